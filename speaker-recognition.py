#!/usr/bin/env python
"""Drop-in for the reference's src/speaker-recognition.py, running on the B200 CUDA path.
See speaker-recognition_b200/cli.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from speaker_recognition_b200.cli import main  # noqa: E402

if __name__ == "__main__":
    main()
