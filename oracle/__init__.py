"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's MFCC + diagonal-GMM hot path, and loaders for the reference's
own C++ compiled in-container (oracle/_ref, built by oracle/Makefile).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (speaker-recognition_b200/) never does and fails loudly without its CUDA library.
"""
