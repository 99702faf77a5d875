"""Command line tool -- a Python 3 mirror of src/speaker-recognition.py:20-100 (same flags, same flow):

    speaker-recognition.py -t enroll  -i "./bob/ ./mary/ ./person*" -m model.out
    speaker-recognition.py -t predict -i "./*.wav"                  -m model.out

Wav files in each input directory are labelled with the directory's basename; wildcard inputs are globbed.
Everything below the facade (features, training, scoring) runs through the CUDA library.
"""
import argparse
import glob
import itertools
import os
import sys

from .interface import ModelInterface


def read_wav(fname):
    """src/gui/utils.py:10-13"""
    from scipy.io import wavfile
    fs, signal = wavfile.read(fname)
    assert len(signal.shape) == 1, "Only Support Mono Wav File!"
    return fs, signal


def get_args(argv=None):
    desc = "Speaker Recognition Command Line Tool"
    epilog = """
Wav files in each input directory will be labeled as the basename of the directory.
Note that wildcard inputs should be *quoted*, and they will be sent to glob.glob module.

Examples:
    Train (enroll a list of person named person*, and mary, with wav files under corresponding directories):
    ./speaker-recognition.py -t enroll -i "./bob/ ./mary/ ./person*" -m model.out

    Predict (predict the speaker of all wav files):
    ./speaker-recognition.py -t predict -i "./*.wav" -m model.out
"""
    parser = argparse.ArgumentParser(description=desc, epilog=epilog, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument('-t', '--task', help='Task to do. Either "enroll" or "predict"', required=True)
    parser.add_argument('-i', '--input', help='Input Files(to predict) or Directories(to enroll)', required=True)
    parser.add_argument('-m', '--model', help='Model file to save(in enroll) or use(in predict)', required=True)
    return parser.parse_args(argv)


def task_enroll(input_dirs, output_model):
    m = ModelInterface()
    input_dirs = [os.path.expanduser(k) for k in input_dirs.strip().split()]
    dirs = itertools.chain(*(glob.glob(d) for d in input_dirs))
    dirs = [d for d in dirs if os.path.isdir(d)]
    if len(dirs) == 0:
        print("No valid directory found!")
        sys.exit(1)
    training_stats = []
    for d in dirs:
        label = os.path.basename(d.rstrip('/'))
        wavs = glob.glob(d + '/*.wav')
        if len(wavs) == 0:
            print("No wav file found in {0}".format(d))
            continue
        print("Label '{0}' has files: {1}".format(label, ', '.join(wavs)))
        total_len = 0
        for wav in wavs:
            fs, signal = read_wav(wav)
            print("   File '{}' has frequency={} and length={}".format(wav, fs, len(signal)))
            total_len += len(signal)
            m.enroll(label, fs, signal)
        training_stats.append((label, total_len))
    print("--------------------------------------------")
    for label, total_len in training_stats:
        print("Total length of training data for '{}' is {}".format(label, total_len))
    print("For best accuracy, please make sure all labels have similar amount of training data!")
    m.train()
    m.dump(output_model)


def task_predict(input_files, input_model):
    m = ModelInterface.load(input_model)
    out = []
    for f in glob.glob(os.path.expanduser(input_files)):
        fs, signal = read_wav(f)
        label = m.predict(fs, signal)
        print(f, '->', label)
        out.append((f, label))
    return out


def main(argv=None):
    args = get_args(argv)
    if args.task == 'enroll':
        task_enroll(args.input, args.model)
    elif args.task == 'predict':
        task_predict(args.input, args.model)
    else:
        print('task must be "enroll" or "predict"')
        sys.exit(2)


if __name__ == '__main__':
    main()
