"""End-to-end drop-in test: the reference's facade (ModelInterface.enroll/train/predict, src/gui/interface.py) on
synthetic speakers, with the float64 oracle pipeline as the checker for the predicted labels."""
import numpy as np
import pytest

from oracle import gmm_ref, mfcc_ref
from speaker_recognition_b200 import synth
from speaker_recognition_b200.interface import ModelInterface

pytestmark = pytest.mark.gpu


def speaker_audio(spk, seconds, seed):
    # distinct spectral envelopes per speaker: different f0 region via the seed
    return synth.audio(int(16000 * seconds), seed=1000 * spk + seed)


def test_enroll_train_predict_dump_load(tmp_path):
    m = ModelInterface()
    names = ["ann", "bob", "cyd", "dee"]
    for i, nm in enumerate(names):
        for k in range(3):
            m.enroll(nm, 16000, speaker_audio(i, 4, k))
    m.train(quiet=True)
    assert sorted(m.gmmset.y) == sorted(names) and m.gmmset.gmms[0].get_nr_mixtures() == 32
    tests = [(nm, speaker_audio(i, 3, 0) if False else speaker_audio(i, 4, 1)[8000:56000]) for i, nm in enumerate(names)]
    preds = [m.predict(16000, sig) for _, sig in tests]
    assert preds == names
    # the float64 oracle (reference arithmetic) agrees on every label given the same trained models
    for (nm, sig), pred in zip(tests, preds):
        feat = mfcc_ref.extract(16000, sig)
        sc = [gmm_ref.score_all(*g.get_params(), feat, gmm_ref.MODE_LOGDOMAIN) for g in m.gmmset.gmms]
        assert m.gmmset.y[int(np.argmax(sc))] == pred
    path = str(tmp_path / "model.out")
    m.dump(path)
    m2 = ModelInterface.load(path)
    assert [m2.predict(16000, sig) for _, sig in tests] == names
    assert m.predict(16000, np.zeros(100, dtype=np.int16)) is None      # extractor exception -> None (interface.py:89-93)
