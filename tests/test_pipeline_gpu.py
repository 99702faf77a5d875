"""End-to-end drop-in test: the reference's facade (ModelInterface.enroll/train/predict, src/gui/interface.py) on
synthetic speakers, with the float64 oracle pipeline as the checker for the predicted labels."""
import os

import numpy as np
import pytest

from oracle import gmm_ref, lpc_ref, mfcc_ref
from speaker_recognition_b200 import _lib, synth
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
        feat = np.concatenate((mfcc_ref.extract(16000, sig), lpc_ref.extract(16000, sig)), axis=1)   # mix_feature
        sc = [gmm_ref.score_all(*g.get_params(), feat, gmm_ref.MODE_LOGDOMAIN) for g in m.gmmset.gmms]
        assert m.gmmset.y[int(np.argmax(sc))] == pred
    path = str(tmp_path / "model.out")
    m.dump(path)
    m2 = ModelInterface.load(path)
    assert [m2.predict(16000, sig) for _, sig in tests] == names
    assert m.predict(16000, np.zeros(100, dtype=np.int16)) is None      # extractor exception -> None (interface.py:89-93)


def test_batch_scorer_matches_per_utterance_path_and_oracle():
    """pipeline.BatchScorer (device-resident, chunked H2D overlap) == per-utterance public API == float64 oracle argmax"""
    from speaker_recognition_b200.feature import MFCC
    from speaker_recognition_b200.gmm.pygmm import GMM
    from speaker_recognition_b200.gmmset import GMMSet
    from speaker_recognition_b200.pipeline import BatchScorer
    S, K, D = 12, 32, 39
    W, M, Sg = synth.models(S, K, D, seed0=1000)
    gs = GMMSet(gmm_order=K)
    for i in range(S):
        gs.gmms.append(GMM(K).set_params(W[i], M[i], Sg[i]))
        gs.y.append("s%d" % i)
    lens = [40000, 9000, 33333, 2561 + 512, 70001, 12000, 8191, 25000, 16000, 48000]   # 10 utterances -> 4 ragged sub-batches
    sigs = [synth.audio(n, seed=300 + i) for i, n in enumerate(lens)]
    bs = BatchScorer(gs, fs=16000, nd=2)
    labels = bs.predict(sigs)
    sums = bs.d_sums.cpu().numpy()
    for i, s in enumerate(sigs):
        feat = MFCC.extract(16000, s, nd=2)
        per_utt = np.array(gs.predict_one_scores(feat))
        assert np.max(np.abs(per_utt - sums[i]) / np.abs(per_utt)) < 1e-5
        ref_feat = mfcc_ref.extract39(16000, s)
        ex = [gmm_ref.score_all(W[k], M[k], Sg[k], ref_feat, gmm_ref.MODE_LOGDOMAIN) for k in range(S)]
        assert labels[i] == "s%d" % int(np.argmax(ex))
        assert np.max(np.abs(np.array(ex) - sums[i]) / np.abs(np.array(ex))) < 2e-3     # features differ at 1e-4 -> sums at ~1e-3


def test_streaming_predictor_equals_the_oracle_on_the_trailing_window():
    """config-5 semantics (gui.py:179-190): each tick's label == the ORACLE's label (NumPy restatement of MFCC.py + exact float64
    GMM scores) on the trailing 1.5 s window of every stream -- checked on several ticks, not only the last one"""
    from speaker_recognition_b200.gmm.pygmm import GMM
    from speaker_recognition_b200.gmmset import GMMSet
    from speaker_recognition_b200.streaming import StreamingPredictor
    S, K, D, B, hop = 6, 16, 39, 5, 400
    W, M, Sg = synth.models(S, K, D, seed0=2000)
    gs = GMMSet(gmm_order=K)
    for i in range(S):
        gs.gmms.append(GMM(K).set_params(W[i], M[i], Sg[i]))
        gs.y.append(i)
    sp = StreamingPredictor(gs, n_streams=B, fs=16000, window_s=1.5, nd=2)
    n_ticks = 24000 // hop + 7
    audio = np.stack([synth.audio(hop * n_ticks, seed=700 + b) for b in range(B)])
    checked = 0
    for t in range(n_ticks):
        out = sp.push(audio[:, t * hop:(t + 1) * hop])
        assert (out is None) == ((t + 1) * hop < 24000)
        if out is None or (t % 3 and t != n_ticks - 1):
            continue
        for b in range(B):
            win = audio[b, (t + 1) * hop - 24000:(t + 1) * hop]
            feat = mfcc_ref.extract39(16000, win)
            assert feat.shape == (90, 39)
            ex = np.array([gmm_ref.score_all(W[k], M[k], Sg[k], feat, gmm_ref.MODE_LOGDOMAIN) for k in range(S)])
            srt = np.sort(ex)
            if srt[-1] - srt[-2] > 1e-3 * abs(srt[-1]):          # a near-tie may legitimately flip at the 1e-4 feature tolerance
                assert int(out[b]) == int(np.argmax(ex)), (t, b)
                checked += 1
            got = sp.sums[b].cpu().numpy()
            assert np.max(np.abs(got - ex) / np.abs(ex)) < 2e-3     # features differ at 1e-4 -> sums at ~1e-3
    assert checked >= 3 * B
    with pytest.raises(AssertionError):
        sp.push(audio[:, :24000])                                   # hop must be shorter than the window


def test_rejection_is_one_launch_and_matches_the_oracle():
    """predict_one_with_rejection (src/testbench/gmmset.py:69-84): the UBM rides in the speakers' launch as model S + 1; the
    accept / reject decision and the margin equal the float64 oracle's"""
    from speaker_recognition_b200.gmm.pygmm import GMM
    from speaker_recognition_b200.gmmset import GMMSet
    L = _lib.lib()
    S, K, D = 5, 32, 13
    wu, mu, su = synth.model(K, D, seed=4000, floor_frac=0.0)
    su = su * 1.5                                                    # a broad "world" model
    ubm = GMM(K).set_params(wu, mu, su)
    rng = np.random.default_rng(3)
    spk = [(wu, mu + 0.8 * rng.standard_normal(mu.shape), su * 0.6) for _ in range(S)]      # MAP-like: moved means, tighter
    gs = GMMSet(ubm=ubm, reject_threshold=0.8)
    for i, (w, m, s) in enumerate(spk):
        gs.gmms.append(GMM(K).set_params(w, m, s))
        gs.y.append("spk%d" % i)
    cases = [synth.draw(*spk[2], 400, seed=1), synth.draw(*spk[4], 250, seed=2), synth.draw(wu, mu, su, 400, seed=3),
             synth.draw(wu, mu + 2.0, su, 300, seed=4)]
    outcomes = set()
    for x in cases:
        n0 = L.srb200_score_launch_count()
        got = gs.predict_one_with_rejection(x)
        assert L.srb200_score_launch_count() - n0 == 1               # speakers and UBM in ONE scoring launch
        sc = np.array([gmm_ref.score_all(w, m, s, x, gmm_ref.MODE_LOGDOMAIN) for (w, m, s) in spk]) / len(x)
        margin = sc.max() - gmm_ref.score_all(wu, mu, su, x, gmm_ref.MODE_LOGDOMAIN) / len(x)
        assert abs(margin - gs.reject_threshold) > 1e-3              # the case is not on the decision boundary
        want = None if margin < gs.reject_threshold else "spk%d" % int(np.argmax(sc))
        assert got == want
        outcomes.add(got is None)
        # the plain prediction ignores the UBM even though it is in the packed set
        assert gs.predict_one(x) == "spk%d" % int(np.argmax(sc))
    assert outcomes == {True, False}                                  # both branches exercised
    # the batched device path ranks only the speakers as well
    from speaker_recognition_b200.pipeline import BatchScorer
    assert L.srb200_gmmset_size(gs._ensure_set()) == S + 1


def test_cli_enroll_and_predict(tmp_path, capsys):
    """speaker-recognition.py -t enroll / -t predict on wav files written to disk (src/speaker-recognition.py:52-90)"""
    from scipy.io import wavfile
    from speaker_recognition_b200 import cli
    names = ["alice", "bert", "carol"]
    for i, nm in enumerate(names):
        d = tmp_path / "train" / nm
        d.mkdir(parents=True)
        for k in range(2):
            wavfile.write(str(d / ("%d.wav" % k)), 16000, speaker_audio(i + 10, 4, k))
    test_dir = tmp_path / "test"
    test_dir.mkdir()
    for i, nm in enumerate(names):
        wavfile.write(str(test_dir / ("%s_x.wav" % nm)), 16000, speaker_audio(i + 10, 4, 1)[4000:52000])
    model = str(tmp_path / "model.out")
    cli.main(["-t", "enroll", "-i", "%s/train/*" % tmp_path, "-m", model])
    out = cli.task_predict(str(test_dir / "*.wav"), model)
    assert len(out) == 3
    for f, label in out:
        assert os.path.basename(f).split("_")[0] == label
    assert "->" in capsys.readouterr().out


def test_ragged_batches_do_not_allocate_after_warm_up():
    """200 batches of random utterance counts and lengths through the public batched API: once the largest batch has been seen, the
    device-side free memory does not move (no cudaMalloc in libsrb200's shape / tile-table caches -- their slots are rewritten in
    place -- nor in the pipeline's buffers) and the per-call latency stays flat"""
    import time
    import torch
    from speaker_recognition_b200.gmm.pygmm import GMM
    from speaker_recognition_b200.gmmset import GMMSet
    from speaker_recognition_b200.pipeline import BatchScorer
    S, K, D = 8, 32, 39
    W, M, Sg = synth.models(S, K, D, seed0=3100)
    gs = GMMSet(gmm_order=K)
    for i in range(S):
        gs.gmms.append(GMM(K).set_params(W[i], M[i], Sg[i]))
        gs.y.append(i)
    bs = BatchScorer(gs, fs=16000, nd=2)
    rng = np.random.default_rng(11)
    Bmax, Lmax = 6, 40000
    pool = torch.from_numpy(synth.audio(Bmax * Lmax, seed=5)).pin_memory()

    def run(lengths):
        n = int(sum(lengths))
        t0 = time.perf_counter()
        lab = bs.predict_pcm(pool[:n], lengths)
        return time.perf_counter() - t0, lab.numpy().copy()
    for _ in range(3):
        run([Lmax] * Bmax)                                       # warm-up with the largest batch
    shapes = [[int(rng.integers(4000, Lmax)) for _ in range(int(rng.integers(1, Bmax + 1)))] for _ in range(40)]
    for sh in shapes[:20]:
        run(sh)                                                   # fills the 16 / 256 cache slots once
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    lat = []
    for i in range(200):
        sh = shapes[i % len(shapes)] if i % 3 else [int(rng.integers(4000, Lmax)) for _ in range(int(rng.integers(1, Bmax + 1)))]
        dt, lab = run(sh)
        lat.append(dt)
        assert lab.shape == (len(sh),)
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info()[0] == free0                  # nothing was allocated (or freed) on the device
    lat = np.sort(np.array(lat))
    assert lat[int(0.95 * len(lat))] < 4.0 * lat[len(lat) // 2]  # no allocation / synchronisation spikes
    # and the answers do not depend on what was cached: a fresh scorer gives the same labels for the last shape
    bs2 = BatchScorer(gs, fs=16000, nd=2)
    n = int(sum(sh))
    assert np.array_equal(bs2.predict_pcm(pool[:n], sh).numpy(), lab)


def test_two_steps_in_flight_give_the_synchronous_answers():
    """submit() / result(): step i+1 is enqueued (its copies on the copy stream, into the other buffer slot) before step i's labels
    are collected; every step's labels equal those of the synchronous predict_pcm on the same input"""
    import torch
    from speaker_recognition_b200.gmm.pygmm import GMM
    from speaker_recognition_b200.gmmset import GMMSet
    from speaker_recognition_b200.pipeline import BatchScorer
    S, K, D = 6, 32, 39
    W, M, Sg = synth.models(S, K, D, seed0=4200)
    gs = GMMSet(gmm_order=K)
    for i in range(S):
        gs.gmms.append(GMM(K).set_params(W[i], M[i], Sg[i]))
        gs.y.append(i)
    bs = BatchScorer(gs, fs=16000, nd=2)
    rng = np.random.default_rng(3)
    steps = []
    for i in range(7):
        lengths = [int(rng.integers(6000, 30000)) for _ in range(int(rng.integers(1, 6)))]
        if i in (2, 3):
            lengths = [30000] * 40                                   # a chunked step (B >= 32) between unchunked ones
        pcm = torch.from_numpy(synth.audio(int(sum(lengths)), seed=100 + i)).pin_memory()
        steps.append((pcm, lengths))
    want = [bs.predict_pcm(p, l).numpy().copy() for p, l in steps]
    got = []
    pending = bs.submit(*steps[0])
    for p, l in steps[1:]:
        nxt = bs.submit(p, l)
        got.append(bs.result(pending).numpy().copy())
        pending = nxt
    got.append(bs.result(pending).numpy().copy())
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
