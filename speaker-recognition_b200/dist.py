"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU
tests) -- and nothing else crosses the fabric.  The path shards three ways:

  * utterance-sharded predict (bench.py): ranks own disjoint utterances; no collective on the data path, labels are
    gathered at the end if the caller wants them in one place            -> shard_range / gather_labels
  * frame-sharded UBM EM (BASELINE config 3): every rank keeps N/world frames resident, runs the E-step kernels on
    its shard, ONE all-reduce(sum) of the float64 statistics vector per iteration, identical float64 M-step on every
    rank (deterministic, so no broadcast)                                 -> em_train
  * speaker-sharded MAP enrolment: names are dealt to the ranks (shard_range), every rank adapts the UBM to ITS speakers in one
    batched launch sequence, the resulting models are exchanged as text at the end -- no collective on the data path
                                                                           -> enroll_sharded
  * speaker-sharded predict (BASELINE config 4): every rank scores the same features against its S/world speakers,
    all-gather of the per-rank [B, S_local] sums, global argmax           -> speaker_sharded_argmax

The compute callables default to the C ABI device entry points (srb200_em_estep_dev / srb200_em_mstep_dev); the
CPU tests inject NumPy stand-ins to exercise exactly this control flow and the collective layout under gloo.
"""
import numpy as np


def shard_range(n, rank, world):
    """contiguous [lo, hi) of n items for `rank` (sizes differ by at most one)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist


def world():
    d = _dist()
    if d.is_available() and d.is_initialized():
        return d.get_rank(), d.get_world_size()
    return 0, 1


def all_reduce_sum_(t, group=None):
    """in-place sum across ranks of a torch tensor (the EM statistics vector); no-op without a process group"""
    d = _dist()
    if d.is_available() and d.is_initialized() and d.get_world_size(group) > 1:
        d.all_reduce(t, op=d.ReduceOp.SUM, group=group)
    return t


def gather_labels(local_labels, n_total, group=None):
    """utterance-sharded predict: every rank contributes the labels of its shard_range; returns the full list on all ranks"""
    d = _dist()
    rank, ws = world()
    if ws == 1:
        return list(local_labels)
    out = [None] * ws
    d.all_gather_object(out, list(local_labels), group=group)
    full = []
    for r in range(ws):
        lo, hi = shard_range(n_total, r, ws)
        assert len(out[r]) == hi - lo
        full.extend(out[r])
    return full


def speaker_sharded_argmax(local_sums, group=None):
    """local_sums: torch tensor [B, S_local] (speakers shard_range(S, rank, world), in rank order).
    Returns (best index in the global speaker numbering [B], gathered sums [B, S]) on every rank."""
    import torch
    d = _dist()
    rank, ws = world()
    if ws == 1:
        return torch.argmax(local_sums, dim=1), local_sums
    # all_gather needs equal shapes: shard sizes differ by at most one, so pad every shard to the largest with -inf
    S_loc = local_sums.shape[1]
    sizes = [torch.zeros(1, dtype=torch.int64, device=local_sums.device) for _ in range(ws)]
    d.all_gather(sizes, torch.tensor([S_loc], dtype=torch.int64, device=local_sums.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    S_max = max(sizes)
    padded = torch.full((local_sums.shape[0], S_max), float("-inf"), dtype=local_sums.dtype, device=local_sums.device)
    padded[:, :S_loc] = local_sums
    parts = [torch.empty_like(padded) for _ in range(ws)]
    d.all_gather(parts, padded, group=group)
    full = torch.cat([p_[:, :n] for p_, n in zip(parts, sizes)], dim=1)
    return torch.argmax(full, dim=1), full      # torch.argmax returns the first maximum, like Python's max()


def enroll_sharded(names, feats, fit_many, group=None):
    """MAP enrolment of many speakers over the ranks (SURVEY.md s8e "MAP enrolment of many speakers: speakers over GPUs, no
    collective").  names: list (same on every rank, same order); feats: dict name -> frame matrix -- a rank only needs the entries
    of its own shard; fit_many(names_local, [feats...]) -> list of picklable models (e.g. GMM text dumps) for those names.
    Returns {name: model} for ALL names on every rank (one all_gather_object of the small model blobs at the end)."""
    d = _dist()
    rank, ws = world()
    lo, hi = shard_range(len(names), rank, ws)
    mine = list(names[lo:hi])
    models = fit_many(mine, [feats[n] for n in mine]) if mine else []
    assert len(models) == len(mine)
    if ws == 1:
        return dict(zip(mine, models))
    out = [None] * ws
    d.all_gather_object(out, list(zip(mine, models)), group=group)
    full = {}
    for part in out:
        full.update(dict(part))
    return full


class DeviceEM(object):
    """The default compute backend of em_train: the C ABI device entry points on this rank's frame shard."""

    def __init__(self, gmm, Xd, n_local, ld, dim, ubm=None):
        import torch
        from . import _lib
        self._lib, self.L, self.torch = _lib, _lib.lib(), torch
        self.gmm, self.ubm = gmm, ubm
        self.Xd, self.n, self.ld, self.dim = Xd, int(n_local), int(ld), int(dim)
        K = self.L.get_nr_mixtures(gmm.gmm)
        self.stats = torch.zeros(self.L.srb200_em_stats_len(K, dim), dtype=torch.float64, device=Xd.device)

    def estep(self):
        st = self.torch.cuda.current_stream().cuda_stream
        self._lib.check(self.L.srb200_em_estep_dev(self.gmm.gmm, self.Xd.data_ptr(), self.n, self.ld, self.dim, self.stats.data_ptr(), st),
                        "em_estep_dev")
        return self.stats

    def mstep(self, stats, n_total, min_covar, mode):
        st = self.torch.cuda.current_stream().cuda_stream
        self._lib.check(self.L.srb200_em_mstep_dev(self.gmm.gmm, stats.data_ptr(), int(n_total), float(min_covar), int(mode),
                                                   self.ubm.gmm if self.ubm is not None else None, st), "em_mstep_dev")

    def finish(self):
        self._lib.check(self.L.srb200_gmm_sync_from_device(self.gmm.gmm), "sync_from_device")


def em_train(backend, n_total, nr_iteration, min_covar=1e-3, threshold=0.01, mode=0, group=None, verbosity=0):
    """Frame-sharded EM with the reference trainer's control flow (GMMTrainerBaseline::train, gmm.cc:613-652):
    iterate; on odd iterations monitor the total log-likelihood of the UPDATED model and stop when
    |dll|/|ll| < threshold and dll < threshold.  The E-step of iteration i+1 already yields the likelihood of the model
    after iteration i (statistics tail), so monitoring costs no extra pass: the check is made at the top of iteration
    i+1, before its M-step is applied.  Returns (iterations applied, last monitored ll)."""
    last_ll = -np.finfo(np.float64).max
    ll = None
    iters = 0
    for i in range(nr_iteration + 1):
        need_check = i >= 1 and ((i - 1) % 2 == 1)          # model after iteration i-1, i-1 odd (gmm.cc:622)
        if i == nr_iteration and not need_check:
            break
        stats = backend.estep()
        all_reduce_sum_(stats, group)                        # the ONE collective of the path
        if need_check:
            ll = float(stats[-2].item())
            if verbosity >= 1:
                print("iter %d: ll %f" % (i - 1, ll))
            ll_diff = ll - last_ll
            if abs(ll_diff) / abs(ll) < threshold and ll_diff < threshold:
                break
            last_ll = ll
        if i == nr_iteration:
            break
        backend.mstep(stats, n_total, min_covar, mode)
        iters += 1
    backend.finish()
    return iters, ll
