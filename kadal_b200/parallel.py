"""Multi-GPU sharding of the hot path: one process per GPU (``torch.distributed``).

The path shards naturally: hyper-parameter candidates and prediction /
Monte-Carlo points are independent units, split contiguously over the ranks; the only
exchange is the final gather of P (or npred) doubles, NCCL over NVLink on the GPU box and
gloo in the CPU tests.  There is no data-path collective inside the kernels.
"""
import numpy as np


def shard_bounds(n, world, rank):
    """Contiguous, balanced [lo, hi) of ``n`` units for ``rank`` of ``world`` (the first
    ``n % world`` ranks get one extra unit)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist


def _world(group=None):
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def comm_device(group=None):
    """Where tensors of a collective are staged: the rank's own GPU for NCCL -- chosen by the same rule as the
    compute path ($KADAL_B200_DEVICE, else $LOCAL_RANK), not torch's current device, which is cuda:0 on every
    rank unless the caller has set it -- and the host for gloo."""
    import os
    import torch
    if _dist().get_backend(group) == "nccl":
        k = int(os.environ.get("KADAL_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        return "cuda:%d" % (k % max(1, torch.cuda.device_count()))
    return "cpu"


def all_gather_rows(local, n_total, group=None, device=None):
    """Gather the per-rank slices (``shard_bounds`` order) of a 1-D / 2-D float64 or int32 array
    into the full array on every rank.  Uses the group's backend: tensors are staged on
    ``device`` ('cuda:k' for NCCL, cpu for gloo)."""
    import torch
    dist = _dist()
    world, rank = _world(group)
    local = np.ascontiguousarray(local)
    if world == 1:
        return local
    if device is None:
        device = comm_device(group)
    counts = [shard_bounds(n_total, world, r) for r in range(world)]
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    maxc = max(hi - lo for lo, hi in counts)
    pad = np.zeros((maxc,) + local.shape[1:], dtype=local.dtype)
    pad[:local.shape[0]] = local
    t = torch.from_numpy(pad).to(device)
    out = torch.empty((world * maxc,) + tuple(local.shape[1:]), dtype=t.dtype, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    out = out.cpu().numpy().reshape((world, maxc) + local.shape[1:])
    return np.concatenate([out[r, :hi - lo] for r, (lo, hi) in enumerate(counts)], axis=0)


def _nccl(group):
    d = _dist()
    return d.is_available() and d.is_initialized() and d.get_backend(group) == "nccl"


def _gather_device(local_t, n_total, group):
    """all_gather of per-rank 1-D device tensors of unequal length (``shard_bounds`` order) ON THE DEVICE: padded to
    the longest shard, one ``all_gather_into_tensor`` over NVLink, trimmed with device slicing.  Returns the full
    device tensor; nothing touches host memory."""
    import torch
    dist = _dist()
    world, rank = _world(group)
    counts = [shard_bounds(n_total, world, r) for r in range(world)]
    maxc = max(hi - lo for lo, hi in counts)
    pad = local_t
    if local_t.shape[0] != maxc:
        pad = torch.zeros(maxc, dtype=local_t.dtype, device=local_t.device)
        pad[:local_t.shape[0]] = local_t
    out = torch.empty(world * maxc, dtype=local_t.dtype, device=local_t.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(hi - lo == maxc for lo, hi in counts):
        return out
    return torch.cat([out[r * maxc:r * maxc + (hi - lo)] for r, (lo, hi) in enumerate(counts)])


def likelihood_batch_sharded(Xpop, KrigInfo, trainvar=False, group=None, eval_fn=None):
    """``likelihood_batch`` with the candidates split over the ranks of ``group``: every rank
    evaluates its contiguous slice on its own GPU and all ranks return the full
    ``(nll[P], info[P])``.  ``eval_fn`` (default: the CUDA ``likelihood_batch``) is injectable
    for the CPU (gloo) tests of the sharding logic.

    Over NCCL the shard is evaluated with device pointers (``kb200_lik_batch_dev`` on the rank's torch stream), the
    P / world doubles are gathered on the device and the full vector comes to the host once.  The factorisation
    schedule is chosen from the GLOBAL population size, so that every candidate gets the same bits as in a
    single-GPU call on the whole population (the two schedules round differently) and the
    arg-min is the same for every world size."""
    Xpop = np.atleast_2d(np.asarray(Xpop, dtype=float))
    world, rank = _world(group)
    lo, hi = shard_bounds(len(Xpop), world, rank)
    if eval_fn is None and world > 1 and _nccl(group):
        import torch
        from .likelihood_func import _fixed_nugget
        from .model import model_for
        m = model_for(KrigInfo)
        dev = torch.device(comm_device(group))
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev)
            x_t = torch.from_numpy(np.ascontiguousarray(Xpop[lo:hi])).to(dev, non_blocking=False)
            nll_t = torch.empty(hi - lo, dtype=torch.float64, device=dev)
            inf_t = torch.empty(hi - lo, dtype=torch.int32, device=dev)
            if hi > lo:
                m.set_schedule_population(len(Xpop))
                try:
                    m.lik_batch_dev(x_t.data_ptr(), hi - lo, Xpop.shape[1], trainvar is True, _fixed_nugget(KrigInfo),
                                    nll_t.data_ptr(), inf_t.data_ptr(), st.cuda_stream)
                finally:
                    m.set_schedule_population(0)
            nll = _gather_device(nll_t, len(Xpop), group)
            info = _gather_device(inf_t, len(Xpop), group)
            return nll.cpu().numpy(), info.cpu().numpy()
    if eval_fn is None:
        from .likelihood_func import likelihood_batch as eval_fn
    if hi > lo:
        nll, info = eval_fn(Xpop[lo:hi], KrigInfo, trainvar)
    else:
        nll, info = np.zeros(0), np.zeros(0, dtype=np.int32)
    nll = all_gather_rows(np.asarray(nll, dtype=np.float64), len(Xpop), group)
    info = all_gather_rows(np.asarray(info, dtype=np.int32), len(Xpop), group)
    return nll, info


def predict_sharded(x, KrigInfo, outputs=("pred",), group=None, gather=True, eval_fn=None):
    """``predict_arrays`` with the points split over the ranks.  With ``gather=False`` every
    rank keeps only its slice (what AK-MCS / Sobol reductions want); otherwise all ranks
    return the full arrays.  Over NCCL the gather runs on the device (the shard's results are copied up once,
    ``all_gather_into_tensor`` over NVLink, one device->host copy of the full array)."""
    if eval_fn is None:
        from .prediction import predict_arrays as eval_fn
        default_fn = True
    else:
        default_fn = False
    x = np.atleast_2d(np.asarray(x, dtype=float))
    world, rank = _world(group)
    lo, hi = shard_bounds(len(x), world, rank)
    res = eval_fn(x[lo:hi], KrigInfo, outputs) if hi > lo else {k: np.zeros(0) for k in outputs}
    if not gather or world == 1:
        return res
    if default_fn and _nccl(group):
        import torch
        dev = torch.device(comm_device(group))
        out = {}
        with torch.cuda.device(dev):
            for k in outputs:
                t = torch.from_numpy(np.ascontiguousarray(np.asarray(res[k], dtype=np.float64).ravel())).to(dev)
                out[k] = _gather_device(t, len(x), group).cpu().numpy()
        return out
    return {k: all_gather_rows(np.asarray(res[k], dtype=np.float64).ravel(), len(x), group) for k in outputs}


def argmin_sharded(Xpop, KrigInfo, trainvar=False, group=None, eval_fn=None):
    """Index and value of the best candidate of a sharded population (first minimum, as
    ``np.argmin`` on the full vector -- identical on every rank and for every world size)."""
    nll, _ = likelihood_batch_sharded(Xpop, KrigInfo, trainvar, group, eval_fn)
    i = int(np.argmin(nll))
    return i, float(nll[i])
