"""Device-side consumers of bulk prediction for AK-MCS (``kadal/reliability_analysis/akmcs.py``):
SURVEY.md section 8 row f3.  The reference predicts the Monte-Carlo population in 10 000-row chunks
(akmcs.py:91-106), keeps G and sigma_G as (N,1) arrays and reduces them with Python list
comprehensions (:267-297).  Here one call streams the population through the GPU and returns only
the scalars; the N predictions never travel back."""
import numpy as np

from .model import model_for
from .prediction import _check_supported, _ensure_predictor, _norm_args


class DevicePopulation:
    """A Monte-Carlo population kept in HBM across AK-MCS iterations.

    ``AKMCS.run`` predicts the SAME population ``init_samp`` after every added sample (akmcs.py:95-106 before the
    loop, :144-155 inside it): uploading its 8 d N bytes once instead of once per iteration removes the PCIe traffic
    the host-pointer path pays every time (480 MB per iteration at BASELINE config 3).  Device memory comes from
    PyTorch (plumbing only); the upload goes through a pinned staging copy.  Pass the object to
    :func:`akmcs_reductions` in place of the numpy array; ``rows(i)`` fetches single rows (``xnew = init_samp[minUloc]``)."""

    def __init__(self, init_samp, device=None):
        import os
        import torch
        x = np.ascontiguousarray(np.atleast_2d(np.asarray(init_samp, dtype=np.float64)))
        if device is None:
            k = int(os.environ.get("KADAL_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            device = "cuda:%d" % (k % max(1, torch.cuda.device_count()))
        self.shape = x.shape
        self.device = torch.device(device)
        self.tensor = torch.empty(x.shape, dtype=torch.float64, device=self.device)
        step = max(1, (64 << 20) // (8 * x.shape[1]))            # 64 MiB pinned staging slices, copy and DMA overlap
        stage = [torch.empty((min(step, len(x)), x.shape[1]), dtype=torch.float64).pin_memory() for _ in range(2)]
        evs = [None, None]
        for k, lo in enumerate(range(0, len(x), step)):
            hi = min(len(x), lo + step)
            b = k & 1
            if evs[b] is not None:
                evs[b].synchronize()
            stage[b][:hi - lo].numpy()[:] = x[lo:hi]
            self.tensor[lo:hi].copy_(stage[b][:hi - lo], non_blocking=True)
            evs[b] = torch.cuda.Event()
            evs[b].record(torch.cuda.current_stream(self.device))
        torch.cuda.synchronize(self.device)

    def __len__(self):
        return self.shape[0]

    def data_ptr(self):
        return self.tensor.data_ptr()

    def rows(self, idx):
        return self.tensor[idx].cpu().numpy()


def akmcs_reductions(init_samp, KrigInfo, group=None):
    """Learning-function and stopping quantities of one AK-MCS iteration.

    Returns a dict with the reference's names: ``minU``, ``minUloc`` (lfucalc :281-285; ``xnew`` is
    ``init_samp[minUloc]``), ``Pf`` (pfcalc :267-271), ``cov`` (covpf :273-279), ``stop_criteria``
    (stopcrit :287-297) and the raw counts.  Under ``torch.distributed`` every rank passes its own
    shard of the population (``parallel.shard_bounds`` order) and all ranks get the global result.
    """
    _check_supported(KrigInfo, None, None)
    m = model_for(KrigInfo)
    _ensure_predictor(m, KrigInfo)
    normtype, na, nb = _norm_args(KrigInfo)
    if isinstance(init_samp, DevicePopulation) or (hasattr(init_samp, "data_ptr") and getattr(init_samp, "is_cuda", False)):
        # resident population (DevicePopulation or a float64 CUDA tensor of shape [N, d]): kb200_akmcs_reduce_dev
        shape = tuple(init_samp.shape)
        if len(shape) != 2 or shape[1] != m.d:
            raise ValueError("population has shape %r, model has %d columns" % (shape, m.d))
        if not isinstance(init_samp, DevicePopulation):
            import torch
            if init_samp.dtype != torch.float64 or not init_samp.is_contiguous():
                raise ValueError("a device-resident population must be a contiguous float64 tensor")
            torch.cuda.current_stream(init_samp.device).synchronize()     # the library runs on its own stream
        min_u, arg, cnt = m.akmcs_reduce_dev(init_samp.data_ptr(), shape[0], normtype, na, nb)
        nmc = shape[0]
    else:
        x = np.atleast_2d(np.asarray(init_samp, dtype=float))
        min_u, arg, cnt = m.akmcs_reduce(x, normtype, na, nb)
        nmc = x.shape[0]
    from . import parallel
    world, rank = parallel._world(group)
    if world > 1:
        import torch
        import torch.distributed as dist
        dev = parallel.comm_device(group)
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        sizes[rank] = nmc
        dist.all_reduce(sizes, group=group)
        offset = int(sizes[:rank].sum().item())
        rec = torch.tensor([[min_u, float(arg + offset)]], dtype=torch.float64, device=dev)
        allrec = torch.empty((world, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allrec, rec, group=group)
        c = torch.tensor(cnt, dtype=torch.int64, device=dev)
        dist.all_reduce(c, group=group)
        allrec = allrec.cpu().numpy()
        best = min(range(world), key=lambda r: (not np.isnan(allrec[r, 0]), allrec[r, 0], allrec[r, 1]))
        min_u, arg = float(allrec[best, 0]), int(allrec[best, 1])
        cnt = [int(v) for v in c.cpu().tolist()]
        nmc = int(sizes.sum().item())
    pf = cnt[0] / nmc
    cov = 1000 if pf == 0 else float(np.sqrt((1 - pf) / (pf * nmc)))
    stop = 100 if pf == 0 else (cnt[1] / nmc - cnt[2] / nmc) / pf
    return {"minU": min_u, "minUloc": arg, "Pf": pf, "cov": cov, "stop_criteria": stop,
            "nGless": cnt[0], "n_minus": cnt[1], "n_plus": cnt[2], "nmc": nmc}
