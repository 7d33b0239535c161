"""Device-side consumers of bulk prediction for AK-MCS (``kadal/reliability_analysis/akmcs.py``):
SURVEY.md section 8 row f3.  The reference predicts the Monte-Carlo population in 10 000-row chunks
(akmcs.py:91-106), keeps G and sigma_G as (N,1) arrays and reduces them with Python list
comprehensions (:267-297).  Here one call streams the population through the GPU and returns only
the scalars; the N predictions never travel back."""
import numpy as np

from .model import model_for
from .prediction import _check_supported, _ensure_predictor, _norm_args


def akmcs_reductions(init_samp, KrigInfo, group=None):
    """Learning-function and stopping quantities of one AK-MCS iteration.

    Returns a dict with the reference's names: ``minU``, ``minUloc`` (lfucalc :281-285; ``xnew`` is
    ``init_samp[minUloc]``), ``Pf`` (pfcalc :267-271), ``cov`` (covpf :273-279), ``stop_criteria``
    (stopcrit :287-297) and the raw counts.  Under ``torch.distributed`` every rank passes its own
    shard of the population (``parallel.shard_bounds`` order) and all ranks get the global result.
    """
    _check_supported(KrigInfo, None, None)
    x = np.atleast_2d(np.asarray(init_samp, dtype=float))
    m = model_for(KrigInfo)
    _ensure_predictor(m, KrigInfo)
    normtype, na, nb = _norm_args(KrigInfo)
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
