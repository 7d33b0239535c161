"""World-size-2 (gloo, CPU) tests of the multi-GPU sharding logic: contiguous candidate /
point shards, gather order, arg-min identical to the single-process result.  The evaluation
function is injected (a deterministic numpy stand-in); the CUDA path is exercised by the
``-m gpu`` tests and by ``bench.py --gpus N``."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_lik(Xpop, KrigInfo, trainvar):
    nll = np.sin(Xpop).sum(1) + (1.0 if trainvar else 0.0)
    info = (Xpop[:, 0] > 0.9).astype(np.int32)
    return nll, info


def _fake_pred(x, KrigInfo, outputs):
    return {"pred": x.sum(1) * 2.0, "s": np.abs(x[:, 0]), "stats": np.zeros(2)}


def _worker(rank, world, port, P, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from kadal_b200 import parallel
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        pop = rng.uniform(-1, 1, (P, 3))
        nll, info = parallel.likelihood_batch_sharded(pop, {}, False, eval_fn=_fake_lik)
        i, v = parallel.argmin_sharded(pop, {}, True, eval_fn=_fake_lik)
        x = rng.uniform(-1, 1, (2 * P + 1, 3))
        full = parallel.predict_sharded(x, {}, ("pred", "s"), eval_fn=_fake_pred)
        mine = parallel.predict_sharded(x, {}, ("pred",), gather=False, eval_fn=_fake_pred)
        # the device-side gather of the NCCL path (unequal shards padded, one all_gather_into_tensor), here on CPU tensors
        import torch
        lo, hi = parallel.shard_bounds(len(x), world, rank)
        g = parallel._gather_device(torch.from_numpy(_fake_pred(x[lo:hi], {}, None)["pred"]), len(x), None).numpy()
        assert np.array_equal(g, full["pred"])
        gi = parallel._gather_device(torch.from_numpy(_fake_lik(pop, {}, False)[1][slice(*parallel.shard_bounds(P, world, rank))]),
                                     P, None).numpy()
        assert np.array_equal(gi, info) and gi.dtype == np.int32
        q.put((rank, nll, info, i, v, full["pred"], full["s"], mine["pred"], parallel.shard_bounds(len(x), world, rank)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P", [7, 16, 1])
def test_sharded_population_and_prediction_match_single_process(P):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    pop = rng.uniform(-1, 1, (P, 3))
    ref_nll, ref_info = _fake_lik(pop, {}, False)
    x = rng.uniform(-1, 1, (2 * P + 1, 3))
    ref_pred = _fake_pred(x, {}, None)
    for rank, nll, info, i, v, pred, s, mine, (lo, hi) in res:
        assert np.array_equal(nll, ref_nll) and np.array_equal(info, ref_info)
        assert i == int(np.argmin(ref_nll + 1.0)) and v == float((ref_nll + 1.0)[i])
        assert np.array_equal(pred, ref_pred["pred"]) and np.array_equal(s, ref_pred["s"])
        assert np.array_equal(mine, ref_pred["pred"][lo:hi])


def test_shard_bounds_cover_everything():
    from kadal_b200.parallel import shard_bounds
    for n in (0, 1, 5, 512, 10 ** 7 + 3):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


# ------------------------------------------------------------------ AK-MCS reductions across ranks
class _FakeAkModel:
    """Stands in for the device model: the per-shard reductions kb200_akmcs_reduce returns, in numpy."""

    def akmcs_reduce(self, x, normtype, na, nb):
        g, s = x[:, 0], np.abs(x[:, 1]) + 0.1
        with np.errstate(invalid="ignore"):
            u = np.abs(g) / s
        arg = int(np.argmin(u)) if not np.isnan(u).any() else int(np.flatnonzero(np.isnan(u))[0])
        return float(u[arg]), arg, [int((g <= 0).sum()), int((g - 1.96 * s <= 0).sum()), int((g + 1.96 * s <= 0).sum())]


def _ak_population(with_nan):
    rng = np.random.default_rng(11)
    x = rng.standard_normal((1001, 2))
    x[700, 0] = 1e-9          # the global minimum of U lives in the second shard
    if with_nan:
        x[900, 0] = np.nan    # np.argmin puts NaN first
    return x


def _ak_worker(rank, world, port, with_nan, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from kadal_b200 import parallel, reliability
    reliability.model_for = lambda K: _FakeAkModel()
    reliability._ensure_predictor = lambda m, K: None
    reliability._check_supported = lambda K, a, b: None
    reliability._norm_args = lambda K: (0, None, None)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        x = _ak_population(with_nan)
        lo, hi = parallel.shard_bounds(len(x), world, rank)
        q.put((rank, reliability.akmcs_reductions(x[lo:hi], {})))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("with_nan", [False, True])
def test_akmcs_reductions_combine_across_ranks(with_nan):
    """reliability.akmcs_reductions under data parallelism: every rank reduces its shard on its GPU (stand-in here), the
    scalars are combined -- global arg-min with the shard offset, NaN first as np.argmin, counts added -- and every rank
    gets what one process computes on the whole population (akmcs.py:267-297)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ak_worker, args=(r, 2, port, with_nan, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = _ak_population(with_nan)
    mu, arg, cnt = _FakeAkModel().akmcs_reduce(x, 0, None, None)
    assert arg == (900 if with_nan else 700)
    pf = cnt[0] / len(x)
    for rank, r in res:
        assert r["minUloc"] == arg and (r["minU"] == mu or (np.isnan(mu) and np.isnan(r["minU"])))
        assert [r["nGless"], r["n_minus"], r["n_plus"]] == cnt and r["nmc"] == len(x)
        assert r["Pf"] == pf and r["stop_criteria"] == (cnt[1] / len(x) - cnt[2] / len(x)) / pf
