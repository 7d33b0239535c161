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
