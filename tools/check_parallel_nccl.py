"""torchrun, NCCL: the sharding helpers of kadal_b200.parallel / reliability / sensitivity against the same
calls on one rank.  Deliberately does NOT call torch.cuda.set_device: the helpers must pick the rank's GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_parallel_nccl.py"""
import os, sys, copy, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import kadal_b200 as kb
from kadal_b200 import parallel, reliability, sensitivity
from oracle import kriging_oracle as ko      # builds the KrigInfo only (test infrastructure)
dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(3)
n, d = 200, 6
X = rng.uniform(-1, 1, (n, d)); y = np.sin(X.sum(1, keepdims=True)) + 0.3 * X[:, :1] ** 2 - 0.4
info = ko.make_kriginfo(X, y)
pop = rng.uniform(-0.5, 1.0, (37, d))
nll, inf = parallel.likelihood_batch_sharded(pop, info, False)
ref, rinf = kb.likelihood_batch(pop, info, False)           # every rank also evaluates everything itself
assert np.array_equal(nll, ref) and np.array_equal(inf, rinf), "sharded likelihood differs"
i, v = parallel.argmin_sharded(pop, info, False)
assert i == int(np.argmin(ref)) and v == ref[i]
kb.likelihood(pop[i].copy(), info, "all", False)
xp = rng.uniform(-1, 1, (100003, d))
full = parallel.predict_sharded(xp, info, ("pred", "s"))
one = kb.predict_arrays(xp, info, ("pred", "s"))
assert np.array_equal(full["pred"], one["pred"]) and np.array_equal(full["s"], one["s"]), "sharded predict differs"
lo, hi = parallel.shard_bounds(len(xp), world, rank)
a = reliability.akmcs_reductions(xp[lo:hi], info, group=dist.group.WORLD)
b = reliability.akmcs_reductions(xp, info, group=None) if world == 1 else None
U = np.abs(one["pred"]) / one["s"]
assert a["minUloc"] == int(np.argmin(U)) and abs(a["minU"] - U.min()) <= 1e-12 * U.min(), (a["minUloc"], int(np.argmin(U)))
assert abs(a["Pf"] - np.mean(one["pred"] <= 0)) < 1e-15
# the same population resident in HBM as row shards (kb200_akmcs_reduce_dev): identical scalars on every rank
dp = reliability.DevicePopulation(xp[lo:hi])
a2 = reliability.akmcs_reductions(dp, info, group=dist.group.WORLD)
assert a2 == a, (a2, a)
# Saltelli sums with A, B drawn on the device, rows sharded: equal (to summation order) to one rank doing all rows
sets = [(k,) for k in range(d)]
N = 200_003
lo2, hi2 = -np.ones(2 * d), np.ones(2 * d)
g = sensitivity.sobol_sums_generated(N, info, sets, lo2, hi2, group=dist.group.WORLD)
if g is not None:
    from kadal_b200 import sampling
    _, S = sampling.sampling("sobolnew", 2 * d, N, result="real", upbound=hi2, lobound=lo2)
    r0, r1 = parallel.shard_bounds(N, world, rank)
    h = sensitivity.sobol_sums(np.ascontiguousarray(S[r0:r1, :d]), np.ascontiguousarray(S[r0:r1, d:]), info, sets,
                               group=dist.group.WORLD)
    for u, v in zip(g, h):
        assert np.array_equal(u, v), "device-generated Saltelli sums differ from the host-generated plan"
print("rank", rank, "of", world, "on", parallel.comm_device(), "ok: argmin", i, "Pf", a["Pf"], "minUloc", a["minUloc"], flush=True)
dist.barrier()
dist.destroy_process_group()
