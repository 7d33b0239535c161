#!/usr/bin/env python
"""bench.py — headline benchmark of the KADAL Kriging hot path on B200.

    python bench.py --gpus 1 --steps 10 --warmup 3            # our arm (CUDA, C ABI)
    python bench.py --impl reference --steps 3 --warmup 1     # reference CPU arm (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): likelihood evals/sec at n=1000, d=10 (configs[1]: CMA-ES population
of 512 theta candidates per GPU); a "step" is one pass of the batched likelihood over one
population.  Weak scaling: every rank evaluates its own 512-candidate population (restarts /
populations are independent) and the NLL vectors are all-gathered with NCCL.
The same JSON line carries the predicted-points/sec figures of configs[2] (n=200, d=6,
pred + s over Monte-Carlo points) under "predict".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_PER_GPU = 512
N_C2, D_C2 = 1000, 10


# ------------------------------------------------------------------ workload (synthetic)
def c2_data():
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (N_C2, D_C2))
    y = np.sin(X.sum(1))[:, None] + 0.1 * rng.standard_normal((N_C2, 1))
    rng.uniform(-1, 1, D_C2)            # known-answer candidate (keeps the stream aligned with SURVEY §8d)
    rng.uniform(-5, 5, (512, D_C2))     # population A (reference search box)
    popB = rng.uniform(-0.5, 1.5, (512, D_C2))   # population B (informative region)
    return X, y, popB


def population(rank):
    X, y, popB = c2_data()
    if rank == 0:
        return X, y, popB
    return X, y, np.random.default_rng(1000 + rank).uniform(-0.5, 1.5, (P_PER_GPU, D_C2))


def c3_data():
    rng = np.random.default_rng(1)
    X = rng.uniform(-1, 1, (200, 6))
    y = (np.sin(X.sum(1)) + 0.3 * X[:, 0] ** 2)[:, None] + 3.0
    return X, y, np.full(6, 0.2)


def krig_info(X, y, kernel=("gaussian",), pls=None, n_princomp=False):
    """Minimal KrigInfo dict of an ordinary-Kriging model on inputs already scaled to [-1, 1]^d (the keys
    likelihood() / prediction() read) -- built here so that the GPU arm imports nothing of oracle/."""
    n, d = X.shape
    return dict(nvar=d, nsamp=n, X=X, X_norm=X, y=y, F=np.ones((n, 1)), kernel=list(kernel), nkernel=len(kernel),
                standardization=True, normtype="default", norm_y=False, type="kriging" if pls is None else "kpls",
                n_princomp=n_princomp, nugget=-6, idx=np.zeros((1, d)), TrendOrder=0, lb=-np.ones(d), ub=np.ones(d),
                **({"plscoeff": pls} if pls is not None else {}))


def headline_config(n_gpus):
    """`config` of the JSON line -- the same dict for both arms (the driver compares them key by key)."""
    return {"workload": "C2 batched likelihood sweep: n=1000, d=10, CMA-ES population 512 theta candidates per GPU",
            "n": N_C2, "d": D_C2, "population_per_gpu": P_PER_GPU, "kernel": "gaussian", "nugget": -6,
            "candidates": "popB: log10(theta) ~ U(-0.5, 1.5) (informative region, cond up to ~9e8)",
            "l2": "working set %.1f GB of factor tiles per step >> 126 MB L2 (inputs larger than L2)" % (P_PER_GPU * 36 * 131072 / 1e9),
            "parallelism": "dp%d (candidates sharded per GPU, NCCL all-gather of NLL)" % n_gpus}


# ------------------------------------------------------------------ algorithmic work
def chol_panel_flops(n, P):
    """Algorithmic flops of all chol_panel_kernel launches of one likelihood pass: for tile
    (i,j), i>j: GEMM update 2*r_i*c_j*(128 j) + triangular solve r_i*c_j^2 (true, unpadded
    row/column counts)."""
    nt = (n + 127) // 128
    size = lambda t: min(128, n - 128 * t)
    fl = 0.0
    for j in range(nt):
        for i in range(j + 1, nt):
            fl += 2.0 * size(i) * size(j) * (128 * j) + size(i) * size(j) ** 2
    return fl * P


def chol_diag_flops(n, P, nq=2):
    """Algorithmic flops of the diagonal-tile kernel: SYRK update c_j^2*(128 j), POTF2
    c_j^3/3, fused forward substitution of [y F] 2*nq*c_j*(128 j) + nq*c_j^2.  The block
    inverse the kernel also forms is an implementation choice, not algorithmic work."""
    nt = (n + 127) // 128
    size = lambda t: min(128, n - 128 * t)
    fl = 0.0
    for j in range(nt):
        c = size(j)
        fl += c * c * (128.0 * j) + c ** 3 / 3.0 + 2.0 * nq * c * (128.0 * j) + nq * c * c
    return fl * P


def assemble_flops(n, d, P, c_exp=20):
    """Lower triangle incl. diagonal: per entry d*(sub, mul, div, add) + exp."""
    return P * (n * (n + 1) / 2.0) * (4.0 * d + c_exp)


def lik_flops_per_eval(n, d, p=1, c_exp=20):
    """n^3/3 + n^2 (1.5 d + p + 1) + (n^2/2) c_exp."""
    return n ** 3 / 3 + n ** 2 * (1.5 * d + p + 1) + (n ** 2 / 2) * c_exp


def predict_flops_per_point(n, d, var=True, c_exp=20):
    f = n * (3 * d + c_exp) + 2 * n
    return f + (n * n + 2 * n if var else 0)


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region: NVML polled from a thread every
    10 ms (nvidia-smi -lms needs ~1 s to start, longer than the timed region); nvidia-smi once as a
    fallback."""
    NAMES = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self.t, self.stop_flag, self.h = index, [], None, False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, mx, pw, rs))
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        if self.h is None:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,power.draw",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                sm, mx, pw = [float(v) for v in out.strip().split(",")]
                return {"sm_mhz": sm, "sm_max_mhz": mx, "power_w_max": pw, "samples": 1, "reasons": [],
                        "note": "NVML unavailable: one nvidia-smi sample after the timed region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.stop_flag = True
        self.t.join(timeout=1)
        sm = [r[0] for r in self.rows]
        reasons = sorted(nm for nm, bit in self.NAMES.items() if any(r[3] & bit for r in self.rows))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[1] for r in self.rows) if sm else None,
                "power_w_max": max(r[2] for r in self.rows) if sm else None, "samples": len(sm), "reasons": reasons}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# ------------------------------------------------------------------ reference arm (CPU)
def all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is entitled to all host threads."""
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=os.cpu_count())
        return max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] or [1])
    except Exception:
        return None


def run_reference(args, rank):
    """The reference's own CPU implementation of the path, timed on this box's host cores.
    KADAL is pure Python (numpy/scipy) and /root/reference does not travel to the GPU box, so
    the timed code is oracle/kriging_oracle.py — the operation-by-operation restatement that
    tests/test_oracle.py pins bit-for-bit to the live reference (kind = "port")."""
    if rank != 0:
        return
    all_host_threads()
    from oracle import kriging_oracle as ko
    X, y, pop = c2_data()
    info = ko.make_kriginfo(X, y)
    per_step = 2
    k = 0
    for _ in range(max(args.warmup, 1)):
        ko.likelihood(pop[k % 512].copy(), info, "default", False)
        k += 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(per_step):
            ko.likelihood(pop[k % 512].copy(), info, "default", False)
            k += 1
    dt = time.perf_counter() - t0
    val = args.steps * per_step / dt
    cores = os.cpu_count()
    sample = "%d of the 512 popB candidates per step, sequential calls as cma/scipy issue them, eigvals check on, BLAS threads = all cores" % per_step
    line = {
        "impl": "reference", "metric": "likelihood_evals_per_sec", "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": headline_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline(budget_s=14.0):
    all_host_threads()
    from oracle import kriging_oracle as ko
    X, y, pop = c2_data()
    info = ko.make_kriginfo(X, y)
    ko.likelihood(pop[0].copy(), info, "default", False)   # warm BLAS
    t0, k = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s * 0.6 and k < 16:
        ko.likelihood(pop[k].copy(), info, "default", False)
        k += 1
    asis = k / (time.perf_counter() - t0)
    t1, k2 = time.perf_counter(), 0
    while time.perf_counter() - t1 < budget_s * 0.4 and k2 < 16:
        ko.likelihood(pop[k2].copy(), info, "default", False, check_pd=False)
        k2 += 1
    fair = k2 / (time.perf_counter() - t1)
    return {"value": asis, "unit": "evals/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "first %d of the 512 popB candidates, sequential oracle.likelihood calls (as cma/scipy call the "
                      "reference), eigvals check on; with the result-neutral eigvals check off: %.2f evals/s (%d evals)"
                      % (k, fair, k2)}


# ------------------------------------------------------------------ our arm (GPU)
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from kadal_b200 import _capi
    from kadal_b200.model import GpuModel, fp64_probe

    _capi.require_gpu()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # (NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION and WARN: main() has pointed fd 1 at
        # stderr, the JSON line goes to the saved descriptor)
        dist.init_process_group("nccl", device_id=dev)

    X, y, pop = population(rank)
    model = GpuModel(X, y, np.ones((N_C2, 1)), [0], device=local_rank)
    P, nb = pop.shape
    x_dev = torch.from_numpy(pop).to(dev)
    nll_dev = torch.empty(P, dtype=torch.float64, device=dev)
    info_dev = torch.empty(P, dtype=torch.int32, device=dev)
    gathered = torch.empty(world * P, dtype=torch.float64, device=dev) if world > 1 else None
    # a real (non-default) stream: the library launches on the handle it is given, and
    # torch.cuda.Event only sees work of torch's *current* stream -> make them the same one
    tstream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        model.lik_batch_dev(x_dev.data_ptr(), P, nb, False, -6.0, nll_dev.data_ptr(), info_dev.data_ptr(), stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, nll_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # NVML poll thread: samples only while the timed steps run
    launches0 = model.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = model.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel durations (the roofline's denominator): the same steps once more with the
    # candidate groups serialised on one stream, so that the CUDA-event brackets around every
    # launch do not overlap (in the timed region above two groups run concurrently)
    model.set_groups(1)
    step()
    barrier()
    model.profile(True)
    pev0, pev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pev0.record()
    for _ in range(args.steps):
        step()
    pev1.record()
    barrier()
    ms_serial = pev0.elapsed_time(pev1)
    model.profile(False)
    prof = model.profile_get()
    model.set_groups(2)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.float64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    value = world * P * args.steps / (ms * 1e-3)
    nll_host = nll_dev.cpu().numpy()
    assert np.all(np.isfinite(nll_host)) and int((info_dev != 0).sum().item()) == 0

    # ---- end to end through the host-pointer C ABI (pinned host buffers, H2D + D2H inside)
    xh = _capi.pinned_empty((P, nb))
    xh[:] = pop
    nll_h = _capi.pinned_empty((P,))
    info_h = _capi.pinned_empty((P,), dtype=np.int32)
    lib = model.lib

    def e2e_step():
        _capi.check(lib.kb200_lik_batch(model.handle, _capi.ptr(xh), P, nb, 0, -6.0, _capi.ptr(nll_h), _capi.ptr(info_h)))

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = world * P * args.steps / e2e_s
    assert np.array_equal(nll_h, nll_host)

    # ---- sustained: back-to-back generations for >= args.sustain_seconds (a CMA-ES training is hundreds of them)
    sustained = sustained_record(args, rank, local_rank, world, torch, dist, barrier, step, P)
    # ---- strong scaling through the product API (kadal_b200.parallel / reliability / sensitivity), checked in the run
    strong = strong_scaling(args, rank, local_rank, world, torch, dist, barrier) if not args.no_strong else None
    # ---- prediction throughput (configs[2]: n=200, d=6, pred + s over 10^7 points per GPU), device resident + e2e
    pred = predict_bench(args, rank, local_rank, world, torch, dist, barrier)
    extra = extra_configs(args, rank, local_rank, world, torch, dist, barrier) if not args.no_extra else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    dmma_tf, dfma_tf = fp64_probe(local_rank)
    peaks = measured_peaks()
    kernel_ms = {k: round(v[0], 4) for k, v in prof.items() if v[1]}
    tot_ms = sum(v[0] for v in prof.values())
    step_tf = lik_flops_per_eval(N_C2, D_C2) * P * args.steps / (tot_ms * 1e-3) / 1e12
    hbm = peaks["hbm_gbs"] if peaks else 6650.0
    # algorithmic flops per launch class -> achieved TFLOP/s of each kernel
    # the panel kernel now also carries the diagonal tiles j >= 1 (SYRK, POTF2, forward substitution)
    diag0 = P * (128.0 ** 3 / 3.0 + 2 * 128.0 * 128.0)
    nt_c2 = (N_C2 + 127) // 128
    half_path = prof["chol_diag"][1] >= nt_c2 * args.steps      # one diagonal launch per tile column
    if half_path:
        cls_flops = {"assemble": assemble_flops(N_C2, D_C2, P), "chol_diag": chol_diag_flops(N_C2, P),
                     "chol_panel": chol_panel_flops(N_C2, P)}
        cls_name = {"assemble": "corr_tile_kernel (Psi assembly, fp64 FMA pipe + exp)",
                    "chol_diag": "chol_diag_half_kernel (half-SM CTA per diagonal tile: DMMA SYRK over the tile row, fused "
                                 "forward substitution, blocked POTF2 on a packed triangle)",
                    "chol_panel": "chol_panel_half_kernel (half-SM CTAs, two per SM: 64 x 128 block of tile (i,j), fp64 DMMA "
                                  "GEMM + right-looking DMMA TRSM against the fully resident L_jj)"}
    else:
        cls_flops = {"assemble": assemble_flops(N_C2, D_C2, P), "chol_diag": diag0,
                     "chol_panel": chol_panel_flops(N_C2, P) + chol_diag_flops(N_C2, P) - diag0}
        cls_name = {"assemble": "corr_tile_kernel (Psi assembly, fp64 FMA pipe + exp)",
                    "chol_diag": "chol_diag_kernel (diagonal tile 0: blocked in-smem POTF2)",
                    "chol_panel": "chol_panel_kernel (fp64 DMMA GEMM + right-looking DMMA TRSM; CTA (j+1,j) also does the "
                                  "SYRK + blocked POTF2 + forward substitution of diagonal tile j+1)"}
    per_kernel = {}
    for k, fl in cls_flops.items():
        ms_k, cnt_k = prof[k]
        if cnt_k:
            tf = fl * args.steps / (ms_k * 1e-3) / 1e12
            per_kernel[k] = {"tflops": tf, "frac_of_dmma_peak": tf / dmma_tf, "share_of_step": ms_k / tot_ms,
                             "launches": cnt_k, "avg_launch_ms": ms_k / cnt_k}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["share_of_step"])
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    line = {
        "metric": "likelihood_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": headline_config(world),
        "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": world * P * nb * 8,
                "d2h_bytes_per_step": world * P * 12},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": cls_name[dom],
                     "achieved": per_kernel[dom]["tflops"], "peak": dmma_tf, "unit": "TFLOP/s",
                     "frac": per_kernel[dom]["frac_of_dmma_peak"], "traffic": traffic,
                     "peak_source": "kb200_fp64_probe: register-resident mma.sync.m8n8k4.f64 loop measured in this run "
                                    "(MEASURED_PEAKS.json has no FP64 figure; DFMA probe: %.1f TFLOP/s)" % dfma_tf,
                     "launches": per_kernel[dom]["launches"], "avg_launch_ms": per_kernel[dom]["avg_launch_ms"],
                     "share_of_step": per_kernel[dom]["share_of_step"],
                     "per_kernel": per_kernel,
                     "per_kernel_note": "per-kernel times from a second pass of the same steps with the two candidate "
                                        "groups serialised on one stream (%.3f ms/step; the timed region overlaps them: "
                                        "%.3f ms/step)" % (ms_serial / args.steps, ms / args.steps),
                     "whole_step_tflops": lik_flops_per_eval(N_C2, D_C2) * P * args.steps / (ms * 1e-3) / 1e12,
                     "whole_step_frac": lik_flops_per_eval(N_C2, D_C2) * P * args.steps / (ms * 1e-3) / 1e12 / dmma_tf,
                     "hbm_view": {"achieved": P * args.steps * 36 * 131072 * 3 / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                  "note": "factor tiles written once by assembly, read+written once by the factorisation "
                                          "(compulsory traffic); not the binding roof (AI >> 6 flop/B)"}},
        "kernel_ms": kernel_ms,
        "sustained": sustained,
        "strong_scaling": strong,
        "predict": pred,
        "other_configs": extra,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def sustained_record(args, rank, local_rank, world, torch, dist, barrier, step, P):
    """The headline step repeated back to back for >= --sustain-seconds with the clock sampler running: the number a
    training of hundreds of generations sees (this pool's B200s drop from 1965 MHz under seconds-long load,
    MEASURED_PEAKS.json)."""
    if args.sustain_seconds <= 0:
        return None
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_done, block = 0, 50
    t0 = time.perf_counter()
    ev0.record()
    while True:
        for _ in range(block):
            step()
        n_done += block
        torch.cuda.current_stream().synchronize()
        go = torch.tensor([1.0 if time.perf_counter() - t0 < args.sustain_seconds else 0.0], device="cuda:%d" % local_rank)
        if world > 1:
            dist.all_reduce(go, op=dist.ReduceOp.MIN)      # every rank runs the same number of blocks
        if go.item() == 0.0:
            break
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda:%d" % local_rank)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"value": world * P * n_done / (ms * 1e-3), "unit": "evals/s", "steps": n_done, "seconds": ms * 1e-3,
            "ms_per_step": ms / n_done, "clocks": clocks,
            "note": "same step as the headline, %d generations back to back (includes one host sync per %d steps)" % (n_done, block)}


def strong_scaling(args, rank, local_rank, world, torch, dist, barrier):
    """Fixed total work split over the ranks THROUGH THE PRODUCT API, with the results checked inside the run:
      * likelihood: ONE 512-candidate population (C2, popB) via kadal_b200.parallel.likelihood_batch_sharded
        (host arrays in, full nll / info vectors out on every rank: device-side gather, H2D / D2H inside the timing);
        asserted bit-equal to this rank evaluating the whole population alone, arg-min identical;
      * prediction: ONE 10^7-point Monte-Carlo population (C3) resident in HBM as row shards, one AK-MCS iteration via
        kadal_b200.reliability.akmcs_reductions(group=...) (scalars all-reduced); asserted equal to rank 0 reducing the
        whole population alone;
      * C4: the 2e7 Saltelli predictions of one Sobol analysis (N = 909 091, d = 20, A, B, 20 x C_i) via
        kadal_b200.sensitivity.sobol_sums_generated(group=...): rows sharded, drawn on the device, sums all-reduced."""
    import kadal_b200 as kb
    from kadal_b200 import parallel, reliability, sensitivity
    dev = torch.device("cuda", local_rank)
    group = dist.group.WORLD if world > 1 else None
    out = {"n_gpus": world, "scaling": "strong"}

    def wall(fn, reps, warm=1):
        for _ in range(warm):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        barrier()
        dt = (time.perf_counter() - t0) / reps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    # ---- likelihood
    X, y, popB = c2_data()
    info = krig_info(X, y)
    res = {}
    dt = wall(lambda: res.update(v=parallel.likelihood_batch_sharded(popB, info, False, group=group)), max(3, min(args.steps, 10)), warm=2)
    nll, inf = res["v"]
    ref_nll, ref_inf = kb.likelihood_batch(popB, info, False)           # this rank alone, whole population
    assert np.array_equal(nll, ref_nll) and np.array_equal(inf, ref_inf), "sharded likelihood differs from the single-GPU result"
    i, v = parallel.argmin_sharded(popB, info, False, group=group)
    assert i == int(np.argmin(ref_nll)) and v == ref_nll[i]
    out["likelihood"] = {"workload": "C2: ONE population of 512 candidates (popB) split over %d GPU(s), n=1000, d=10" % world,
                         "api": "kadal_b200.parallel.likelihood_batch_sharded (host arrays in / out, device-side gather)",
                         "value": len(popB) / dt, "unit": "evals/s", "ms_per_generation": dt * 1e3,
                         "candidates_per_gpu": len(popB) // world, "argmin": i,
                         "checked": "bit-equal to one GPU evaluating all 512 candidates; arg-min identical"}
    # ---- prediction (AK-MCS iteration over a resident population)
    X3, y3, th3 = c3_data()
    info3 = krig_info(X3, y3 - 3.6)                                      # the limit state G = 0 cuts through the population
    kb.likelihood(th3.copy(), info3, "all", False)
    ntot = int(args.predict_points)
    lo, hi = parallel.shard_bounds(ntot, world, rank)
    # the whole population is a pure function of the row index, so every rank can make its shard (and rank 0 the rest)
    def rows(a, b):
        g = torch.Generator(device=dev)
        blocks = []
        B = 1 << 20
        for b0 in range(a - a % B, b, B):
            g.manual_seed(77 + b0 // B)
            t = (torch.randn(B, 6, generator=g, device=dev, dtype=torch.float64) * 0.5).clamp_(-1, 1)
            blocks.append(t[max(a - b0, 0):min(b - b0, B)])
        return torch.cat(blocks).contiguous()
    shard = rows(lo, hi)
    torch.cuda.synchronize()
    r = {}
    dt = wall(lambda: r.update(v=reliability.akmcs_reductions(shard, info3, group=group)), 3, warm=1)
    got = r["v"]
    # single-rank reference of the same reduction: rank 0 alone over the whole population (no group)
    if rank == 0:
        full = rows(0, ntot) if world > 1 else shard
        torch.cuda.synchronize()
        from kadal_b200.model import model_for
        from kadal_b200.prediction import _norm_args
        m3 = model_for(info3)
        nt_, na_, nb_ = _norm_args(info3)
        mu, arg, cnt = m3.akmcs_reduce_dev(full.data_ptr(), ntot, nt_, na_, nb_)
        assert got["minUloc"] == arg and got["minU"] == mu and [got["nGless"], got["n_minus"], got["n_plus"]] == cnt, \
            "sharded AK-MCS reduction differs from the single-GPU result"
        del full
    out["akmcs_iteration"] = {"workload": "C3: ONE Monte-Carlo population of %d points (n=200, d=6) resident in HBM as row shards over %d GPU(s)" % (ntot, world),
                              "api": "kadal_b200.reliability.akmcs_reductions(resident population, group): pred + s + U-function reductions, scalars all-reduced",
                              "value": ntot / dt, "unit": "pts/s", "ms_per_iteration": dt * 1e3, "Pf": got["Pf"], "minUloc": got["minUloc"],
                              "checked": "minU, arg-min index and the three counts equal to one GPU reducing the whole population"}
    # ---- C4 Sobol analysis, rows drawn on the device
    rng = np.random.default_rng(7)
    X4 = rng.uniform(-1, 1, (800, 20))
    y4 = (np.sin(X4[:, :3].sum(1)) + 0.05 * rng.standard_normal(800))[:, None] + 3.0
    info4 = krig_info(X4, y4)
    kb.likelihood(np.full(20, 0.4), info4, "all", False)
    N4, sets = 909_091, [(i,) for i in range(20)]
    lo2, hi2 = -np.ones(40), np.ones(40)
    r4 = {}
    dt = wall(lambda: r4.update(v=sensitivity.sobol_sums_generated(N4, info4, sets, lo2, hi2, group=group)), 2, warm=1)
    if r4["v"] is not None:
        base, yayc, ybyc = r4["v"]
        fo2 = (base[0] / N4) ** 2
        den = base[1] / N4 - fo2
        s1 = (yayc / N4 - fo2) / den
        out["sobol_analysis"] = {"workload": "C4: Sobol indices via Kriging surrogate, n=800, d=20, N=909091 base rows -> 22 x N = %.2e Saltelli predictions split over %d GPU(s)" % (22 * N4, world),
                                 "api": "kadal_b200.sensitivity.sobol_sums_generated(group): A, B drawn on the device (bit-identical to the host Sobol plan), C_i composed on the device, sums all-reduced",
                                 "value": 22 * N4 / dt, "unit": "pts/s", "ms_per_analysis": dt * 1e3,
                                 "pcie_bytes_per_analysis": 40 * 30 * 8 + 80 * 8 + 20 * 20 + 44 * 8,
                                 "first_order_indices_x1_x2_x3": [float(v) for v in s1[:3]], "max_abs_first_order_x4_x20": float(np.abs(s1[3:]).max())}
    return out


def predict_bench(args, rank, local_rank, world, torch, dist, barrier):
    from kadal_b200 import _capi
    from kadal_b200.model import GpuModel
    dev = torch.device("cuda", local_rank)
    X, y, th = c3_data()
    n, d = X.shape
    m = GpuModel(X, y, np.ones((n, 1)), [0], device=local_rank)
    st = m.fit_state(th, False, -6.0, want_U=False, want_Psi=False)
    assert st["info"] == 0
    npts = int(args.predict_points)
    g = torch.Generator(device=dev)
    g.manual_seed(2 + rank)
    xs = (torch.randn(npts, d, generator=g, device=dev, dtype=torch.float64) * 0.5).clamp_(-1, 1)
    o_pred = torch.empty(npts, dtype=torch.float64, device=dev)
    o_s = torch.empty(npts, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream   # run_ours installed a non-default stream
    assert stream != 0
    lb, ub = -np.ones(d), np.ones(d)
    want = _capi.OUT_PRED | _capi.OUT_S

    def step():
        m.predict_dev(xs.data_ptr(), npts, want, [o_pred.data_ptr(), 0, o_s.data_ptr(), 0, 0, 0, 0],
                      normtype=_capi.NORM_DEFAULT, norm_a=lb, norm_b=ub, stream=stream)

    steps = max(2, min(args.steps, 5))
    for _ in range(3):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    # per-kernel-class times from a second pass of the same steps with the profiler events on: the library then keeps the
    # variance chunks on ONE stream (the timed pass above rotates them over two, overlapping the launch tails)
    m.profile(True)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for _ in range(steps):
        step()
    ev3.record()
    barrier()
    ms_serial = ev2.elapsed_time(ev3)
    m.profile(False)
    prof = m.profile_get()
    # e2e: pinned host points in, pinned host results out
    xh = _capi.pinned_empty((npts, d))
    xh[:] = xs.cpu().numpy()
    out = {"pred": _capi.pinned_empty((npts,)), "s": _capi.pinned_empty((npts,))}
    m.predict(xh, want, _capi.NORM_DEFAULT, lb, ub, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        m.predict(xh, want, _capi.NORM_DEFAULT, lb, ub, out=out)
    barrier()
    e2e_s = (time.perf_counter() - t0) / 2
    if world > 1:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0].item()), float(t[1].item())
    assert np.allclose(out["pred"], o_pred.cpu().numpy(), rtol=0, atol=0)
    pts = world * npts * steps / (ms * 1e-3)
    fl = predict_flops_per_point(n, d, True)
    rec = {"metric": "predicted_points_per_sec", "value": pts, "unit": "pts/s", "n_gpus": world, "scaling": "weak",
           "e2e": {"value": world * npts / e2e_s, "unit": "pts/s", "h2d_bytes_per_step": world * npts * d * 8,
                   "d2h_bytes_per_step": world * npts * 16,
                   "note": "kb200_predict with pinned host points in and pinned host results out, copies inside the timing"},
           "e2e_value": world * npts / e2e_s,
           "config": {"workload": "C3 AK-MCS: n=200, d=6, pred + s over %d Monte-Carlo points per GPU" % npts,
                      "n": n, "d": d, "points_per_gpu": npts, "outputs": "pred + s",
                      "l2": "%.0f MB of points + %.0f MB of results per step >> 126 MB L2" % (npts * d * 8 / 1e6, npts * 16 / 1e6)},
           "steps": steps, "ms_per_step": ms / steps, "kernel_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1]},
           "fp64_tflops_algorithmic": pts * fl / 1e12 / world,
           "hbm_gbs_algorithmic": pts * (8 * d + 16) / 1e9 / world,
           "h2d_bytes_per_step": world * npts * d * 8, "d2h_bytes_per_step": world * npts * 16}
    if rank == 0:
        from kadal_b200.model import fp64_probe
        dmma_tf, _ = fp64_probe(local_rank)
        peaks = measured_peaks()
        k_ms, k_cnt = prof["pred_solve"]
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("predict_tiled")
        except Exception:
            pass
        if k_cnt:
            # dominant kernel class of the step: the variance solve v = L^-1 psi (n^2 + 2 n flops per point); the psi kernel
            # (n (3 d + 20) + 2 n flops per point, FP64 FMA pipe) and the epilogue are in `kernel_ms` and in `whole_frac`
            fl_solve = float(n * n + 2 * n)
            tf = fl_solve * npts * steps / (k_ms * 1e-3) / 1e12
            rec["roofline"] = {"bound": "tensor",
                               "kernel": "variance solve of the tiled predictor: solve_rows_half_kernel (sample tile 0: resident L_00, "
                                         "8-row strips per warp) + chol_panel_half_kernel in prediction mode (sample tile 1: DMMA "
                                         "GEMM update + solve, padded columns of the ragged tile skipped)",
                               "achieved": tf, "peak": dmma_tf, "unit": "TFLOP/s", "frac": tf / dmma_tf, "traffic": traffic,
                               "launches": k_cnt, "avg_launch_ms": k_ms / k_cnt,
                               "flops_per_point": fl_solve, "points_per_launch": npts * steps * (n + 127) // 128 / max(1, k_cnt),
                               "whole_frac": pts / world * fl / 1e12 / dmma_tf,
                               "per_kernel_note": "kernel_ms / achieved from a second pass with the variance chunks on one stream "
                                                  "(%.3f ms per step; the timed pass rotates them over two streams: %.3f ms)" % (ms_serial / steps, ms / steps),
                               "whole_note": "all kernels of the step (psi + solve + epilogue), %d algorithmic flops per point" % fl,
                               "peak_source": "kb200_fp64_probe (DMMA, this run)",
                               "hbm_view": {"achieved": npts * steps * (8 * d + 16) / (ms * 1e-3) / 1e9,
                                            "peak": peaks["hbm_gbs"] if peaks else 6650.0, "unit": "GB/s",
                                            "note": "algorithmic bytes 8 d + 16 per point; arithmetic intensity %.0f flop/B against a machine "
                                                    "balance of ~5.7: the FP64 pipe binds, not HBM (SURVEY.md section 8d)" % (fl / (8 * d + 16))}}
        if world == 1 and not args.no_cpu_baseline:
            rec["cpu_baseline"] = cpu_baseline_predict()
    return rec


def cpu_baseline_predict(budget_s=8.0):
    """The reference's prediction() (oracle port) on this box's host cores: C3 model, ['pred', 's'] on 10 000-point chunks
    as AKMCS.run issues them (akmcs.py:95-106), and the C4 model, 'pred' on 10 000-point chunks (sobol_ind.py:48-58)."""
    all_host_threads()
    from oracle import kriging_oracle as ko
    out = {"unit": "pts/s", "cores": os.cpu_count(), "kind": "port"}
    X, y, th = c3_data()
    info = ko.make_kriginfo(X, y)
    ko.likelihood(th.copy(), info, "all", False, check_pd=False)
    rng = np.random.default_rng(2)
    chunk = np.clip(rng.standard_normal((10_000, 6)) * 0.5, -1, 1)
    ko.prediction(chunk, info, ["pred", "s"])
    t0, k = time.perf_counter(), 0
    while k < 10 and time.perf_counter() - t0 < budget_s * 0.5:
        ko.prediction(chunk, info, ["pred", "s"])
        k += 1
    out["value"] = k * 10_000 / (time.perf_counter() - t0)
    out["sample"] = "%d chunks of 10 000 points, oracle.prediction(chunk, ['pred', 's']) on the C3 model (n=200, d=6)" % k
    rng = np.random.default_rng(7)
    X4 = rng.uniform(-1, 1, (800, 20))
    y4 = (np.sin(X4[:, :3].sum(1)) + 0.05 * rng.standard_normal(800))[:, None] + 3.0
    info4 = ko.make_kriginfo(X4, y4)
    ko.likelihood(np.full(20, 0.4), info4, "all", False, check_pd=False)
    chunk4 = rng.uniform(-1, 1, (10_000, 20))
    t0, k = time.perf_counter(), 0
    while k < 3 and time.perf_counter() - t0 < budget_s * 0.5:
        ko.prediction(chunk4, info4, ["pred"])
        k += 1
    out["c4_value"] = k * 10_000 / (time.perf_counter() - t0)
    out["c4_sample"] = "%d chunks of 10 000 points, oracle.prediction(chunk, ['pred']) on the C4 model (n=800, d=20)" % k
    return out


def extra_configs(args, rank, local_rank, world, torch, dist, barrier):
    """BASELINE configs[3] and [4] on this rank's GPU (rank 0 reports; every rank runs the same work so the
    numbers are per GPU): C4 Sobol surrogate n=800, d=20 'pred' over Saltelli points, C5 KPLS n=4000,
    d=100, h=3: batched likelihood of a small CMA-ES population and EHVI-style pred + SSqr scoring."""
    from kadal_b200 import _capi
    from kadal_b200.model import GpuModel
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    out = {}

    def timed(fn, reps):
        for _ in range(2):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / reps

    # ---- small-n likelihood batches (north star (2): one CTA per matrix held in shared memory): C1 training
    # (n=50, d=2: 7 restarts x 3-point L-BFGS-B stencils = 21 candidates per batch) and C3 training (n=200, d=6, a CMA-ES
    # population of 512), each also through the tile path the same sizes took in round 1 ($KB200_NO_SMALL_LIK=1)
    small = {}
    for tag, n_s, d_s, P_s, seed in (("c1_train_lik", 50, 2, 21, 21), ("n128_lik", 128, 6, 512, 23), ("c3_train_lik", 200, 6, 512, 22)):
        rng = np.random.default_rng(seed)
        Xs = rng.uniform(-1, 1, (n_s, d_s))
        ys = (np.sin(Xs.sum(1)) + 0.3 * Xs[:, 0] ** 2)[:, None] + 3.0
        ms_ = GpuModel(Xs, ys, np.ones((n_s, 1)), [0], device=local_rank)
        pops = rng.uniform(-0.5, 1.0, (P_s, d_s))
        xd = torch.from_numpy(pops).to(dev)
        nll = torch.empty(P_s, dtype=torch.float64, device=dev)
        info = torch.empty(P_s, dtype=torch.int32, device=dev)
        fn = lambda: ms_.lik_batch_dev(xd.data_ptr(), P_s, d_s, False, -6.0, nll.data_ptr(), info.data_ptr(), stream)
        os.environ["KB200_SMALL_LIK_NMAX"] = "256"           # (bench only: n = 200 sits just above the size policy's cut)
        t_small = timed(fn, 20)
        v_small = nll.cpu().numpy().copy()
        os.environ["KB200_NO_SMALL_LIK"] = "1"
        t_tile = timed(fn, 20)
        del os.environ["KB200_NO_SMALL_LIK"]
        del os.environ["KB200_SMALL_LIK_NMAX"]
        v_tile = nll.cpu().numpy()
        assert int((info != 0).sum().item()) == 0 and np.allclose(v_small, v_tile, rtol=1e-7, atol=1e-7)
        fl = lik_flops_per_eval(n_s, d_s) * P_s
        small[tag] = {"workload": "n=%d, d=%d, %d candidates per batch and GPU" % (n_s, d_s, P_s),
                      "kernel": "lik_small_kernel (assembly + blocked POTF2 + forward substitution of one candidate per CTA in shared memory)",
                      "evals_per_s_per_gpu": P_s / (t_small * 1e-3), "ms_per_batch": t_small,
                      "fp64_tflops_algorithmic": fl / (t_small * 1e-3) / 1e12,
                      "tile_path": {"evals_per_s_per_gpu": P_s / (t_tile * 1e-3), "ms_per_batch": t_tile},
                      "default_path": "lik_small_kernel" if n_s <= 192 else "tile path (n > 192: measured equal or faster, DESIGN.md section 4)"}
        ms_.close()
    out["small_n_likelihood"] = small
    # ---- BASELINE configs[0], KrigDemo (KrigDemo.py:22-118): Ordinary Kriging, Gaussian kernel, Branin 2-D, 50 Sobol samples,
    # L-BFGS-B from 7 start points, predict a 100 x 100 grid -- through the drop-in class, wall clock as the user sees it
    if rank == 0:
        try:
            import contextlib
            import io
            from scipy.stats import qmc
            from kadal_b200 import Kriging
            lbk, ubk = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
            Xk = qmc.scale(qmc.Sobol(2, scramble=False).random(51)[1:], lbk, ubk)
            x1, x2 = Xk[:, 0], Xk[:, 1]
            yk = ((x2 - 5.1 / (4 * np.pi ** 2) * x1 ** 2 + 5 / np.pi * x1 - 6) ** 2 + 10 * (1 - 1 / (8 * np.pi)) * np.cos(x1) + 10)[:, None]
            g1, g2 = np.meshgrid(np.linspace(lbk[0], ubk[0], 100), np.linspace(lbk[1], ubk[1], 100))
            xg = np.column_stack([g1.ravel(), g2.ravel()])
            times = []
            for _ in range(3):
                K = {"X": Xk, "y": yk, "lb": lbk, "ub": ubk, "nvar": 2, "problem": "branin", "nugget": -6, "nrestart": 7,
                     "kernel": ["gaussian"], "optimizer": "lbfgsb"}
                with contextlib.redirect_stdout(io.StringIO()):
                    t0 = time.perf_counter()
                    mk = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
                    mk.train(disp=False)
                    t1 = time.perf_counter()
                    pk = mk.predict(xg, ["pred", "SSqr"])
                    t2 = time.perf_counter()
                times.append((t1 - t0, t2 - t1))
            tr, pr = min(t[0] for t in times), min(t[1] for t in times)
            out["c1_krigdemo"] = {"workload": "C1 KrigDemo: n=50, d=2, train (L-BFGS-B, 7 restarts in lock step) + predict a 100 x 100 grid (pred, SSqr), through kadal_b200.Kriging",
                                  "train_ms": tr * 1e3, "predict_10k_ms": pr * 1e3, "likelihood_evals": mk.train_stats["evals"],
                                  "gpu_batches": mk.train_stats["batches"], "NegLnLike": float(mk.KrigInfo["NegLnLike"]),
                                  "reference_on_8_cores_s": {"train": 1.07, "predict_10k": 0.13, "source": "SURVEY.md section 6 (measured in the build container)"}}
        except Exception as exc:      # scipy.stats.qmc missing etc.: the record is optional
            out["c1_krigdemo"] = {"unavailable": str(exc)[:120]}
    # ---- C4
    rng = np.random.default_rng(7)
    X = rng.uniform(-1, 1, (800, 20))
    y = (np.sin(X[:, :3].sum(1)) + 0.05 * rng.standard_normal(800))[:, None] + 3.0
    m = GpuModel(X, y, np.ones((800, 1)), [0], device=local_rank)
    assert m.fit_state(np.full(20, 0.4), False, -6.0, want_U=False, want_Psi=False)["info"] == 0
    npts = 2_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(7 + rank)
    xs = torch.rand(npts, 20, generator=g, device=dev, dtype=torch.float64) * 2 - 1
    o = torch.empty(npts, dtype=torch.float64, device=dev)
    lb, ub = -np.ones(20), np.ones(20)
    ms = timed(lambda: m.predict_dev(xs.data_ptr(), npts, _capi.OUT_PRED, [o.data_ptr(), 0, 0, 0, 0, 0, 0],
                                     normtype=_capi.NORM_DEFAULT, norm_a=lb, norm_b=ub, stream=stream), 3)
    pts = npts / (ms * 1e-3)
    out["c4_sobol_pred"] = {"workload": "C4: n=800, d=20, 'pred' over %d Saltelli points per GPU" % npts, "pts_per_s_per_gpu": pts,
                            "ms": ms, "fp64_tflops_algorithmic": pts * predict_flops_per_point(800, 20, False) / 1e12,
                            "hbm_gbs_algorithmic": pts * (8 * 20 + 8) / 1e9}
    o2 = torch.empty(npts, dtype=torch.float64, device=dev)
    npv = 1_000_000
    ms = timed(lambda: m.predict_dev(xs.data_ptr(), npv, _capi.OUT_PRED | _capi.OUT_S, [o.data_ptr(), 0, o2.data_ptr(), 0, 0, 0, 0],
                                     normtype=_capi.NORM_DEFAULT, norm_a=lb, norm_b=ub, stream=stream), 2)
    out["c4_pred_s"] = {"workload": "C4 model (n=800, d=20): pred + s over %d points per GPU (general tiled path, n > 256)" % npv,
                        "pts_per_s_per_gpu": npv / (ms * 1e-3), "ms": ms,
                        "fp64_tflops_algorithmic": npv / (ms * 1e-3) * predict_flops_per_point(800, 20, True) / 1e12}
    m.close()
    del xs, o, o2
    # ---- C5
    rng = np.random.default_rng(11)
    n, d, h = 4000, 100, 3
    X = rng.uniform(-1, 1, (n, d))
    y = (X[:, :5].sum(1) ** 2)[:, None] + 3.0
    pls = rng.standard_normal((d, h)) / 10
    th = np.array([0.3, 0.1, -0.1])
    m = GpuModel(X, y, np.ones((n, 1)), [0], pls=pls, device=local_rank)
    P5 = 16
    pop = th[None, :] + np.random.default_rng(3).uniform(-0.2, 0.2, (P5, h))
    xd = torch.from_numpy(pop).to(dev)
    nll = torch.empty(P5, dtype=torch.float64, device=dev)
    info = torch.empty(P5, dtype=torch.int32, device=dev)
    ms = timed(lambda: m.lik_batch_dev(xd.data_ptr(), P5, h, False, -6.0, nll.data_ptr(), info.data_ptr(), stream), 2)
    assert int((info != 0).sum().item()) == 0
    out["c5_kpls_lik"] = {"workload": "C5: KPLS n=4000, d=100, h=3, population of %d candidates per GPU" % P5,
                          "evals_per_s_per_gpu": P5 / (ms * 1e-3), "ms": ms,
                          "chol_tflops": n ** 3 / 3.0 * P5 / (ms * 1e-3) / 1e12}
    assert m.fit_state(th, False, -6.0, want_U=False, want_Psi=False)["info"] == 0
    nc = 100_000
    xs = torch.rand(nc, d, generator=g, device=dev, dtype=torch.float64) * 2 - 1
    o1 = torch.empty(nc, dtype=torch.float64, device=dev)
    o2 = torch.empty(nc, dtype=torch.float64, device=dev)
    lb, ub = -np.ones(d), np.ones(d)
    ms = timed(lambda: m.predict_dev(xs.data_ptr(), nc, _capi.OUT_PRED | _capi.OUT_SSQR,
                                     [o1.data_ptr(), o2.data_ptr(), 0, 0, 0, 0, 0], normtype=_capi.NORM_DEFAULT,
                                     norm_a=lb, norm_b=ub, stream=stream), 2)
    pts = nc / (ms * 1e-3)
    out["c5_kpls_score"] = {"workload": "C5: pred + SSqr (EHVI inputs) over %d candidates per GPU" % nc, "pts_per_s_per_gpu": pts,
                            "ms": ms, "fp64_tflops_algorithmic": pts * (n * n + n * (3 * d + 6 * h + 20)) / 1e12}
    # the batch a MOBO generation really scores: n_pop = 1500 candidates per objective (EHVIcomputation.py:169-172),
    # right-looking sweep (few point tiles, 32 sample tiles); with the left-looking sweep forced beside it
    nc2 = 1500
    fn = lambda: m.predict_dev(xs.data_ptr(), nc2, _capi.OUT_PRED | _capi.OUT_SSQR, [o1.data_ptr(), o2.data_ptr(), 0, 0, 0, 0, 0],
                               normtype=_capi.NORM_DEFAULT, norm_a=lb, norm_b=ub, stream=stream)
    ms_rl = timed(fn, 5)
    os.environ["KB200_PRED_RL"] = "0"
    ms_ll = timed(fn, 5)
    del os.environ["KB200_PRED_RL"]
    out["c5_kpls_score_generation"] = {"workload": "C5: pred + SSqr of ONE generation of %d candidates per GPU (n=4000, d=100, h=3)" % nc2,
                                       "ms": ms_rl, "pts_per_s_per_gpu": nc2 / (ms_rl * 1e-3),
                                       "fp64_tflops_algorithmic": nc2 / (ms_rl * 1e-3) * (n * n + n * (3 * d + 6 * h + 20)) / 1e12,
                                       "left_looking_sweep_ms": ms_ll}
    m.close()
    out["c5_ehvi3d"] = ehvi3d_bench(rank, local_rank, world, torch, timed, dev)
    return out


def ehvi3d_bench(rank, local_rank, world, torch, timed, dev):
    """C5's 3-objective scoring step: kb200_ehvi3d over 1000 candidates and a 100-point front (the largest batch
    the reference's static tables hold), both batch semantics, device-resident inputs; at N=1 the reference's own
    C++ (oracle/_ref/libkmac_ref.so, built from /root/reference in the build container) is timed beside it."""
    import ctypes
    from kadal_b200 import _capi
    lib = _capi.load()
    rng = np.random.default_rng(5)
    k, npop = 100, 1000
    v = np.abs(rng.standard_normal((k, 3)))
    P = -10 + 9.5 * v / np.linalg.norm(v, axis=1, keepdims=True)
    mu, sd = rng.uniform(-10.2, -0.5, (npop, 3)), rng.uniform(0.3, 2.5, (npop, 3))
    ref = np.array([-10.0, -10.0, -10.0])
    dP, dmu, dsd = (torch.from_numpy(a).to(dev) for a in (P, mu, sd))
    o = torch.empty(npop, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    res = {"workload": "C5 scoring: 3-objective EHVI (KMAC) of %d candidates over a %d-point front per GPU" % (npop, k)}
    for mode, name in ((0, "reference_semantics"), (1, "independent")):
        def run():
            _capi.check(lib.kb200_ehvi3d_dev(local_rank, dP.data_ptr(), k, _capi.ptr(ref), dmu.data_ptr(), dsd.data_ptr(),
                                             3, 1, npop, mode, 1.0, o.data_ptr(), ctypes.c_void_p(stream)))
        ms = timed(run, 5)
        res[name] = {"ms": ms, "candidates_per_s_per_gpu": npop / (ms * 1e-3), "positive": int((o > 0).sum().item())}
    if rank == 0 and world == 1:
        try:
            from oracle import kmac_ref_loader as kr   # checker / CPU baseline only
            if kr.available():
                t0 = time.perf_counter()
                kr.ehvi3d_multi(P, ref, mu, sd)
                t_batch = time.perf_counter() - t0
                t0 = time.perf_counter()
                for i in range(20):
                    kr.ehvi3d_multi(P, ref, mu[i:i + 1], sd[i:i + 1])
                t_one = (time.perf_counter() - t0) / 20
                res["cpu_reference"] = {"kind": "reference", "cores": 1, "batch_ms": t_batch * 1e3,
                                        "one_candidate_alone_ms": t_one * 1e3,
                                        "sample": "the batch once (its early exit skips most of the work); 20 candidates one by one"}
        except Exception as exc:   # the baseline is optional
            res["cpu_reference"] = {"unavailable": str(exc)[:100]}
    return res


_REAL_STDOUT = None


def emit(obj):
    """The one JSON line, on the process's real stdout."""
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries write to fd 1 behind Python's back (NCCL's version banner, torchrun notices): point fd 1 at stderr
    # for the whole run and keep the real stdout for the JSON line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--predict-points", type=float, default=1e7, help="C3 Monte-Carlo points per GPU (BASELINE configs[2]: 10^7)")
    ap.add_argument("--sustain-seconds", type=float, default=5.0, help="length of the sustained-throughput record (0: skip)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling records")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C4 / C5 sub-benchmarks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, stdout=_REAL_STDOUT))   # the child gets the real stdout as its fd 1
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
