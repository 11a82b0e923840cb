#!/usr/bin/env python
"""bench.py -- headline benchmark of the fvGP training hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (ours;  torchrun launches N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): LML + gradient evaluations per second on config C2 -- single-task GP,
3-D inputs, N = 50 000, default anisotropic Matern-3/2 kernel, dense FP64 Cholesky.  One "step"
is one evaluation of log_likelihood(theta_k) AND neg_log_likelihood_gradient(theta_k) at a new
theta_k (synthetic data of SURVEY.md section 8d, seeded).

  value        evals/s with x / y / noise resident in HBM, timed with CUDA events on the launching
               stream, barrier + synchronize on both sides, max over ranks.
  e2e          same metric through the public API with HOST buffers: every step re-uploads x from
               pinned host memory and reads LML and gradient back (copies inside the timed region).
  roofline     dominant kernel = the DMMA GEMM behind POTRF + POTRI: algorithmic N^3 flop / (CUDA-event
               time of the potrf + potri phases); peak = DMMA issue rate measured live in this run
               (MEASURED_PEAKS.json has no FP64 figure).  roofline_kfill: the K-assembly kernel against HBM.
  cpu_baseline the numpy/scipy oracle port of the reference algorithm on the host cores, on a bounded
               sample (smaller N), extrapolated with N^3 (the reference gradient needs (3H+3) 8 N^2 bytes
               = 300 GB at N = 50k and cannot run); reported, not the target.
N > 1: hyperparameter proposals are independent evaluations (MCMC / DE populations), so each rank
evaluates its own theta sequence on a replica of the data -- no data-path collective ("weak").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv and "TORCHELASTIC_RUN_ID" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
    # torchrun exports OMP_NUM_THREADS=1 when it starts more than one rank; the CPU arm (rank 0 only) is meant
    # to use every host core, and OpenBLAS sizes its thread pool when numpy is first imported
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def synthetic_c2(n, seed=2):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 3))
    y = np.sin(5 * x[:, 0]) * np.cos(3 * x[:, 1]) + x[:, 2] + 0.1 * rng.standard_normal(n)
    return x, y, np.full(n, 1e-2)


def synthetic_c4(n, seed=4):
    """uniform points in the unit cube, Morton-ordered (the permutation is part of the
    workload definition and is applied before the CPU baseline too), ~100 neighbours inside the support."""
    from fvgp_b200.utils import morton_order
    rng = np.random.default_rng(seed)
    x = rng.random((n, 3))
    x = np.ascontiguousarray(x[morton_order(x)])
    y = np.sin(8 * np.linalg.norm(x, axis=1)) + 0.1 * rng.standard_normal(n)
    return x, y, np.full(n, 1e-2)


def theta_c4(k, n, rank=0):
    base = 0.029 * (1e6 / n) ** (1.0 / 3.0)            # keeps ~102 nnz/row when n is reduced for tests
    return np.array([1.0, base, base, base]) * np.array([1.0] + [1.0 + 0.005 * ((k + 3 * rank) % 10 - 5)] * 3)


def theta_k(k, rank=0):
    return np.array([1.0, .3, .4, .5]) * (1.0 + 0.02 * ((k + 7 * rank) % 20))


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region.  In-process NVML (the library nvidia-smi itself
    reads; the recipe's clocks line: clocks.sm, clocks.max.sm, clocks_event_reasons.*): forking nvidia-smi from a
    process with a CUDA context costs tens of milliseconds per sample and perturbs short timed regions (the C1
    population step is ~4 ms).  Falls back to the nvidia-smi subprocess when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = self.handle = None
        self.mode = os.environ.get("FVGP_BENCH_SAMPLER", "nvml")      # nvml | smi | none (A/B of the sampler's own cost)
        try:
            if self.mode != "nvml":
                raise RuntimeError("NVML sampler disabled")
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.source = "nvml"
        except Exception:
            self.source = "nvidia-smi"

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
        return [str(sm), str(mx), "0"] + ["Active" if (r & b) else "Not Active" for b in bits]

    def run(self):
        while not self.stop_flag and self.mode != "none":
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                         timeout=5).stdout
                    self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.05 if self.nvml is not None else 0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = sorted({self.NAMES[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": self.source}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 for nproc > 1; the CPU arm is meant to use every host core."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    return os.cpu_count()


def cpu_baseline_step(x, y, noise, theta, orc):
    """Reference algorithm (stacked LU solves of KV against dK/dtheta) on the host cores."""
    lml = orc.dense_log_likelihood(x, y, theta, noise)
    grad = orc.dense_neg_log_likelihood_gradient(x, y, theta, noise, economical=False)
    return lml, grad


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (numpy/scipy oracle port; the reference itself is
    pure Python and absent on the GPU box) timed on the host cores, bounded sample, N^3-extrapolated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import fvgp_oracle as orc
    use_all_host_threads()
    ns = args.cpu_sample_n
    x, y, noise = synthetic_c2(ns)
    for k in range(min(args.warmup, 1)):
        cpu_baseline_step(x, y, noise, theta_k(k), orc)
    t0 = time.perf_counter()
    for k in range(args.steps):
        cpu_baseline_step(x, y, noise, theta_k(k), orc)
    per = (time.perf_counter() - t0) / args.steps
    scale = (args.n / ns) ** 3
    value = 1.0 / (per * scale)
    cores = os.cpu_count()
    sample = (f"oracle port of the reference algorithm (K-fill numpy, scipy cho_factor, H stacked LU solves) at N={ns}: "
              f"{per:.2f} s per LML+gradient on {cores} host threads; extrapolated x(N/{ns})^3 to N={args.n} "
              f"(the reference gradient needs ~{(3 * 4 + 3) * 8 * args.n ** 2 / 1e9:.0f} GB at N={args.n})")
    line = {"impl": "reference", "metric": "LML+gradient evals/s (dense, N=50k, 3D, ARD Matern-3/2)",
            "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per * scale * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_c4(args):
    """Secondary workload (not the driver's headline line): gp2Scale LML evaluations per second, N = 1M,
    3-D, anisotropic Wendland, K assembled straight to CSR on the device, block-Jacobi PCG + SLQ logdet.
    There is no gradient under gp2Scale in the reference (gp_marginal_likelihood.py:240)."""
    n = args.n
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        from oracle import fvgp_oracle as orc
        import scipy.sparse.linalg as spla
        use_all_host_threads()
        ns = min(n, 20000)
        x, y, noise = synthetic_c4(ns)
        th = theta_c4(0, ns)
        t0 = time.perf_counter()
        K = orc.gp2scale_covariance(x, x, th, batch=2000, symmetric=True)      # the DEFINING dense-block path
        t_fill = time.perf_counter() - t0
        KV = orc.add_kv(K, noise)
        t0 = time.perf_counter()
        sol, iters = orc.sparse_cg(KV, (y - y.mean())[:, None], rtol=1e-5)
        t_cg = time.perf_counter() - t0
        per = t_fill * (n / ns) ** 2 + t_cg * (n / ns)
        line = {"impl": "reference", "metric": "gp2Scale LML evals/s (N=1M, 3D, Wendland)", "value": 1.0 / per,
                "unit": "evals/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": per * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C4: gp2Scale N={n}", "n": n},
                "cpu_baseline": {"value": 1.0 / per, "unit": "evals/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"oracle port at N={ns}: dense-block fill {t_fill:.1f} s (x(N/{ns})^2), scipy cg "
                                           f"{t_cg:.2f} s / {iters[0]} iterations (x N/{ns}); the stochastic logdet (imate) is "
                                           f"not available and not included"},
                "e2e": {"value": 1.0 / per, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    import torch
    from fvgp_b200 import GP, ops, parallel
    from fvgp_b200 import _lib as L
    rank, local_rank, world = parallel.init()
    lib = L.load()
    x, y, noise = synthetic_c4(n)
    mode_args = {"sparse_cg_tol": 1e-5, "random_logdet_lanczos_degree": 20, "random_logdet_min_num_samples": 10,
                 "random_logdet_max_num_samples": 10}
    t0 = time.perf_counter()
    gp = GP(x, y, init_hyperparameters=theta_c4(0, n), noise_variances=noise, gp2Scale=True,
            linalg_mode="sparseCGpre", args=mode_args)
    t_ctor = time.perf_counter() - t0
    for k in range(args.warmup):
        gp.log_likelihood(theta_c4(k + 1, n, rank))
    parallel.barrier()
    torch.cuda.synchronize()
    launches0 = lib.fvgp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        lml = gp.log_likelihood(theta_c4(args.warmup + k + 1, n, rank))
    e1.record()
    torch.cuda.synchronize()
    t_dev = parallel.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    launches = lib.fvgp_launch_count() - launches0
    if args.profile_host and rank == 0:                      # where does the host side of one evaluation go?
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        gp.log_likelihood(theta_c4(args.warmup + args.steps + 2, n, rank))
        torch.cuda.synchronize()
        pr.disable()
        with open(args.profile_host, "w") as fh:
            pstats.Stats(pr, stream=fh).sort_stats("cumulative").print_stats(35)
    # phase breakdown of one evaluation
    xd = gp.data.x_device()
    th = theta_c4(1, n)
    nd = L.to_dev(noise)

    def timed(fn, reps=3):
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e-3)
        return best, out
    t_fill, KV = timed(lambda: ops.wendland_csr(xd, xd, th, noise=nd))
    stats = torch.zeros(1, dtype=torch.int64, device="cuda")
    ops.wendland_csr(xd, xd, th, noise=nd, stats=stats)
    tile_pairs = int(stats.item())
    v = L.to_dev(y - y.mean())
    yv = L.dev_empty((n,))
    t_spmv, _ = timed(lambda: ops.spmv(KV, v, yv), reps=10)
    t_pre, M = timed(lambda: ops.bjacobi(KV))
    t_cg, res = timed(lambda: ops.pcg(KV, v, rtol=1e-5, precond=M), reps=2)
    t_cg0, res0 = timed(lambda: ops.pcg(KV, v, rtol=1e-5), reps=2)
    t_slq, _ = timed(lambda: ops.slq_logdet(KV, degree=20, probes=10, seed=0), reps=1)
    if rank != 0:
        return
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    nnz = KV.nnz
    spmv_gbs = (12.0 * nnz + 16.0 * n) / t_spmv / 1e9
    fill_bytes = 12.0 * nnz + 4.0 * (n + 1) + 8.0 * n * 3
    line = {"metric": "gp2Scale LML evals/s (N=1M, 3D, Wendland, sparse assembly + PCG + SLQ logdet)",
            "value": world * args.steps / t_dev, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C4: gp2Scale, 3-D input, N={n}, wendland_anisotropic, CSR assembly on device, "
                                   f"block-Jacobi PCG rtol 1e-5, SLQ logdet (degree 20, 10 probes)", "n": n, "nnz": nnz,
                       "nnz_per_row": nnz / n, "l2_policy": f"CSR is {12 * nnz / 1e9:.2f} GB > L2"},
            "gpu_launches": int(launches), "constructor_seconds": t_ctor, "last_lml": lml,
            "phases_seconds": {"csr_count+scan+fill": t_fill, "spmv": t_spmv, "bjacobi_build": t_pre,
                               "pcg_bjacobi": t_cg, "pcg_bjacobi_iters": res[2], "pcg_plain": t_cg0,
                               "pcg_plain_iters": res0[2], "slq_10x20": t_slq},
            "roofline": {"bound": "hbm", "kernel": "spmv_kernel (CSR-vector), the inner kernel of PCG and SLQ",
                         "achieved": spmv_gbs, "peak": hbm, "unit": "GB/s", "frac": spmv_gbs / hbm,
                         # dram read + write of one launch on this matrix (ncu --set full,
                         # profiles/r01/ncu_spmv.v6.summary.txt): 1.216 + 0.011 GB
                         "traffic": 1.227e9 if n == 1000000 else None,
                         "algorithmic_bytes": 12.0 * nnz + 16.0 * n},
            "roofline_fill": {"bound": "hbm", "kernel": "wendland_csr_kernel count + fill (output-sensitive bytes)",
                              "achieved": fill_bytes / t_fill / 1e9, "peak": hbm, "unit": "GB/s",
                              "frac": fill_bytes / t_fill / 1e9 / hbm, "algorithmic_bytes": fill_bytes,
                              "tile_pairs_tested": tile_pairs, "pair_tests_per_stored_entry": tile_pairs * 1024.0 / nnz,
                              "note": "output-sensitive bytes; the geometry passes are FP64-issue bound "
                                      "(3 DP instructions per axis and candidate pair), not HBM bound"}}
    print(json.dumps(line), flush=True)


def synthetic_c5(n, seed=5):
    """2-D inputs, smooth signal + noise."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, 2))
    y = np.sin(4 * x[:, 0]) * np.cos(3 * x[:, 1]) + 0.1 * rng.standard_normal(n)
    return x, y, np.full(n, 1e-2)


def run_c5(args):
    """Secondary workload: exact dense GP whose KV is sharded 2-D block-cyclically over all ranks (C5: N = 200 000 =
    320 GB at 8 GPUs; smaller N by default on fewer GPUs so that a step stays within a minute).  One step = one
    log_likelihood (+ gradient with --grad) through the public GP API with args["dense_sharded"] = True; every rank
    calls collectively.  The reference has no multi-GPU dense path and cannot hold this matrix."""
    import torch
    from fvgp_b200 import GP, parallel
    from fvgp_b200 import _lib as L
    rank, local_rank, world = parallel.init()
    lib = L.load()
    n = args.n
    x, y, noise = synthetic_c5(n)
    th0 = np.array([1.0, .2, .2])
    t0 = time.perf_counter()
    gp = GP(x, y, init_hyperparameters=th0, noise_variances=noise, args={"dense_sharded": True})
    t_ctor = time.perf_counter() - t0

    def step(k):
        th = th0 * (1.0 + 0.02 * ((k + 1) % 10))
        lml = gp.log_likelihood(th)
        grad = gp.neg_log_likelihood_gradient(th) if args.grad else None
        return lml, grad
    for k in range(args.warmup):
        step(k)
    parallel.barrier()
    torch.cuda.synchronize()
    launches0 = lib.fvgp_launch_count()
    E = gp.kv._sharded_eval
    E.comm.bytes_received = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        lml, grad = step(args.warmup + k)
    e1.record()
    torch.cuda.synchronize()
    parallel.barrier()
    t_dev = parallel.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    if rank != 0:
        return
    import ctypes
    scratch = L.dev_empty((148 * 8 * 256,))
    pk = ctypes.c_double()
    lib.fvgp_bench_fp64_peak(0, 2, 20000, L.ptr(scratch), ctypes.byref(pk), L.stream_ptr())
    flops = float(n) ** 3 / 3.0 * (3.0 if args.grad else 1.0)
    per_gpu = flops * args.steps / t_dev / world / 1e12
    what = "LML + gradient" if args.grad else "LML"
    line = {"metric": f"dense {what} evals/s, KV 2-D block-cyclic over the GPUs (N={n})", "value": args.steps / t_dev,
            "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong within a run (one matrix over all GPUs); default N grows with the GPU count",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C5-type: exact dense GP, 2-D input, N={n} ({8e-9 * n * n:.0f} GB KV), block-cyclic "
                                   f"Cholesky{' + inverse + gradient traces' if args.grad else ''} over NCCL",
                       "n": n, "grid": list(E.grid), "block": E.nb, "local_GB": E._matrix().local_bytes() / 1e9},
            "gpu_launches": int(lib.fvgp_launch_count() - launches0), "constructor_seconds": t_ctor, "last_lml": lml,
            "last_grad": None if grad is None else [float(g) for g in grad],
            "recv_GB_per_rank_per_step": E.comm.bytes_received / 1e9 / args.steps,
            "roofline": {"bound": "tensor", "kernel": "dgemm_mma_kernel (DMMA.8x8x4), per GPU", "achieved": per_gpu,
                         "peak": pk.value, "unit": "TFLOP/s", "frac": per_gpu / pk.value, "traffic": None,
                         "algorithmic_flops_per_step": flops,
                         "note": "whole step (fill, factor, solves, logdet" + (", inverse, traces" if args.grad else "")
                                 + ") over the algorithmic flops, max over ranks"}}
    print(json.dumps(line), flush=True)


def synthetic_c1(n=1000, seed=1):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 1))
    y = np.sin(5 * x[:, 0]) + np.cos(10 * x[:, 0]) + 0.05 * rng.standard_normal(n)
    return x, y, np.full(n, 1e-2)


def run_c1(args):
    """Secondary workload (BASELINE config 1): what train() does at the reference's own CPU-runnable size
    -- populations of hyperparameter proposals (a differential-evolution generation = pop_size x H = 40 individuals at
    the defaults) on a 1-D, N = 1000 GP.  A step = one population of `--population` LML evaluations through
    GP.log_likelihood_population (host thetas in, host LMLs out).  Reported next to the same proposals evaluated one at
    a time through GP.log_likelihood and to the oracle port on the host cores."""
    import torch
    from fvgp_b200 import GP, ops
    from fvgp_b200 import _lib as L
    lib = L.load()
    n = args.n
    B = args.population
    x, y, noise = synthetic_c1(n)
    h0 = np.array([1.0, 0.3])
    gp = GP(x, y, init_hyperparameters=h0, noise_variances=noise)
    rng = np.random.default_rng(0)

    def thetas(k):
        return h0 * (0.6 + 0.8 * np.random.default_rng(k).random((B, 2)))

    for k in range(args.warmup):
        gp.log_likelihood_population(thetas(k))
    sampler = ClockSampler(0)
    torch.cuda.synchronize()
    sampler.start()
    launches0 = lib.fvgp_launch_count()
    ops.start_phase_timing()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("timed")
    e0.record()
    per_step = []
    for k in range(args.steps):
        t_s = time.perf_counter()
        lml = gp.log_likelihood_population(thetas(args.warmup + k))        # synchronises internally (host LMLs out)
        per_step.append(time.perf_counter() - t_s)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    t_gpu = ops.stop_phase_timing().get("population", 0.0) / args.steps      # first to last kernel of the C-ABI call
    t_pop = e0.elapsed_time(e1) * 1e-3 / args.steps
    launches = (lib.fvgp_launch_count() - launches0) // args.steps
    clocks = sampler.summary()
    # with gradients (multi-start local optimisers, the finite-difference Hessian)
    gp.marginal_likelihood.evaluate_population(thetas(0), with_gradient=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        gp.marginal_likelihood.evaluate_population(thetas(args.warmup + k), with_gradient=True)
    torch.cuda.synchronize()
    t_pop_grad = (time.perf_counter() - t0) / args.steps
    # the same proposals one at a time (what the optimisers did before; also the MCMC path)
    T = thetas(args.warmup)
    for t in T[:3]:
        gp.log_likelihood(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    one = np.array([gp.log_likelihood(t) for t in T])
    torch.cuda.synchronize()
    t_seq = time.perf_counter() - t0
    assert np.array_equal(one, gp.log_likelihood_population(T)), "population and one-at-a-time LML differ"
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import fvgp_oracle as orc
        use_all_host_threads()
        orc.dense_log_likelihood(x, y, T[0], noise)
        t0 = time.perf_counter()
        ref = [orc.dense_log_likelihood(x, y, t, noise) for t in T[:10]]
        per = (time.perf_counter() - t0) / 10
        assert max(abs(a / b - 1) for a, b in zip(one[:10], ref)) <= 1e-8
        cpu = {"value": 1.0 / per, "unit": "evals/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"oracle port, 10 of the {B} proposals at N={n}, {per * 1e3:.1f} ms each on {os.cpu_count()} host "
                         f"threads (parity of the GPU LMLs against these: <= 1e-8 checked in this run)"}
    flops = B * (n ** 3 / 3.0 + 2.0 * n * n)
    line = {"metric": f"LML evals/s (dense, N={n}, 1D, populations of {B} proposals)", "value": B / t_pop,
            "unit": "evals/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_pop * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C1: single-task GP, 1-D input, N={n}, default kernel, dense Cholesky; one step = one "
                                   f"population of {B} LML evaluations (GP.log_likelihood_population)",
                       "n": n, "population": B, "l2_policy": "latency-bound workload: every proposal refills its own K"},
            "clocks": clocks,
            "e2e": {"value": B / t_pop, "unit": "evals/s", "h2d_bytes_per_step": B * 4 * 8 + 2 * n * 8,
                    "d2h_bytes_per_step": B * (n + 2) * 8 + B * 4},
            "gpu_launches": int(launches),
            "device_ms_per_step": t_gpu * 1e3, "ms_per_step_median": float(np.median(per_step)) * 1e3,
            "ms_per_step_max": float(np.max(per_step)) * 1e3,
            "one_at_a_time": {"value": B / t_seq, "unit": "evals/s", "ms_per_eval": t_seq / B * 1e3},
            "population_with_gradient": {"value": B / t_pop_grad, "unit": "evals/s"},
            "roofline": {"bound": "tensor", "kernel": "whole population (latency-bound chains of small launches)",
                         "achieved": flops / t_pop / 1e12, "peak": 37.1, "unit": "TFLOP/s",
                         "frac": flops / t_pop / 1e12 / 37.1, "traffic": None,
                         "note": "N^3/3 + 2N^2 flop per proposal; at this size the bound is launch latency and the "
                                 "serial 128-column tile factorisations, not the tensor pipe"},
            "cpu_baseline": cpu, "last_lml": float(lml[-1])}
    print(json.dumps(line), flush=True)


def gemm_dram_traffic(n):
    """(bytes, note): what the DMMA GEMM launches of one N = 50 000 evaluation moved through DRAM, from the committed
    ncu launch list (tools/launch_summary.py JSON); (None, why) when no capture exists for this size."""
    import glob
    if n != 50000:
        return None, "no ncu capture at this size"
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*", "launches_bench_n50k*.json")), reverse=True):
        try:
            with open(path) as fh:
                d = json.load(fh)
            gemm = {k: v for k, v in d["kernels"].items() if "dgemm_mma_kernel" in k}
            tot = sum(v["dram_read"] + v["dram_write"] for v in gemm.values())
            if tot > 0:
                note = (f"dram__bytes_read.sum + dram__bytes_write.sum over {sum(v['launches'] for v in gemm.values())} "
                        f"dgemm_mma_kernel launches ({sum(v['ms'] for v in gemm.values()):.0f} ms of kernel time) of one "
                        f"evaluation, {os.path.relpath(path, ROOT)}")
                if "partial" in os.path.basename(path):
                    note += ("; PARTIAL: the capture covers POTRF and TRTRI completely and 106 of the 453 LAUUM launches "
                             "(about 2/3 of the N^3 flop), ncu ran into its time limit")
                return float(tot), note
        except Exception:
            continue
    return None, "no ncu capture committed"


def workload_config(args):
    return {"workload": f"C2: single-task GP, 3-D input, N={args.n}, anisotropic Matern-3/2 (default kernel), dense FP64 "
                        f"Cholesky, LML + hyperparameter gradient per step",
            "n": args.n, "dim": 3, "hyperparameters": 4, "parallelism": f"replicas x{args.gpus} (one theta proposal per GPU)",
            "l2_policy": f"inputs larger than L2: K is {8 * args.n ** 2 / 1e9:.1f} GB"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--size", "--n", dest="n", type=int, default=50000,
                    help="problem size N (use --size under torchrun: its own parser rejects a bare --n as ambiguous)")
    ap.add_argument("--cpu-sample-n", type=int, default=3000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-host", default="", help="c4 only: write a cProfile of one evaluation to this file")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c4", "c5"],
                    help="c2 (default, the headline): dense N=50k LML+gradient; c4: gp2Scale N=1M LML; "
                         "c5: dense LML with KV block-cyclic over all GPUs; c1: populations of proposals at N=1000")
    ap.add_argument("--population", type=int, default=40, help="c1 only: proposals per step")
    ap.add_argument("--grad", action="store_true", help="c5 only: add the gradient to every step")
    args = ap.parse_args()
    if args.workload == "c1":
        if args.n == 50000:
            args.n = 1000
        if args.steps == 3:
            args.steps = 20            # a step is ~4 ms: more of them, so that one host hiccup does not decide the line
        return run_c1(args)
    if args.workload == "c4":
        if args.n == 50000:
            args.n = 1000000
        return run_c4(args)
    if args.workload == "c5":
        if args.n == 50000:
            world = int(os.environ.get("WORLD_SIZE", "1"))
            args.n = {1: 60000, 2: 100000, 4: 140000}.get(world, 200000)
        return run_c5(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from fvgp_b200 import GP, ops, parallel
    from fvgp_b200 import _lib as L
    rank, local_rank, world = parallel.init()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    lib = L.load()
    n = args.n
    x, y, noise = synthetic_c2(n)
    x_pinned = torch.from_numpy(x).pin_memory()
    gp = GP(x, y, init_hyperparameters=theta_k(0), noise_variances=noise)
    H = 4

    def step(k):
        th = theta_k(k, rank)
        return gp.log_likelihood(th), gp.neg_log_likelihood_gradient(th)

    # ---- device-resident timing ---------------------------------------------------------------
    for k in range(args.warmup):
        step(k)
    sampler = ClockSampler(local_rank)
    parallel.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = lib.fvgp_launch_count()
    ops.start_phase_timing()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("timed")          # ncu --nvtx --nvtx-include "timed/" lists exactly these launches
    e0.record()
    for k in range(args.steps):
        out = step(args.warmup + k)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    parallel.barrier()
    phases = ops.stop_phase_timing()
    launches = lib.fvgp_launch_count() - launches0
    clocks = sampler.summary()
    t_dev = parallel.max_over_ranks(e0.elapsed_time(e1) * 1e-3)

    # ---- end to end: host buffers in, host results out ---------------------------------------
    parallel.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        gp.data._x_dev = x_pinned.cuda(non_blocking=True)          # H2D of this step's inputs
        lml, grad = step(args.warmup + args.steps + k)
        assert np.isfinite(lml) and np.all(np.isfinite(grad))
    torch.cuda.synchronize()
    t_e2e = parallel.max_over_ranks(time.perf_counter() - t0)
    h2d = x.nbytes + 2 * n * 8                                      # x, y - m, noise diagonal
    d2h = n * 8 + 8 + 4 + H * 8                                     # KVinvY, logdet, potrf status, traces

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (DMMA GEMM inside potrf + potri) ----------------------
    import ctypes
    scratch = L.dev_empty((148 * 8 * 256,))
    pk = ctypes.c_double()
    lib.fvgp_bench_fp64_peak(0, 2, 20000, L.ptr(scratch), ctypes.byref(pk), L.stream_ptr())
    t_tensor = (phases.get("potrf", 0.0) + phases.get("potri", 0.0)) / args.steps
    achieved = n ** 3 / t_tensor / 1e12
    roofline = {"bound": "tensor", "kernel": "dgemm_mma_kernel (DMMA.8x8x4) inside potrf + potri", "achieved": achieved,
                "peak": pk.value, "unit": "TFLOP/s", "frac": achieved / pk.value,
                # dram__bytes_read.sum + dram__bytes_write.sum summed over every dgemm_mma_kernel launch of ONE timed
                # evaluation (ncu launch list of this command, profiles/r01/launches_bench_n50k.*.json)
                "traffic": gemm_dram_traffic(n)[0], "traffic_note": gemm_dram_traffic(n)[1],
                "peak_source": "measured live: register-resident DMMA.8x8x4 issue rate (fvgp_bench_fp64_peak)",
                "algorithmic_flops_per_step": float(n) ** 3,
                "phase_seconds_per_step": {k: v / args.steps for k, v in phases.items()}}
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    out_buf = L.dev_matrix(n, n)
    xd, nd = gp.data.x_device(), L.to_dev(noise)
    th = theta_k(1)
    best = 1e30
    bounds = (x.min(axis=0), x.max(axis=0))      # host reductions stay outside the event pair
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.kfill(L.K_MATERN32, xd, xd, th[0], 1 / th[1:], 1.0, noise=nd, mode=L.FILL_SYMMETRIC, out=out_buf, bounds=bounds)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    kfill_gbs = 8.0 * n * n / best / 1e9
    roofline_kfill = {"bound": "hbm", "kernel": "kfill_kernel<MATERN32,3> symmetric (full square + noise diagonal)",
                      "achieved": kfill_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": kfill_gbs / hbm_peak,
                      # dram__bytes_read.sum + dram__bytes_write.sum of one launch at N = 50 000
                      # (ncu --set full, profiles/r01/ncu_kfill.v6.summary.txt): 0.081 + 19.941 GB
                      "traffic": 20.022e9 if n == 50000 else None, "peak_source": hbm_src,
                      "algorithmic_bytes": 8.0 * n * n}
    del out_buf

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import fvgp_oracle as orc
        use_all_host_threads()
        ns = args.cpu_sample_n
        xs, ys, vs = synthetic_c2(ns)
        t0 = time.perf_counter()
        cpu_baseline_step(xs, ys, vs, theta_k(1), orc)
        per = time.perf_counter() - t0
        cpu = {"value": 1.0 / (per * (n / ns) ** 3), "unit": "evals/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"oracle port of the reference algorithm, one LML+gradient at N={ns} took {per:.2f} s on "
                         f"{os.cpu_count()} host threads, extrapolated x(N/{ns})^3 to N={n}"}

    line = {"metric": "LML+gradient evals/s (dense, N=50k, 3D, ARD Matern-3/2)",
            "value": world * args.steps / t_dev, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args), "clocks": clocks,
            "e2e": {"value": world * args.steps / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_kfill": roofline_kfill,
            "cpu_baseline": cpu, "last_lml": out[0]}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
