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
  e2e          same metric through the public API with HOST buffers: args["host_inputs_every_call"]
               makes every call copy x from pinned host memory; y - m and the noise diagonal are
               uploaded and LML / gradient read back by every call anyway (copies inside the timed region).
  roofline     dominant kernel = the DMMA GEMM behind POTRF + POTRI: algorithmic N^3 flop / (CUDA-event
               time of the potrf + potri phases); peak = DMMA issue rate measured live in this run
               (MEASURED_PEAKS.json has no FP64 figure).  roofline_kfill: the K-assembly kernel against
               HBM, in the symmetric mode (full square written) and in the lower mode the LML path uses.
  cpu_baseline the numpy/scipy oracle port of the reference algorithm on the host cores, bounded sample.
  parity       (N = 1 only, after the timed regions) the GPU results at the BENCHMARKED sizes against the
               oracle on the host: C2 LML at N = 50 000, gradient at N = 8 000 / 16 000, C4 sampled blocks
               of the N = 1M pattern, gp2Scale sparseLU LML at N = 50 000, C3-shaped fvGP through the
               sharded evaluator.  pass / fail per check, tolerances in the record.
  c4 / c1      the other single-GPU workloads of BASELINE.json's metric, as extra keys of the same line.
  sharded      (N > 1) the DATA-sharded paths measured in the same run: C3-size dense LML + gradient with KV
               2-D block-cyclic over all ranks (+ agreement with one GPU at the same size), C5 at 8 GPUs,
               gp2Scale with the CSR rows sharded over the ranks.
N > 1 headline: hyperparameter proposals are independent evaluations (MCMC / DE populations), so each
rank evaluates its own theta sequence on a replica of the data -- no data-path collective ("weak").

--impl reference: the UNMODIFIED reference package (baseline/_ref, see baseline/install_ref.py) through
its own GP.log_likelihood / neg_log_likelihood_gradient on the host cores; each step a bounded sample
(N = --cpu-sample-n), plus one measurement per size of a ladder for the a N^2 + b N^3 fit that extrapolates
to N = 50 000 (the reference's gradient needs ~300 GB there) and a SAME-N measured point (N = 8 000) that
the GPU arm measures too (`same_n`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

T_START = time.time()
import faulthandler

faulthandler.enable()                 # a native crash must leave a Python traceback on stderr


def log(msg):
    """Progress marker on stderr (stdout carries the one JSON line only)."""
    print(f"[bench {time.time() - T_START:7.1f}s] {msg}", file=sys.stderr, flush=True)


if "reference" in sys.argv and "TORCHELASTIC_RUN_ID" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
    # torchrun exports OMP_NUM_THREADS=1 when it starts more than one rank; the CPU arm (rank 0 only) is meant
    # to use every host core, and OpenBLAS sizes its thread pool when numpy is first imported
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HOLD = {"line": None, "printed": False}      # what the watchdog prints if an optional section hangs
SAME_N = 8000                      # size both arms measure directly (no extrapolation in that ratio)
METRIC_C2 = "LML+gradient evals/s (dense, N=50k, 3D, ARD Matern-3/2)"


# ------------------------------------------------------------------------------------------ synthetic data
def synthetic_c2(n, seed=2):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 3))
    y = np.sin(5 * x[:, 0]) * np.cos(3 * x[:, 1]) + x[:, 2] + 0.1 * rng.standard_normal(n)
    return x, y, np.full(n, 1e-2)


def synthetic_c3(points=20000, tasks=5, seed=3):
    """2-D inputs x 5 tasks -> (points * tasks)-row K on the 3-D index set."""
    rng = np.random.default_rng(seed)
    x = rng.random((points, 2))
    y = np.stack([np.sin((3 + t) * x[:, 0]) + np.cos(2 * x[:, 1]) + 0.1 * rng.standard_normal(points)
                  for t in range(tasks)], axis=1)
    return x, y, np.full((points, tasks), 1e-2)


THETA_C3 = np.array([1.0, .3, .3, 2.0])


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


def synthetic_c5(n, seed=5):
    """2-D inputs, smooth signal + noise."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, 2))
    y = np.sin(4 * x[:, 0]) * np.cos(3 * x[:, 1]) + 0.1 * rng.standard_normal(n)
    return x, y, np.full(n, 1e-2)


def synthetic_c1(n=1000, seed=1):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 1))
    y = np.sin(5 * x[:, 0]) + np.cos(10 * x[:, 0]) + 0.05 * rng.standard_normal(n)
    return x, y, np.full(n, 1e-2)


# ------------------------------------------------------------------------------------------ helpers
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


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def event_timed(fn, reps=1):
    """Best-of-reps CUDA-event time of fn() on torch's current stream (seconds), and its last result."""
    import torch
    best, out = 1e30, None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best, out


class Deadline:
    """Sub-records are optional evidence: each one runs only while the run is inside its time budget, and the
    decision is taken on rank 0 and broadcast so that collective sections are entered by all ranks or none."""

    def __init__(self, seconds):
        self.limit = float(seconds)

    def left(self):
        return self.limit - (time.time() - T_START)

    def allows(self, need):
        import torch
        import torch.distributed as dist
        ok = self.left() > need
        if dist.is_available() and dist.is_initialized():
            t = torch.tensor([1 if ok else 0], device="cuda")
            dist.broadcast(t, 0)
            ok = bool(t.item())
        return ok


def emit(line, rank):
    if rank == 0 and not HOLD["printed"]:
        HOLD["printed"] = True
        print(json.dumps(line), flush=True)


def start_watchdog(limit_s, rank):
    """The headline is measured first; the optional sections after it (multi-rank collectives, host oracles) must
    never be able to lose it.  Past the hard limit every rank leaves, rank 0 after printing what it has."""
    def run():
        while time.time() - T_START < limit_s:
            time.sleep(1.0)
        line = HOLD["line"]
        if line is not None and not HOLD["printed"]:
            line["watchdog"] = f"optional sections cut off after {limit_s:.0f} s"
            emit(line, rank)
        sys.stdout.flush()
        os._exit(0 if line is not None else 3)
    threading.Thread(target=run, daemon=True).start()


def guarded(name, fn, sink):
    """Run one optional section; a failure is recorded in the line instead of killing the headline."""
    t0 = time.time()
    log(f"section {name} ...")
    try:
        sink[name] = fn()
    except Exception as e:                      # noqa: BLE001 -- evidence sections must never lose the headline
        import traceback
        sink[name] = {"error": f"{type(e).__name__}: {e}", "trace_tail": traceback.format_exc()[-600:]}
    if isinstance(sink.get(name), dict):
        sink[name]["section_seconds"] = round(time.time() - t0, 2)


# ------------------------------------------------------------------------------------------ reference arm
def cpu_port_step(x, y, noise, theta, orc):
    """Reference ALGORITHM (stacked LU solves of KV against dK/dtheta) from the oracle port."""
    lml = orc.dense_log_likelihood(x, y, theta, noise)
    grad = orc.dense_neg_log_likelihood_gradient(x, y, theta, noise, economical=False)
    return lml, grad


def _reference_gp(n):
    """The unmodified reference's GP on the C2 data of size n (None when baseline/_ref is absent)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import install_ref
    fv = install_ref.import_reference()
    x, y, noise = synthetic_c2(n)
    import warnings
    warnings.filterwarnings("ignore")
    return fv.GP(x, y, init_hyperparameters=theta_k(0), noise_variances=noise)


def fit_n2_n3(ns, ts):
    """Non-negative least squares of t = a N^2 + b N^3 (relative residuals)."""
    from scipy.optimize import nnls
    ns, ts = np.asarray(ns, dtype=float), np.asarray(ts, dtype=float)
    A = np.stack([ns ** 2, ns ** 3], axis=1) / ts[:, None]
    coef, _ = nnls(A, np.ones(len(ts)))
    return float(coef[0]), float(coef[1])


def run_reference(args):
    """--impl reference: see the module docstring."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = use_all_host_threads()
    budget = float(os.environ.get("FVGP_REF_BUDGET_S", "330"))
    ns = args.cpu_sample_n
    kind = "reference"
    try:
        gp = _reference_gp(ns)

        def step(g, k):
            th = theta_k(k)
            return g.log_likelihood(th), g.neg_log_likelihood_gradient(th)
    except Exception as e:                                         # baseline/_ref missing: the oracle port
        kind = "port"
        why = f"{type(e).__name__}: {e}"
        from oracle import fvgp_oracle as orc

        class _Port:
            def __init__(self, n):
                self.x, self.y, self.noise = synthetic_c2(n)
        gp = _Port(ns)

        def step(g, k):
            return cpu_port_step(g.x, g.y, g.noise, theta_k(k), orc)
    for k in range(min(args.warmup, 2)):
        step(gp, k)
    t0 = time.perf_counter()
    for k in range(args.steps):
        out = step(gp, k)
    per = (time.perf_counter() - t0) / args.steps
    measured = {ns: per}
    # ladder for the fit: one evaluation per size while the budget lasts (cost grows ~8x per doubling)
    make = _reference_gp if kind == "reference" else (lambda n: _Port(n))
    for n2 in (1000, 4000, SAME_N, 16000):
        if n2 in measured:
            continue
        ref_n = max(measured)
        predict = measured[ref_n] * (n2 / ref_n) ** 3 * 1.3 + 2.0
        if n2 > ref_n and (time.time() - T_START) + predict > budget:
            continue
        g2 = make(n2)
        t0 = time.perf_counter()
        step(g2, 1)
        measured[n2] = time.perf_counter() - t0
        del g2
    sizes = sorted(measured)
    a, b = fit_n2_n3(sizes, [measured[s] for s in sizes])
    t_full = a * args.n ** 2 + b * args.n ** 3
    value = 1.0 / t_full
    sample = (f"{'UNMODIFIED reference (baseline/_ref) GP.log_likelihood + GP.neg_log_likelihood_gradient' if kind == 'reference' else 'oracle port (baseline/_ref missing: ' + why + ')'}"
              f" on {cores} host threads; timed steps at N={ns}: {per:.2f} s each; one evaluation per size "
              f"{ {s: round(measured[s], 2) for s in sizes} } s; fit t = a N^2 + b N^3 (a={a:.3e}, b={b:.3e}) extrapolated to "
              f"N={args.n}: {t_full:.0f} s (the reference gradient needs ~{(3 * 4 + 3) * 8 * args.n ** 2 / 1e9:.0f} GB "
              f"there and cannot run); `value` is that extrapolation, `ms_per_step` the measured step at N={ns}")
    line = {"impl": "reference", "metric": METRIC_C2,
            "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "extrapolated": True, "measured_seconds_by_n": {str(s): measured[s] for s in sizes},
            "fit": {"model": "t = a N^2 + b N^3", "a": a, "b": b, "extrapolated_seconds_at_n": t_full},
            "same_n": {"n": SAME_N, "seconds": measured.get(SAME_N), "evals_per_s": (1.0 / measured[SAME_N]) if SAME_N in measured else None,
                       "note": "measured, not extrapolated; the GPU arm reports the same size under `same_n`"},
            "last_lml": float(out[0]),
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.with_c4 and (time.time() - T_START) < budget:
        try:
            line["c4"] = reference_c4(min(args.c4_cpu_n, 200000), 1000000)
        except Exception as e:                                     # noqa: BLE001
            line["c4"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)


def reference_c4(ns, n_full):
    """gp2Scale on the host, the reference's OWN fast path, called through the
    reference's own entry points on a bounded sample of the C4 point set (N = ns, support radius scaled to keep
    ~102 nnz/row):
      fill   gp2Scale_covariance.distributed_covariance (gp2Scale_covariance.py:313-431; blockwise, B = 10 000, synchronous
             stub client -- dask is not installable here) with kernels.wendland_anisotropic_gp2Scale_cpu_sparse
             (kernels.py:724, KD-tree ball queries behind the AABB cull) + assemble_triplets
      solve  gp_lin_alg.calculate_sparse_conj_grad (gp_lin_alg.py:1213-1291; scipy cg, no preconditioner: the default ILU
             is impractical)
    The stochastic log-determinant (imate) is not installed and NOT included, which favours the reference.  The
    defining dense-block kernel (kernels.py:502-528) is timed on two 10k x 10k blocks as a labelled second number."""
    import warnings
    warnings.filterwarnings("ignore")
    x, y, noise = synthetic_c4(ns)
    th = theta_c4(0, ns)
    rhs = (y - y.mean())[:, None]
    out = {"n_sample": ns, "cores": os.cpu_count()}
    from oracle import fvgp_oracle as orc
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import install_ref
        install_ref.import_reference()
        import ref_stubs
        from fvgp import gp2Scale_covariance as g2s
        from fvgp import gp_lin_alg as rla
        from fvgp import kernels as rk
        client = ref_stubs.Client()
        fut = client.scatter(x)
        t0 = time.perf_counter()
        K = g2s.distributed_covariance(client, rk.wendland_anisotropic_gp2Scale_cpu_sparse, th, fut, ns, fut, ns, 10000,
                                       symmetric=True, distribution="blockwise", k_n_params=3, args={})
        t_fill = time.perf_counter() - t0
        KV = orc.add_kv(K.tocsr(), noise)
        t0 = time.perf_counter()
        sol = rla.calculate_sparse_conj_grad(KV, rhs, args={"sparse_cg_tol": 1e-5})
        t_cg = time.perf_counter() - t0
        out["kind"], iters = "reference", None
    except Exception as e:                                         # noqa: BLE001 -- baseline/_ref missing: oracle port
        out["kind"] = "port"
        out["reference_error"] = f"{type(e).__name__}: {e}"[:300]
        ns = min(ns, 20000)
        x, y, noise = synthetic_c4(ns)
        th = theta_c4(0, ns)
        t0 = time.perf_counter()
        K = orc.gp2scale_covariance(x, x, th, batch=2000, symmetric=True)
        t_fill = time.perf_counter() - t0
        KV = orc.add_kv(K, noise)
        t0 = time.perf_counter()
        sol, it = orc.sparse_cg(KV, (y - y.mean())[:, None], rtol=1e-5)
        t_cg = time.perf_counter() - t0
        iters = int(it[0])
        out["n_sample"] = ns
    t0 = time.perf_counter()
    nb = 0
    for i0 in range(0, min(ns, 20000), 10000):
        orc.wendland_block(x[i0:i0 + 10000], x[i0:i0 + 10000], th)
        nb += 1
    t_dense_block = (time.perf_counter() - t0) / max(nb, 1)
    scale = n_full / ns
    per = (t_fill + t_cg) * scale                                  # both phases are O(N) at fixed nnz/row
    out.update({"nnz": int(K.nnz), "fill_seconds": t_fill, "cg_seconds": t_cg, "cg_iterations": iters,
                "value": 1.0 / per, "unit": "evals/s", "extrapolated": True,
                "sample": f"KD-tree block kernel + assemble_triplets at N={ns}: {t_fill:.1f} s; scipy cg (no preconditioner, rtol 1e-5): "
                          f"{t_cg:.1f} s; scaled x{scale:.0f} (linear in N at fixed nnz/row) to N={n_full}; log-determinant not "
                          f"included (imate absent) -- favours the reference",
                "dense_block_seconds_per_10k_block": t_dense_block,
                "dense_block_path_seconds_at_full_n": t_dense_block * (n_full / 10000) * (n_full / 10000 + 1) / 2})
    return out


# ------------------------------------------------------------------------------------------ C4 (gp2Scale)
C4_ARGS = {"sparse_cg_tol": 1e-5, "random_logdet_lanczos_degree": 20, "random_logdet_min_num_samples": 10,
           "random_logdet_max_num_samples": 10}


def c4_record(n, steps, warmup, rank=0, world=1, phases=True, sharded=False):
    """gp2Scale LML evaluations per second: K assembled straight to CSR on the device, block-Jacobi PCG + SLQ logdet.
    There is no gradient under gp2Scale in the reference (gp_marginal_likelihood.py:240).  sharded=True: the CSR rows
    and the SLQ probes are split over the ranks (fvgp_b200/sharded_sparse.py), one LML for the whole job per step;
    otherwise every rank evaluates its own theta (replicas)."""
    import torch
    from fvgp_b200 import GP, ops, parallel
    from fvgp_b200 import _lib as L
    lib = L.load()
    x, y, noise = synthetic_c4(n)
    mode_args = dict(C4_ARGS)
    if sharded:
        mode_args["gp2Scale_sharded"] = True
    t0 = time.perf_counter()
    gp = GP(x, y, init_hyperparameters=theta_c4(0, n), noise_variances=noise, gp2Scale=True,
            linalg_mode="sparseCGpre", args=mode_args)
    t_ctor = time.perf_counter() - t0
    trank = 0 if sharded else rank
    for k in range(warmup):
        gp.log_likelihood(theta_c4(k + 1, n, trank))
    parallel.barrier()
    torch.cuda.synchronize()
    launches0 = lib.fvgp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    variances = []
    for k in range(steps):
        lml = gp.log_likelihood(theta_c4(warmup + k + 1, n, trank))
        variances.append(gp.log_likelihood_variance())
    e1.record()
    torch.cuda.synchronize()
    t_dev = parallel.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    launches = lib.fvgp_launch_count() - launches0
    evals = steps if sharded else world * steps
    rec = {"metric": "gp2Scale LML evals/s (N=1M, 3D, Wendland, sparse assembly + PCG + SLQ logdet)",
           "value": evals / t_dev, "unit": "evals/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": t_dev / steps * 1e3, "scaling": "strong (rows and probes sharded)" if sharded else "weak (replicas)",
           "config": {"workload": f"C4: gp2Scale, 3-D input, N={n}, wendland_anisotropic, CSR assembly on device, "
                                  f"block-Jacobi PCG rtol 1e-5, SLQ logdet (degree 20, 10 probes)", "n": n,
                      "parallelism": ("CSR row slabs + SLQ probes over the ranks" if sharded else f"replicas x{world}")},
           "gpu_launches": int(launches), "constructor_seconds": t_ctor, "last_lml": lml,
           "log_likelihood_variance": variances[-1],
           "log_likelihood_std": None if variances[-1] is None else float(np.sqrt(variances[-1])),
           "cg": {k2: v for k2, v in gp.kv.state.info.items() if k2.startswith("cg_")} if gp.kv.state is not None else None}
    if sharded:
        rec["sharding"] = getattr(gp.kv, "last_sharded_sparse_info", None)
        ev_s = getattr(gp.kv, "_sparse_eval", None)
        if ev_s is not None and ev_s.timing is not None:          # FVGP_SHARDED_TIMING=1 (diagnosis: marks synchronise)
            calls = warmup + steps + 1
            rec["phase_ms_per_evaluation"] = {k: 1e3 * v / calls for k, v in ev_s.timing.items()}
    if not phases:
        del gp
        return rec
    # phase breakdown of one evaluation (single-GPU kernels)
    xd = gp.data.x_device()
    th = theta_c4(1, n)
    nd = L.to_dev(noise)
    t_fill, KV = event_timed(lambda: ops.wendland_csr(xd, xd, th, noise=nd), reps=3)
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    ops.wendland_csr(xd, xd, th, noise=nd, stats=stats)
    tile_pairs = int(stats[0].item())
    v = L.to_dev(y - y.mean())
    yv = L.dev_empty((n,))
    t_spmv, _ = event_timed(lambda: ops.spmv(KV, v, yv), reps=10)
    t_pre, M = event_timed(lambda: ops.bjacobi(KV), reps=3)
    t_cg, res = event_timed(lambda: ops.pcg(KV, v, rtol=1e-5, precond=M), reps=2)
    t_cg0, res0 = event_timed(lambda: ops.pcg(KV, v, rtol=1e-5), reps=2)
    t_slq, _ = event_timed(lambda: ops.slq_logdet(KV, degree=20, probes=10, seed=0), reps=1)
    hbm = load_peaks().get("hbm_gbs", 6650.0)
    nnz = KV.nnz
    spmv_bytes = 12.0 * nnz + 16.0 * n
    fill_bytes = 12.0 * nnz + 4.0 * (n + 1) + 8.0 * n * 3
    pcg_iter_bytes = 12.0 * nnz + 5 * 8.0 * n
    rec["config"].update({"nnz": nnz, "nnz_per_row": nnz / n, "l2_policy": f"CSR is {12 * nnz / 1e9:.2f} GB > L2"})
    rec["phases_seconds"] = {"csr_count+scan+fill": t_fill, "spmv": t_spmv, "bjacobi_build": t_pre,
                             "pcg_bjacobi": t_cg, "pcg_bjacobi_iters": res[2], "pcg_plain": t_cg0,
                             "pcg_plain_iters": res0[2], "slq_10x20": t_slq,
                             "pcg_iteration_over_spmv": (t_cg / max(res[2], 1)) / t_spmv}
    rec["roofline"] = {"bound": "hbm", "kernel": "spmv_kernel (CSR-vector), the inner kernel of PCG and SLQ",
                       "achieved": spmv_bytes / t_spmv / 1e9, "peak": hbm, "unit": "GB/s",
                       "frac": spmv_bytes / t_spmv / 1e9 / hbm,
                       # dram read + write of one launch on this matrix (ncu --set full,
                       # profiles/r01/ncu_spmv.v6.summary.txt): 1.216 + 0.011 GB
                       "traffic": 1.227e9 if n == 1000000 else None, "algorithmic_bytes": spmv_bytes}
    rec["roofline_pcg"] = {"bound": "hbm", "kernel": "one PCG iteration (SpMV + vector updates + block-Jacobi apply)",
                           "achieved": pcg_iter_bytes * res[2] / t_cg / 1e9, "peak": hbm, "unit": "GB/s",
                           "frac": pcg_iter_bytes * res[2] / t_cg / 1e9 / hbm,
                           "algorithmic_bytes_per_iteration": pcg_iter_bytes}
    rec["roofline_fill"] = {"bound": "hbm", "kernel": "wendland_csr_kernel count + fill (output-sensitive bytes)",
                            "achieved": fill_bytes / t_fill / 1e9, "peak": hbm, "unit": "GB/s",
                            "frac": fill_bytes / t_fill / 1e9 / hbm, "algorithmic_bytes": fill_bytes,
                            "tile_pairs_tested": tile_pairs,
                            "pair_tests_per_stored_entry": float(stats[1].item()) / nnz if int(stats[1].item()) else tile_pairs * 1024.0 / nnz,
                            "note": "output-sensitive bytes; the geometry passes are FP64-issue bound, not HBM bound"}
    del gp, KV
    return rec


def run_c4(args):
    """--workload c4: the gp2Scale line on its own."""
    n = args.n
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        use_all_host_threads()
        rec = reference_c4(min(args.c4_cpu_n, n), n)
        line = {"impl": "reference", "metric": "gp2Scale LML evals/s (N=1M, 3D, Wendland)", "value": rec["value"],
                "unit": "evals/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": 1e3 / rec["value"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C4: gp2Scale N={n}", "n": n},
                "cpu_baseline": {"value": rec["value"], "unit": "evals/s", "cores": os.cpu_count(), "kind": rec["kind"],
                                 "sample": rec["sample"]},
                "detail": rec,
                "e2e": {"value": rec["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    from fvgp_b200 import parallel
    rank, local_rank, world = parallel.init()
    rec = c4_record(n, args.steps, args.warmup, rank, world, phases=True, sharded=args.sharded and world > 1)
    if rank != 0:
        return
    rec.update({"higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic"})
    print(json.dumps(rec), flush=True)


# ------------------------------------------------------------------------------------------ C5 / sharded dense
def sharded_dense_record(gp_factory, n, steps, warmup, with_grad, label):
    """LML (+ gradient) with KV 2-D block-cyclic over ALL ranks through the public API (args["dense_sharded"]).
    Every rank calls collectively with the same theta.  Returns the record (meaningful on rank 0)."""
    import ctypes
    import torch
    from fvgp_b200 import parallel
    from fvgp_b200 import _lib as L
    lib = L.load()
    rank, _, world = parallel.dist_env()
    t0 = time.perf_counter()
    gp, th0 = gp_factory()
    t_ctor = time.perf_counter() - t0

    def step(k):
        th = th0 * (1.0 + 0.02 * ((k + 1) % 10))
        lml = gp.log_likelihood(th)
        grad = gp.neg_log_likelihood_gradient(th) if with_grad else None
        return th, lml, grad
    for k in range(warmup):
        step(k)
    parallel.barrier()
    torch.cuda.synchronize()
    launches0 = lib.fvgp_launch_count()
    E = gp.kv._sharded_eval
    E.comm.bytes_received = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        th, lml, grad = step(warmup + k)
    e1.record()
    torch.cuda.synchronize()
    parallel.barrier()
    t_dev = parallel.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    scratch = L.dev_empty((148 * 8 * 256,))
    pk = ctypes.c_double()
    lib.fvgp_bench_fp64_peak(0, 2, 20000, L.ptr(scratch), ctypes.byref(pk), L.stream_ptr())
    flops = float(n) ** 3 / 3.0 * (3.0 if with_grad else 1.0)
    per_gpu = flops * steps / t_dev / world / 1e12
    rec = {"what": label, "n": n, "n_gpus": world, "grid": list(E.grid), "block": E.nb, "steps": steps, "warmup": warmup,
           "seconds_per_step": t_dev / steps, "evals_per_s": steps / t_dev, "with_gradient": bool(with_grad),
           "tflops_per_gpu": per_gpu, "frac_of_fp64_tensor_peak": per_gpu / pk.value, "peak_tflops": pk.value,
           "algorithmic_flops_per_step": flops, "local_GB": E._matrix().local_bytes() / 1e9,
           "recv_GB_per_rank_per_step": E.comm.bytes_received / 1e9 / steps,
           "gpu_launches_per_step": int(lib.fvgp_launch_count() - launches0) // steps,
           "constructor_seconds": t_ctor, "theta": [float(t) for t in th], "lml": lml,
           "grad": None if grad is None else [float(g) for g in grad],
           "phase_seconds_last_step": E.phase_seconds() if hasattr(E, "phase_seconds") else None}
    return rec, gp


def release(gp):
    """Drop the device buffers a GP holds (memoised evaluation, sharded matrix) before the next section."""
    import gc
    import torch
    try:
        gp.kv._memo = None
        gp.kv.state = None
        gp.kv._sharded_eval = None
    except Exception:
        pass
    del gp
    gc.collect()
    torch.cuda.empty_cache()


def single_gpu_check(make_gp, th, with_grad=True):
    """The same LML (+ gradient) on ONE GPU (rank 0 only): agreement flag and the strong-scaling baseline."""
    import torch
    gp = make_gp()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lml = gp.log_likelihood(th)
    grad = gp.neg_log_likelihood_gradient(th) if with_grad else None
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    release(gp)
    return lml, grad, dt


def sharded_section(args, deadline, rank, world):
    """N > 1: the data-sharded paths, measured in the same run as the replica headline."""
    from fvgp_b200 import GP, fvGP, parallel
    out = {}

    # ---- C3: fvGP 20 000 points x 5 tasks -> N = 100 000 rows, default kernel on the 3-D index set, LML + gradient
    def c3():
        pts = args.c3_points
        x, y, noise = synthetic_c3(pts)
        n = pts * y.shape[1]

        def factory():
            return fvGP(x, y, init_hyperparameters=THETA_C3, noise_variances=noise, args={"dense_sharded": True}), THETA_C3
        rec, gp = sharded_dense_record(factory, n, steps=max(3, min(args.steps, 3)), warmup=1, with_grad=True,
                                       label=f"C3: fvGP {pts} points x {y.shape[1]} tasks, dense LML + gradient, KV block-cyclic")
        th = np.array(rec["theta"])
        release(gp)
        parallel.barrier()
        if 10.0 * n * n < 0.85 * 180e9 and deadline.allows(150):           # fits one GPU: agreement + strong-scaling baseline
            if rank == 0:
                try:                                                # a failure here must not keep rank 0 from the barrier
                    from fvgp_b200 import _lib as L_
                    lib_ = L_.load()

                    def one():
                        return fvGP(x, y, init_hyperparameters=THETA_C3, noise_variances=noise, args={"dense_sharded": False})
                    # both paths in their default arithmetic: INT8-slice trailing updates (tcgen05) where the blocks are large
                    # enough, DMMA elsewhere
                    l1, g1, dt = single_gpu_check(one, th)
                    rec["single_gpu"] = {"seconds_per_step": dt, "lml": l1, "grad": [float(g) for g in g1],
                                         "arithmetic": "default (INT8-slice trailing updates on both paths)",
                                         "lml_rel_diff": abs(rec["lml"] / l1 - 1), "grad_rel_diff": relerr(rec["grad"], g1),
                                         "agree_1e-8": bool(abs(rec["lml"] / l1 - 1) <= 1e-8 and relerr(rec["grad"], g1) <= 1e-8)}
                    rec["strong_scaling_efficiency"] = dt / (world * rec["seconds_per_step"])
                    if lib_.fvgp_ozaki_slices() > 0 and deadline.left() > 150:  # the all-DMMA single-GPU step, for reference
                        old = lib_.fvgp_set_ozaki(0)
                        try:
                            l2, g2, dt2 = single_gpu_check(one, th)
                        finally:
                            lib_.fvgp_set_ozaki(old)
                        rec["single_gpu_dmma_only"] = {"seconds_per_step": dt2, "lml_rel_diff": abs(rec["lml"] / l2 - 1),
                                                       "grad_rel_diff": relerr(rec["grad"], g2)}
                except Exception as exc:
                    import traceback
                    traceback.print_exc()
                    rec["single_gpu_error"] = f"{type(exc).__name__}: {exc}"[:300]
                    import gc
                    import torch as _t
                    gc.collect()
                    _t.cuda.empty_cache()
            parallel.barrier()
        return rec
    if deadline.allows(200):
        guarded("c3_dense_block_cyclic", c3, out)

    # ---- C5: N = 200 000 (320 GB) on 8 GPUs; smaller N on fewer GPUs so that KV fills a comparable share of HBM
    def c5():
        n = args.c5_n or {2: 100000, 4: 140000}.get(world, 200000)
        x, y, noise = synthetic_c5(n)
        th0 = np.array([1.0, .2, .2])

        def factory():
            return GP(x, y, init_hyperparameters=th0, noise_variances=noise, args={"dense_sharded": True}), th0
        rec, gp = sharded_dense_record(factory, n, steps=2, warmup=0, with_grad=False,
                                       label=f"C5: exact dense GP, 2-D input, N={n} ({8e-9 * n * n:.0f} GB KV), LML")
        release(gp)
        return rec
    if world >= 4 and deadline.allows(240):
        guarded("c5_dense_block_cyclic", c5, out)

    # ---- C4 with the CSR rows and the SLQ probes sharded over the ranks
    def c4s():
        return c4_record(args.c4_n, steps=5, warmup=2, rank=rank, world=world, phases=False, sharded=True)
    if deadline.allows(120):
        guarded("c4_gp2scale_sharded", c4s, out)
    return out


def run_c5(args):
    """--workload c5: the block-cyclic dense line on its own."""
    from fvgp_b200 import GP, parallel
    rank, local_rank, world = parallel.init()
    n = args.n
    x, y, noise = synthetic_c5(n)
    th0 = np.array([1.0, .2, .2])

    def factory():
        return GP(x, y, init_hyperparameters=th0, noise_variances=noise, args={"dense_sharded": True}), th0
    rec, gp = sharded_dense_record(factory, n, args.steps, args.warmup, args.grad, f"C5-type dense, N={n}")
    if rank != 0:
        return
    what = "LML + gradient" if args.grad else "LML"
    line = {"metric": f"dense {what} evals/s, KV 2-D block-cyclic over the GPUs (N={n})", "value": rec["evals_per_s"],
            "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": rec["seconds_per_step"] * 1e3, "higher_is_better": True,
            "scaling": "strong within a run (one matrix over all GPUs); default N grows with the GPU count",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C5-type: exact dense GP, 2-D input, N={n} ({8e-9 * n * n:.0f} GB KV), block-cyclic "
                                   f"Cholesky{' + inverse + gradient traces' if args.grad else ''} over NCCL",
                       "n": n, "grid": rec["grid"], "block": rec["block"], "local_GB": rec["local_GB"]},
            "detail": rec,
            "roofline": {"bound": "tensor", "kernel": "dgemm_mma_kernel (DMMA.8x8x4), per GPU", "achieved": rec["tflops_per_gpu"],
                         "peak": rec["peak_tflops"], "unit": "TFLOP/s", "frac": rec["frac_of_fp64_tensor_peak"], "traffic": None,
                         "algorithmic_flops_per_step": rec["algorithmic_flops_per_step"]}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ C1
def c1_record(n, B, steps, warmup, cpu=True):
    """What train() does at the reference's own CPU-runnable size (BASELINE config 1): populations of
    hyperparameter proposals on a 1-D, N = 1000 GP.  A step = one population of B LML evaluations through
    GP.log_likelihood_population (host thetas in, host LMLs out)."""
    import torch
    from fvgp_b200 import GP, ops
    from fvgp_b200 import _lib as L
    lib = L.load()
    x, y, noise = synthetic_c1(n)
    h0 = np.array([1.0, 0.3])
    gp = GP(x, y, init_hyperparameters=h0, noise_variances=noise)

    def thetas(k):
        return h0 * (0.6 + 0.8 * np.random.default_rng(k).random((B, 2)))

    for k in range(warmup):
        gp.log_likelihood_population(thetas(k))
    sampler = ClockSampler(0)
    torch.cuda.synchronize()
    sampler.start()
    launches0 = lib.fvgp_launch_count()
    ops.start_phase_timing()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    per_step = []
    for k in range(steps):
        t_s = time.perf_counter()
        lml = gp.log_likelihood_population(thetas(warmup + k))        # synchronises internally (host LMLs out)
        per_step.append(time.perf_counter() - t_s)
    e1.record()
    torch.cuda.synchronize()
    t_gpu = ops.stop_phase_timing().get("population", 0.0) / steps      # first to last kernel of the C-ABI call
    t_pop = e0.elapsed_time(e1) * 1e-3 / steps
    launches = (lib.fvgp_launch_count() - launches0) // steps
    clocks = sampler.summary()
    gp.marginal_likelihood.evaluate_population(thetas(0), with_gradient=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        gp.marginal_likelihood.evaluate_population(thetas(warmup + k), with_gradient=True)
    torch.cuda.synchronize()
    t_pop_grad = (time.perf_counter() - t0) / steps
    T = thetas(warmup)
    for t in T[:3]:
        gp.log_likelihood(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    one = np.array([gp.log_likelihood(t) for t in T])
    torch.cuda.synchronize()
    t_seq = time.perf_counter() - t0
    bitwise = bool(np.array_equal(one, gp.log_likelihood_population(T)))
    cpu_rec, parity = None, None
    if cpu:
        from oracle import fvgp_oracle as orc
        use_all_host_threads()
        orc.dense_log_likelihood(x, y, T[0], noise)
        t0 = time.perf_counter()
        ref = [orc.dense_log_likelihood(x, y, t, noise) for t in T[:10]]
        per = (time.perf_counter() - t0) / 10
        parity = {"lml_max_rel_vs_oracle": relerr(one[:10], ref), "tol": 1e-8, "pass": bool(relerr(one[:10], ref) <= 1e-8),
                  "population_bitwise_equals_one_at_a_time": bitwise}
        cpu_rec = {"value": 1.0 / per, "unit": "evals/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"oracle port, 10 of the {B} proposals at N={n}, {per * 1e3:.1f} ms each"}
    flops = B * (n ** 3 / 3.0 + 2.0 * n * n)
    return {"metric": f"LML evals/s (dense, N={n}, 1D, populations of {B} proposals)", "value": B / t_pop,
            "unit": "evals/s", "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": t_pop * 1e3,
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
            "cpu_baseline": cpu_rec, "parity": parity, "last_lml": float(lml[-1])}


def run_c1(args):
    rec = c1_record(args.n, args.population, args.steps, args.warmup, cpu=not args.no_cpu_baseline)
    rec.update({"higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic"})
    print(json.dumps(rec), flush=True)


# ------------------------------------------------------------------------------------------ parity at size
def parity_c2(args, gp, deadline):
    """GPU results at the BENCHMARKED sizes against the oracle on the host cores."""
    from fvgp_b200 import GP
    from oracle import fvgp_oracle as orc
    use_all_host_threads()
    out = {}
    n = args.n
    x, y, noise = synthetic_c2(n)
    th = theta_k(3)
    # gradient at N = 8 000 and 16 000 against the oracle (LAPACK dpotrf / dpotri on the host, blocked dK)
    for ng in (8000, 16000):
        if deadline.left() < (60 if ng == 8000 else 150):
            out[f"c2_grad_n{ng}"] = {"skipped": "time budget"}
            continue
        xs, ys, vs = synthetic_c2(ng)
        g = GP(xs, ys, init_hyperparameters=theta_k(0), noise_variances=vs)
        lml, grad = g.log_likelihood(th), g.neg_log_likelihood_gradient(th)
        release(g)
        t0 = time.perf_counter()
        lml_ref, grad_ref = orc.dense_neg_log_likelihood_gradient_blocked(xs, ys, th, vs)
        out[f"c2_grad_n{ng}"] = {"n": ng, "lml_rel": abs(lml / lml_ref - 1), "grad_max_rel": relerr(grad, grad_ref), "tol": 1e-8,
                                 "pass": bool(abs(lml / lml_ref - 1) <= 1e-8 and relerr(grad, grad_ref) <= 1e-8),
                                 "oracle_seconds": time.perf_counter() - t0, "grad": [float(v) for v in grad],
                                 "grad_oracle": [float(v) for v in grad_ref]}
    # LML at the full benchmarked N (20 GB on the host; the oracle's gradient does not fit at this size)
    need_ram = 8.0 * n * n * 1.15
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    if avail < need_ram:
        out[f"c2_lml_n{n}"] = {"skipped": f"host RAM: need {need_ram / 1e9:.0f} GB, available {avail / 1e9:.0f} GB"}
    elif deadline.left() < 120:
        out[f"c2_lml_n{n}"] = {"skipped": "time budget"}
    else:
        lml = gp.log_likelihood(th)
        t0 = time.perf_counter()
        lml_ref = orc.dense_log_likelihood_blocked(x, y, th, noise)
        out[f"c2_lml_n{n}"] = {"n": n, "ours": lml, "oracle": lml_ref, "rel": abs(lml / lml_ref - 1), "tol": 1e-8,
                               "pass": bool(abs(lml / lml_ref - 1) <= 1e-8), "oracle_seconds": time.perf_counter() - t0,
                               "host_cores": os.cpu_count()}
    return out


def parity_c4_blocks(x, th, K_host, blocks, batch=10000, threads=8):
    """Seeded blocks of the assembled CSR against the DEFINING dense block kernel (kernels.py:502-528 through
    oracle.wendland_block; pattern = np.nonzero, gp2Scale_covariance.py:147): pattern bit-exact, values <= 1e-12."""
    import concurrent.futures as cf
    from oracle import fvgp_oracle as orc

    def one(ij):
        i, j = ij
        r0, r1, c0, c1 = i * batch, min((i + 1) * batch, len(x)), j * batch, min((j + 1) * batch, len(x))
        ref = orc.wendland_block(x[r0:r1], x[c0:c1], th)
        rr, cc = np.nonzero(ref)
        sub = K_host[r0:r1].tocsc()[:, c0:c1].tocsr()
        sub.sort_indices()
        ro, co = sub.nonzero()
        same = len(rr) == len(ro) and np.array_equal(rr, ro) and np.array_equal(cc, co)
        vrel = 0.0
        if same and len(rr):
            vrel = relerr(np.asarray(sub[rr, cc]).ravel(), ref[rr, cc])
        return same, vrel, len(rr)
    with cf.ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(one, blocks))
    return {"blocks": len(blocks), "pattern_bit_exact": bool(all(r[0] for r in res)),
            "values_max_rel": float(max((r[1] for r in res), default=0.0)), "entries_compared": int(sum(r[2] for r in res)),
            "nonempty_blocks": int(sum(1 for r in res if r[2] > 0))}


def choose_blocks(K_host, n, count, batch=10000, seed=0):
    """count seeded (row-block, column-block) pairs: 2/5 diagonal, 2/5 non-empty off-diagonal, 1/5 uniformly random."""
    rng = np.random.default_rng(seed)
    nb = (n + batch - 1) // batch
    coo_r = np.repeat(np.arange(n) // batch, np.diff(K_host.indptr))
    pairs = np.unique(np.stack([coo_r, K_host.indices // batch], axis=1), axis=0)
    off = pairs[pairs[:, 0] != pairs[:, 1]]
    out = [(int(i), int(i)) for i in rng.choice(nb, size=min(nb, 2 * count // 5), replace=False)]
    if len(off):
        out += [tuple(int(v) for v in off[k]) for k in rng.choice(len(off), size=min(len(off), 2 * count // 5), replace=False)]
    while len(out) < count:
        out.append((int(rng.integers(nb)), int(rng.integers(nb))))
    return out


def parity_c4(args, deadline, nblocks):
    from fvgp_b200 import GP, ops
    from fvgp_b200 import _lib as L
    from oracle import fvgp_oracle as orc
    out = {}
    n = args.c4_n
    if deadline.left() > 150:
        x, y, noise = synthetic_c4(n)
        th = theta_c4(0, n)
        xd = L.to_dev(x)
        K_host = ops.wendland_csr(xd, xd, th).to_scipy()
        blocks = choose_blocks(K_host, n, nblocks)
        t0 = time.perf_counter()
        rec = parity_c4_blocks(x, th, K_host, blocks, threads=min(8, os.cpu_count()))
        rec.update({"n": n, "nnz": int(K_host.nnz), "symmetric_pattern": bool((abs(K_host - K_host.T)).nnz == 0) if n <= 200000 else None,
                    "oracle_seconds": time.perf_counter() - t0, "pass": bool(rec["pattern_bit_exact"] and rec["values_max_rel"] <= 1e-12),
                    "tol": "pattern bit-exact, values 1e-12"})
        out[f"c4_blocks_n{n}"] = rec
        del K_host, xd
    else:
        out[f"c4_blocks_n{n}"] = {"skipped": "time budget"}
    # N = 50 000 LML with the exact sparse LU (the reference's sparseLU mode) against the oracle
    if deadline.left() > 120:
        ns = 50000
        x, y, noise = synthetic_c4(ns)
        th = theta_c4(0, ns) * np.array([1.0, 0.4, 0.4, 0.4])          # ~7 nnz/row: SuperLU's fill-in stays at seconds
        gp = GP(x, y, init_hyperparameters=th, noise_variances=noise, gp2Scale=True, linalg_mode="sparseLU")
        lml = gp.log_likelihood(th)
        K = gp.K
        release(gp)
        t0 = time.perf_counter()
        Kref = orc.gp2scale_covariance(x, x, th, batch=2500, symmetric=True, threads=min(16, os.cpu_count()))
        lml_ref = orc.log_likelihood_from(*_lu(orc, Kref, noise, y), (y - y.mean())[:, None])
        out["c4_sparselu_n50000"] = {"n": ns, "nnz": int(K.nnz), "pattern_bit_exact": bool(np.array_equal(K.indptr, Kref.indptr) and np.array_equal(K.indices, Kref.indices)),
                                     "ours": lml, "oracle": lml_ref, "rel": abs(lml / lml_ref - 1), "tol": 1e-8,
                                     "pass": bool(abs(lml / lml_ref - 1) <= 1e-8 and np.array_equal(K.indices, Kref.indices)),
                                     "oracle_seconds": time.perf_counter() - t0}
    else:
        out["c4_sparselu_n50000"] = {"skipped": "time budget"}
    return out


def _lu(orc, K, noise, y):
    KV = orc.add_kv(K, noise)
    alpha, logdet = orc.sparse_lu_solve_logdet(KV, (y - y.mean())[:, None])
    return alpha.reshape(-1, 1), logdet


def parity_c3(points=2000):
    """C3 shape at 2000 points x 5 tasks through fvGP(..., dense_sharded) on this rank's grid vs the oracle."""
    from fvgp_b200 import fvGP
    from oracle import fvgp_oracle as orc
    x, y, noise = synthetic_c3(points)
    gp = fvGP(x, y, init_hyperparameters=THETA_C3, noise_variances=noise, args={"dense_sharded": True})
    th = THETA_C3 * 1.03
    lml, grad = gp.log_likelihood(th), gp.neg_log_likelihood_gradient(th)
    release(gp)
    xi, yi, vi = orc.fvgp_transform(x, y, noise)
    lml_ref, grad_ref = orc.dense_neg_log_likelihood_gradient_blocked(xi, yi, th, vi)
    return {"n": len(xi), "lml_rel": abs(lml / lml_ref - 1), "grad_max_rel": relerr(grad, grad_ref), "tol": 1e-8,
            "pass": bool(abs(lml / lml_ref - 1) <= 1e-8 and relerr(grad, grad_ref) <= 1e-8)}


# ------------------------------------------------------------------------------------------ headline
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
            "arithmetic": "FP64 results throughout; for N >= 40 000 the large products of the Cholesky factorisation (trailing "
                          "updates) and of the inversion (TRTRI products, LAUUM) run as error-free INT8-slice products on the "
                          "tcgen05 tensor cores (8 x 6-bit slices per entry, exact int32 accumulation, FP64 recombination; "
                          "FVGP_OZAKI=0 keeps everything on the FP64 tensor pipe): LML / gradient within 2e-12 / 7e-10 of the "
                          "pure FP64 path (`int8_trailing_updates_ab`), oracle parity at N = 50 000 in `parity` and, with the "
                          "size gate lowered, at N = 16 000 in tests/test_gpu_parity_at_size.py",
            "l2_policy": f"inputs larger than L2: K is {8 * args.n ** 2 / 1e9:.1f} GB"}


def i8_gemm_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the int8 GEMM at the `kernel_alone` shape, from the
    committed ncu --set full capture (profiles/r02/ncu_i8_gemm.v19.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02", "ncu_i8_gemm.v19.json")) as fh:
            d = json.load(fh)
        return float(d["dram_bytes"]), (f"per launch at m, n, K = {d['shape_mnk']} (the `kernel_alone` shape): {d['dram_bytes'] / 1e9:.1f} GB "
                                        f"of DRAM traffic vs {d['algorithmic_bytes'] / 1e9:.2f} GB algorithmic (operands once + int32 "
                                        f"output): operand panels are re-read across tiles (L2 hit rate {d['l2_hit_rate_pct']:.0f} %), "
                                        f"DRAM at {d['dram_throughput_pct']:.0f} % of peak, tensor pipe active "
                                        f"{d['tensor_pipe_active_pct_elapsed']:.1f} % of the launch: tensor-bound ({d['source']})")
    except Exception as exc:
        return None, f"no capture on file ({exc})"


def kfill_rooflines(n, gp, x, noise):
    """K-assembly against HBM: the symmetric mode (full square + noise diagonal, 8 N^2 bytes written) and FILL_LOWER,
    the mode the LML path uses (tiles on / below the diagonal only: 4 N^2 + 256 N bytes)."""
    import torch
    from fvgp_b200 import ops
    from fvgp_b200 import _lib as L
    peaks = load_peaks()
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    out_buf = L.dev_matrix(n, n)
    xd, nd = gp.data.x_device(), L.to_dev(noise)
    th = theta_k(1)
    bounds = (x.min(axis=0), x.max(axis=0))      # host reductions stay outside the event pair
    recs = {}
    for mode, name in ((L.FILL_SYMMETRIC, "symmetric"), (L.FILL_LOWER, "lower")):
        best, _ = event_timed(lambda: ops.kfill(L.K_MATERN32, xd, xd, th[0], 1 / th[1:], 1.0, noise=nd, mode=mode,
                                                out=out_buf, bounds=bounds), reps=5)
        tiles = (n + 63) // 64
        nbytes = 8.0 * n * n if mode == L.FILL_SYMMETRIC else 8.0 * 64 * 64 * tiles * (tiles + 1) / 2
        recs[name] = {"seconds": best, "bytes": nbytes, "GB/s": nbytes / best / 1e9, "frac": nbytes / best / 1e9 / hbm_peak,
                      "entries_evaluated_per_s": (n * (n + 64.0) / 2) / best}
    del out_buf
    torch.cuda.empty_cache()
    sym = recs["symmetric"]
    return {"bound": "hbm", "kernel": "kfill_kernel<MATERN32,3> symmetric (full square + noise diagonal)",
            "achieved": sym["GB/s"], "peak": hbm_peak, "unit": "GB/s", "frac": sym["frac"],
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch at N = 50 000
            # (ncu --set full, profiles/r01/ncu_kfill.v6.summary.txt): 0.081 + 19.941 GB
            "traffic": 20.022e9 if n == 50000 else None, "peak_source": hbm_src, "algorithmic_bytes": 8.0 * n * n,
            "lower_mode_on_the_lml_path": {"achieved": recs["lower"]["GB/s"], "frac": recs["lower"]["frac"],
                                           "seconds": recs["lower"]["seconds"], "algorithmic_bytes": recs["lower"]["bytes"],
                                           "note": "half the bytes, the same number of evaluated entries: FP64-issue bound"},
            "symmetric_seconds": sym["seconds"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--size", "--n", dest="n", type=int, default=50000,
                    help="problem size N (use --size under torchrun: its own parser rejects a bare --n as ambiguous)")
    ap.add_argument("--cpu-sample-n", type=int, default=2000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c4", "c5"],
                    help="c2 (default, the headline; carries c4 / c1 / parity / sharded sub-records): dense N=50k LML+gradient; "
                         "c4: gp2Scale N=1M LML; c5: dense LML with KV block-cyclic over all GPUs; c1: populations at N=1000")
    ap.add_argument("--population", type=int, default=40, help="c1 only: proposals per step")
    ap.add_argument("--grad", action="store_true", help="c5 only: add the gradient to every step")
    ap.add_argument("--sharded", action="store_true", help="c4 only: shard CSR rows and SLQ probes over the ranks")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip parity / c4 / c1 / sharded sub-records")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-fresh-c4", action="store_true", help="skip the C4 measurement in a fresh subprocess")
    ap.add_argument("--parity-blocks", type=int, default=12, help="C4 blocks compared with the oracle inside bench.py "
                    "(the -m gpu suite compares 50)")
    ap.add_argument("--c4-n", type=int, default=1000000)
    ap.add_argument("--c4-cpu-n", type=int, default=200000)
    ap.add_argument("--c3-points", type=int, default=20000)
    ap.add_argument("--c5-n", type=int, default=0)
    ap.add_argument("--no-c4", dest="with_c4", action="store_false", help="reference arm: skip the gp2Scale sub-record")
    ap.add_argument("--budget", type=float, default=float(os.environ.get("FVGP_BENCH_BUDGET_S", "740")),
                    help="seconds after which optional sub-records are skipped")
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
    deadline = Deadline(args.budget)
    n = args.n
    x, y, noise = synthetic_c2(n)
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
    macs0 = lib.fvgp_ozaki_mac_count()
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
    i8_macs = (lib.fvgp_ozaki_mac_count() - macs0) / args.steps       # int8 multiply-accumulates per step (counted per launch)
    clocks = sampler.summary()
    t_dev = parallel.max_over_ranks(e0.elapsed_time(e1) * 1e-3)

    log(f"device-timed region done: {t_dev / args.steps:.3f} s per step")
    # ---- end to end: host buffers in, host results out, through the public API -------------------
    gp.args = dict(gp.args, host_inputs_every_call=True)           # every call copies x from pinned host memory
    step(args.warmup + args.steps)                                  # pins the staging buffer (untimed)
    parallel.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        lml, grad = step(args.warmup + args.steps + 1 + k)
        assert np.isfinite(lml) and np.all(np.isfinite(grad))
    torch.cuda.synchronize()
    t_e2e = parallel.max_over_ranks(time.perf_counter() - t0)
    gp.args = {k: v for k, v in gp.args.items() if k != "host_inputs_every_call"}
    h2d = 2 * x.nbytes + 2 * n * 8                                  # x (LML call and gradient call), y - m, noise diagonal
    d2h = n * 8 + 8 + 4 + H * 8                                     # KVinvY, logdet, potrf status, traces

    line = {"metric": METRIC_C2,
            "value": world * args.steps / t_dev, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args), "clocks": clocks,
            "e2e": {"value": world * args.steps / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "last_lml": out[0], "last_grad": [float(g) for g in out[1]]}
    HOLD["line"] = line
    start_watchdog(args.budget + 60.0, rank)

    log(f"e2e region done: {t_e2e / args.steps:.3f} s per step")
    # ---- roofline of the dominant kernel (DMMA GEMM inside potrf + potri) ----------------------
    if rank == 0:
        import ctypes
        scratch = L.dev_empty((148 * 8 * 256,))
        pk = ctypes.c_double()
        lib.fvgp_bench_fp64_peak(0, 2, 20000, L.ptr(scratch), ctypes.byref(pk), L.stream_ptr())
        t_potrf, t_potri = phases.get("potrf", 0.0) / args.steps, phases.get("potri", 0.0) / args.steps
        t_tensor = t_potrf + t_potri
        achieved = n ** 3 / t_tensor / 1e12
        step_tflops = n ** 3 / (t_dev / args.steps) / 1e12
        potri_tflops = 2.0 * n ** 3 / 3.0 / t_potri / 1e12
        int8_on = i8_macs > 0
        peaks = load_peaks()
        fp64 = {"what": "N^3 flop of POTRF + POTRI over their time, against the DMMA (FP64 tensor pipe) issue rate measured live "
                        "(fvgp_bench_fp64_peak).  With the INT8-slice products on, these are FP64-EQUIVALENT flop: most of "
                        "them are executed as int8 MMAs, so the fraction can pass 1; the all-DMMA figures are in "
                        "`int8_trailing_updates_ab` (and in round 1's line)",
                "achieved": achieved, "peak": pk.value, "unit": "TFLOP/s", "frac": achieved / pk.value,
                "potrf": {"seconds": t_potrf, "achieved": n ** 3 / 3.0 / t_potrf / 1e12},
                "potri": {"seconds": t_potri, "achieved": potri_tflops},
                "whole_step": {"achieved": step_tflops, "frac": step_tflops / pk.value,
                               "note": "N^3 flop over ms_per_step (fill, solves, traces and host time included)"},
                "peak_source": "measured live: register-resident DMMA.8x8x4 issue rate (fvgp_bench_fp64_peak)"}
        if int8_on:
            # dominant kernel: the int8 GEMM (tcgen05.mma kind::i8, CTA pair, 256 x 256 x 128 tile) behind the INT8-slice
            # products.  achieved = 2 x the multiply-accumulates actually launched per step (counted per GEMM launch in
            # csrc/ozaki.cu) over the time of the two phases that contain them -- slicing, recombination, panels, tile
            # kernels and the remaining DMMA products included, i.e. a LOWER bound on the kernel's own rate.  peak: the
            # dense int8 rate is twice the bf16 rate on this part; MEASURED_PEAKS.json holds the bf16 figures.
            bf16 = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
            peak8, src8 = ((2.0 * bf16, "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (kernel timed inside a long step)")
                           if bf16 else (4500.0, "fallback: nominal dense INT8 rate of B200"))
            ach8 = 2.0 * i8_macs / t_tensor / 1e12
            burst = lib.fvgp_ozaki_i8_seconds(6272, 24912, 100352, 3, 3, None)
            line["roofline"] = {"bound": "tensor",
                                "kernel": "int8 GEMM of the INT8-slice products (CuTe/CUTLASS sm100 collective: TMA + tcgen05.mma "
                                          "kind::i8 cta_group::2, 256x256x128, int32 accumulators in TMEM) inside potrf + potri",
                                "achieved": ach8, "peak": peak8, "unit": "TFLOP/s", "frac": ach8 / peak8,
                                "op": "int8 multiply-accumulate = 2 ops (TOP/s)", "peak_source": src8,
                                "nominal_peak": 4500.0, "frac_of_nominal": ach8 / 4500.0,
                                "int8_macs_per_step": i8_macs,
                                "algorithmic_note": "36 int8 MACs per FP64 MAC (8 slices: 8*9/2 slice pairs); the step's FP64 "
                                                    "work is N^3/2 MACs, of which the INT8 path covers the products with >= 8192 "
                                                    "rows; chunking of the triangular products adds 1/8 on those",
                                "kernel_alone": ({"shape_mnk": [6272, 24912, 100352], "seconds": burst,
                                                  "achieved": 2.0 * 6272 * 24912 * 100352 / burst / 1e12,
                                                  "frac_of_nominal": 2.0 * 6272 * 24912 * 100352 / burst / 1e12 / 4500.0,
                                                  "note": "one TRTRI-chunk-shaped launch timed alone with CUDA events (burst clocks)"}
                                                 if burst > 0 else None),
                                "fp64_equivalent": fp64,
                                "potrf": {"seconds": t_potrf, "achieved_fp64_equivalent": n ** 3 / 3.0 / t_potrf / 1e12,
                                          "int8_trailing_updates": True},
                                "potri": {"seconds": t_potri, "achieved": potri_tflops, "frac": potri_tflops / pk.value,
                                          "what": "TRTRI + LAUUM, 2 N^3 / 3 flop, FP64-equivalent"},
                                "whole_step": fp64["whole_step"],
                                "traffic": i8_gemm_traffic()[0], "traffic_note": i8_gemm_traffic()[1],
                                "algorithmic_flops_per_step": float(n) ** 3,
                                "phase_seconds_per_step": {k: v / args.steps for k, v in phases.items()}}
        else:
            line["roofline"] = {"bound": "tensor", "kernel": "dgemm_mma_kernel (DMMA.8x8x4) inside potrf + potri",
                                "achieved": achieved, "peak": pk.value, "unit": "TFLOP/s", "frac": achieved / pk.value,
                                "potri": {"what": "TRTRI + LAUUM, 2 N^3 / 3 flop", "achieved": potri_tflops,
                                          "frac": potri_tflops / pk.value},
                                "potrf": {"seconds": t_potrf, "achieved_fp64_equivalent": n ** 3 / 3.0 / t_potrf / 1e12,
                                          "int8_trailing_updates": False},
                                "whole_step": fp64["whole_step"],
                                # dram__bytes_read.sum + dram__bytes_write.sum summed over every dgemm_mma_kernel launch of
                                # ONE timed evaluation (ncu launch list, profiles/*/launches_bench_n50k.*.json)
                                "traffic": gemm_dram_traffic(n)[0], "traffic_note": gemm_dram_traffic(n)[1],
                                "peak_source": fp64["peak_source"],
                                "algorithmic_flops_per_step": float(n) ** 3,
                                "phase_seconds_per_step": {k: v / args.steps for k, v in phases.items()}}
        log("fp64 peak probe done; K-fill rooflines ...")
        line["roofline_kfill"] = kfill_rooflines(n, gp, x, noise)
        log("same-N point ...")

        # the size both arms measure directly: ours through the public API with host buffers
        xs, ys, vs = synthetic_c2(SAME_N)
        g8 = GP(xs, ys, init_hyperparameters=theta_k(0), noise_variances=vs, args={"host_inputs_every_call": True})
        for k in range(2):
            g8.log_likelihood(theta_k(k + 1)), g8.neg_log_likelihood_gradient(theta_k(k + 1))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(5):
            g8.log_likelihood(theta_k(k + 3)), g8.neg_log_likelihood_gradient(theta_k(k + 3))
        torch.cuda.synchronize()
        dt8 = (time.perf_counter() - t0) / 5
        release(g8)
        line["same_n"] = {"n": SAME_N, "seconds": dt8, "evals_per_s": 1.0 / dt8,
                          "note": "LML + gradient through the public API, host buffers; the reference arm measures this size too"}

        log("cpu baseline sample ...")
        line["cpu_baseline"] = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import fvgp_oracle as orc
            use_all_host_threads()
            ns = 3000
            xs, ys, vs = synthetic_c2(ns)
            t0 = time.perf_counter()
            cpu_port_step(xs, ys, vs, theta_k(1), orc)
            per = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1.0 / (per * (n / ns) ** 3), "unit": "evals/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"oracle port of the reference algorithm, one LML+gradient at N={ns} took {per:.2f} s on "
                                              f"{os.cpu_count()} host threads, extrapolated x(N/{ns})^3 to N={n}; the reference arm "
                                              f"(--impl reference) times the unmodified reference over a ladder of sizes instead"}

    extras = not args.no_extras
    # ---- A/B of the INT8-slice trailing updates (default on at this size): the same steps on the DMMA pipe only ------
    if extras and rank == 0 and lib.fvgp_ozaki_available() and deadline.left() > 120:
        def ozaki_section():
            th = theta_k(4)
            gp.kv._memo = None
            on_lml, on_grad = gp.log_likelihood(th), gp.neg_log_likelihood_gradient(th)
            old = lib.fvgp_set_ozaki(0)
            rec = {"slices_default": int(old), "what": "trailing updates of the look-ahead POTRF, SYRK of LAUUM and the chunked "
                   "triangular products of TRTRI / LAUUM (>= 8192 rows, N >= 40 000) as INT8-slice GEMMs on tcgen05 (kind::i8, "
                   "TMEM) vs everything on the DMMA pipe (fvgp_set_ozaki(0))"}
            try:
                ops.start_phase_timing()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for k in range(3):
                    o = step(100 + k)
                torch.cuda.synchronize()
                rec["dmma_seconds_per_step"] = (time.perf_counter() - t0) / 3
                ph = ops.stop_phase_timing()
                rec["dmma_potrf_seconds"] = ph.get("potrf", 0.0) / 3
                rec["dmma_potri_seconds"] = ph.get("potri", 0.0) / 3
                rec["int8_potri_seconds"] = phases.get("potri", 0.0) / args.steps
                gp.kv._memo = None
                lml, grad = gp.log_likelihood(th), gp.neg_log_likelihood_gradient(th)
                rec.update({"int8_seconds_per_step": t_dev / args.steps, "int8_potrf_seconds": phases.get("potrf", 0.0) / args.steps,
                            "int8_potrf_tflops_fp64_equivalent": n ** 3 / 3.0 / (phases.get("potrf", 1e30) / args.steps) / 1e12,
                            "dmma_potrf_tflops": n ** 3 / 3.0 / rec["dmma_potrf_seconds"] / 1e12,
                            "lml_rel_diff_int8_vs_dmma": abs(on_lml / lml - 1),
                            "grad_max_rel_diff_int8_vs_dmma": relerr(on_grad, grad),
                            "within_1e-8": bool(abs(on_lml / lml - 1) <= 1e-8 and relerr(on_grad, grad) <= 1e-8)})
                assert np.isfinite(o[0])
            finally:
                lib.fvgp_set_ozaki(old)
                gp.kv._memo = None
            return rec
        guarded("int8_trailing_updates_ab", ozaki_section, line)
    # ---- parity at the benchmarked sizes (one GPU; the host needs its cores) ----------------------
    if extras and world == 1 and not args.no_parity:
        par = {}

        def merge(name, fn):
            tmp = {}
            guarded(name, fn, tmp)
            res = tmp[name]
            if "error" in res:
                par[name + "_error"] = res
            else:
                res.pop("section_seconds", None)
                par.update(res)
        merge("c2", lambda: parity_c2(args, gp, deadline))
        release(gp)
        gp = None
        merge("c4", lambda: parity_c4(args, deadline, args.parity_blocks))
        guarded("c3_n10000_fvgp_dense_sharded_1rank", parity_c3, par)
        checks = [v for v in par.values() if isinstance(v, dict) and "skipped" not in v]
        par["all_pass"] = bool(checks and all(v.get("pass", False) for v in checks))
        line["parity"] = par
        HOLD["line"] = line
    if gp is not None:
        release(gp)
        gp = None

    # ---- the other workloads of BASELINE.json's metric --------------------------------------------
    if extras and world == 1:
        if deadline.allows(60):
            guarded("c4", lambda: c4_record(args.c4_n, steps=max(5, min(args.steps, 10)), warmup=3), line)
        if deadline.allows(30):
            guarded("c1", lambda: c1_record(1000, 40, 20, 3), line)
        if deadline.allows(90) and not args.no_fresh_c4:
            # the same C4 workload in a fresh process of the same run: inside this process the step has been measured
            # ~25 ms slower than standalone (host state after the dense phases and the host oracles); both are reported
            def fresh_c4():
                cmd = [sys.executable, os.path.abspath(__file__), "--workload", "c4", "--steps", "8", "--warmup", "3",
                       "--size", str(args.c4_n)]
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=max(60.0, deadline.left()))
                rows = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
                if not rows:
                    raise RuntimeError("no JSON line from the C4 subprocess: " + out.stderr[-300:])
                d = json.loads(rows[-1])
                return {k: d.get(k) for k in ("value", "unit", "ms_per_step", "steps", "warmup", "last_lml", "log_likelihood_std",
                                              "phases_seconds", "gpu_launches")}
            guarded("c4_fresh_process", fresh_c4, line)
    if extras and world > 1:
        sh = sharded_section(args, deadline, rank, world)
        line["sharded"] = sh
    line["wall_seconds"] = round(time.time() - T_START, 1)
    emit(line, rank)
    parallel.barrier()


if __name__ == "__main__":
    main()
