"""Time the symmetric / lower K-fill (default kernel, D = 3) at the sizes given on the command line."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fvgp_b200 import _lib as L, ops

hog = []
if "--hog" in sys.argv:                      # occupy HBM first (does placement of the output matter?)
    gb = int(sys.argv[sys.argv.index("--hog") + 1])
    hog = [torch.empty(int(1e9), dtype=torch.float64, device="cuda").fill_(1.0) for _ in range(gb // 8)]
    print("hogging", 8 * len(hog), "GB")
if "--after-gemm" in sys.argv:               # heat the chip with ~8 s of DMMA first (do clocks matter?)
    a = torch.randn(16384, 16384, dtype=torch.float64, device="cuda")
    c = torch.zeros_like(a)
    for _ in range(30):
        ops.dgemm_nt(a, a, c)
    torch.cuda.synchronize()
    del a, c
    print("after 30 DGEMMs of 16384^3")
if "--bulk" in sys.argv:                     # mirror path of the symmetric fill: 0 plain, 1 CTA-wide bulk rows, 2 per-warp bulk
    mode_b = int(sys.argv[sys.argv.index("--bulk") + 1])
    L.load().fvgp_set_bulk_store(mode_b)
    print("mirror store mode", mode_b)
sizes = [int(a) for a in sys.argv[1:] if a.isdigit() and (sys.argv[sys.argv.index(a) - 1] not in ("--hog", "--bulk"))]
for n in sizes or [30000, 50000]:
    rng = np.random.default_rng(0)
    x = L.to_dev(rng.random((n, 3)))
    noise = L.to_dev(np.full(n, 1e-2))
    out = L.dev_matrix(n, n)
    th = np.array([1.0, .3, .4, .5])
    for name, mode, nbytes in (("symmetric", L.FILL_SYMMETRIC, 8.0 * n * n), ("lower", L.FILL_LOWER, 4.0 * n * n)):
        best = 1e30
        for _ in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.kfill(L.K_MATERN32, x, x, th[0], 1 / th[1:], 1.0, noise=noise, mode=mode, out=out,
                      bounds=(np.zeros(3), np.ones(3)))
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e-3)
        print(f"band={os.environ.get('FVGP_FILL_BAND', 'default')} n={n} {name}: {best * 1e3:.3f} ms -> {nbytes / best / 1e9:.0f} GB/s", flush=True)
    del out
