"""Two launches (warm-up + one) of the int8 GEMM at a TRTRI-chunk shape, for an `ncu --set full` capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fvgp_b200 import _lib as L  # noqa: E402

torch.zeros(1, device="cuda")
m, n, k = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (6272, 24912, 100352)))
t = L.load().fvgp_ozaki_i8_seconds(m, n, k, 3, 1, None)
print(f"m={m} n={n} K={k}: {t * 1e3:.2f} ms; algorithmic bytes {m * k + n * k + 4 * m * n}")
