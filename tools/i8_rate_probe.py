"""Raw rate of the int8 GEMM behind the INT8-slice path (tcgen05 kind::i8) per tile configuration and shape:
1 = 128x128x128 one SM, 2 = 256x128x128 CTA pair (default), 3 = 256x256x128 CTA pair, 4 = 256x128x128 cluster 2x2.
Nominal dense INT8 peak of B200: 4.5 POP/s = 2.25 PMAC/s."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fvgp_b200 import _lib as L  # noqa: E402

lib = L.load()
torch.cuda.init()
torch.zeros(1, device="cuda")
shapes = [(32768, 4096, 2048), (32768, 4096, 8192), (32768, 4096, 16384), (16384, 16384, 16384), (6272, 24912, 100352),
          (3200, 24912, 200000)]
for m, n, k in shapes:
    line = [f"m={m} n={n} K={k}:"]
    for tile in (1, 2, 3, 4):
        t = lib.fvgp_ozaki_i8_seconds(m, n, k, tile, 3, None)
        line.append(f"tile{tile} {t * 1e3:.2f} ms = {m * n * k / t / 1e15:.2f} PMAC/s" if t > 0 else f"tile{tile} failed ({t})")
    print("  ".join(line), flush=True)
