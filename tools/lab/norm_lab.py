"""Lab: GroupNorm consumer timings (tsd_bench_norm) at the UNet / decoder shapes."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context  # noqa: E402
ctx = Context(0)
for (n, h, w, c, g) in ((1, 64, 64, 320, 32), (1, 32, 32, 640, 32), (1, 16, 16, 1280, 32), (1, 64, 64, 320, 320), (2, 64, 64, 320, 32),
                        (1, 128, 128, 512, 16)):
    row = []
    for mode in (0, 1, 2):
        ms = C.c_double()
        ctx._ck(ctx.L.tsd_bench_norm(ctx.h, n, h, w, c, g, mode, 30, C.byref(ms)))
        row.append(ms.value * 1e3)
    ms = ctx.bench_conv(n, h, w, c, c, 3, 1, iters=30) * 1e3
    mb = n * h * w * c * 4 / 1e6
    print(f"norm n={n} {h}x{w} C={c} G={g} ({mb:.1f} MB): stand-alone {row[0]:6.1f} us | partial-stats consumer {row[1]:6.1f} us | conv+norm pair {row[2]:6.1f} us (conv alone {ms:6.1f} us)", flush=True)
