"""Lab: attention core timings (tsd_bench_attention, device events) for the first-generation kernel and the
software-pipelined one, per softmax axis; with TSD_ATTN_TRACE=1 the v2 kernel prints its in-kernel cycle counts."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context  # noqa: E402

ctx = Context(0)
shapes = [(8, 4096, 4096, 40), (8, 4096, 77, 40), (8, 1024, 1024, 80), (8, 1024, 77, 80), (8, 256, 256, 160),
          (8, 256, 77, 160), (8, 4096, 4096, 80), (8, 4096, 4096, 160), (16, 4096, 4096, 40)]
iters = int(os.environ.get("ITERS", "20"))
for axis in (0, 1):
    ctx.set_option("softmax_axis", axis)
    for (h, tq, tk, d) in shapes:
        row = []
        for v2, poly in ((0, 0), (1, 0), (1, 1)):
            ctx.set_option("attn_v2", v2)
            ctx.set_option("attn_poly", poly)
            ms = C.c_double()
            ctx._ck(ctx.L.tsd_bench_attention(ctx.h, h, tq, tk, d, iters, C.byref(ms)))
            row.append(ms.value * 1e3)
        fl = 4.0 * h * tq * tk * d
        print(f"axis={'query' if axis == 0 else 'key  '} h={h} tq={tq} tk={tk} d={d}: v1 {row[0]:8.1f} us  v2 {row[1]:8.1f} us  v2+poly {row[2]:8.1f} us  "
              f"({fl / row[2] / 1e6:6.1f} TFLOP/s, exp floor {2.0 * h * tq * tk / (16 * 148 * 1.965e9) * 1e6:6.1f} us)", flush=True)
ctx.set_option("softmax_axis", 0)
ctx.close()
