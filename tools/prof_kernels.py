"""Launches the hot kernels once each at their UNet shapes (for `ncu --set full -k regex:...`).
Also prints CUDA-event timings of the same shapes when run without a profiler."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context

ctx = Context(0)
iters = int(os.environ.get("ITERS", "1"))
shapes = [("conv", (1, 64, 64, 320, 320, 3, 1)), ("conv", (1, 64, 64, 640, 320, 3, 1)), ("conv", (1, 32, 32, 640, 640, 3, 1)),
          ("conv", (1, 16, 16, 2560, 1280, 3, 1)), ("conv", (1, 16, 16, 1280, 1280, 3, 1)),
          ("gemm", (4096, 320, 320, 1, 0)), ("gemm", (4096, 960, 320, 1, 0)), ("gemm", (4096, 2560, 320, 1, 1)), ("gemm", (4096, 320, 1280, 1, 0)),
          ("gemm", (1024, 5120, 640, 1, 1)), ("gemm", (1024, 640, 2560, 1, 0)), ("gemm", (256, 10240, 1280, 1, 1)), ("gemm", (256, 1280, 5120, 1, 0))]
for kind, a in shapes:
    if kind == "conv":
        n, h, w, cin, cout, k, s = a
        ms = ctx.bench_conv(n, h, w, cin, cout, k, s, iters=iters)
        fl = 2.0 * n * h * w * cout * cin * k * k
    else:
        m, n, k, b, geglu = a
        ms = ctx.bench_gemm(m, n, k, b, geglu, iters=iters)
        fl = 2.0 * m * n * k * b
    print(f"{kind} {a}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
rng = np.random.default_rng(0)
for (h, tq, tk, d) in ((8, 4096, 4096, 40), (8, 4096, 77, 40), (8, 1024, 1024, 80), (8, 256, 256, 160)):
    q = rng.standard_normal((h, tq, d), dtype=np.float32)
    k = rng.standard_normal((h, tk, d), dtype=np.float32)
    v = rng.standard_normal((h, tk, d), dtype=np.float32)
    t0 = time.perf_counter()
    ctx.attention_core(q, k, v)
    print(f"attention h={h} tq={tq} tk={tk} d={d}: host call {1e3 * (time.perf_counter() - t0):.2f} ms (incl. copies)")
