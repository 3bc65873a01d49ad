"""In-kernel cycle trace (CTA 0) of the most frequent UNet GEMMs with their shipped tile plans.
Needs the trace build: TSD_LAB_TRACE=1 TSD_BUILD_DIR=build_trace TSD_OUT=libtsd_b200_trace.so bash csrc/build.sh,
then TSD_LIB=.../libtsd_b200_trace.so python tools/lab/gemm_trace.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context  # noqa: E402

ctx = Context(0)
ctx.set_option("autotune", 0)
cases = [("gemm", (4096, 320, 320, 0), 80, 1), ("gemm", (4096, 320, 320, 0), 80, 2), ("gemm", (4096, 960, 320, 0), 240, 2),
         ("gemm", (4096, 2560, 320, 1), 256, 1), ("gemm", (4096, 320, 1280, 0), 80, 2),
         ("conv", (1, 64, 64, 320, 320), 80, 2), ("conv", (1, 64, 64, 320, 320), 160, 2),
         ("gemm", (1024, 640, 640, 0), 64, 1), ("gemm", (256, 1280, 1280, 0), 32, 2)]
for kind, shp, bn, cg in cases:
    ctx.set_option("gemm_cg", cg)
    ctx.set_option("gemm_debug", 0)
    if kind == "gemm":
        m, n, k, geglu = shp
        ms = ctx.bench_gemm(m, n, k, 1, geglu, bn, 1, iters=20)
        ctx.set_option("gemm_debug", 8)
        ctx.bench_gemm(m, n, k, 1, geglu, bn, 1, iters=1)
    else:
        n_, h, w, cin, cout = shp
        ms = ctx.bench_conv(n_, h, w, cin, cout, 3, 1, bn, 1, iters=20)
        ctx.set_option("gemm_debug", 8)
        ctx.bench_conv(n_, h, w, cin, cout, 3, 1, bn, 1, iters=1)
    ctx.synchronize()
    print(f"^^ {kind} {shp} bn={bn} cg={cg}: {ms * 1e3:.1f} us (trace build)", flush=True)
