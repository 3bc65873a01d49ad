"""Lab: split-K GEMMs of the 16x16 / 32x32 UNet layers - partial tiles + reduce kernel vs the in-kernel cluster reduction."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context  # noqa: E402
ctx = Context(0)
ctx.set_option("autotune", 0)
trace = os.environ.get("TRACE", "0") == "1"
cases = [("conv", (1, 16, 16, 1280, 1280), 160, 8, 1), ("conv", (1, 16, 16, 1280, 1280), 160, 4, 2), ("conv", (1, 16, 16, 1280, 1280), 160, 8, 2),
         ("conv", (1, 32, 32, 640, 640), 160, 4, 2), ("conv", (1, 32, 32, 640, 640), 128, 3, 2), ("conv", (1, 16, 16, 2560, 1280), 160, 8, 1),
         ("gemm", (256, 1280, 5120), 128, 6, 2), ("gemm", (1024, 640, 2560), 160, 4, 2)]
for kind, shp, bn, sp, cg in cases:
    ctx.set_option("gemm_cg", cg)
    out = []
    for cl in (0, 1):
        ctx.set_option("splitk_cluster", cl)
        ctx.set_option("splitk_cluster_max", 16)
        ctx.set_option("gemm_debug", 0)
        if kind == "conv":
            n_, h, w, cin, cout = shp
            ms = ctx.bench_conv(n_, h, w, cin, cout, 3, 1, bn, sp, iters=20)
            if trace:
                ctx.set_option("gemm_debug", 8)
                ctx.bench_conv(n_, h, w, cin, cout, 3, 1, bn, sp, iters=1)
        else:
            m, n, k = shp
            ms = ctx.bench_gemm(m, n, k, 1, 0, bn, sp, iters=20)
            if trace:
                ctx.set_option("gemm_debug", 8)
                ctx.bench_gemm(m, n, k, 1, 0, bn, sp, iters=1)
        ctx.synchronize()
        out.append(ms * 1e3)
    print(f"{kind} {shp} bn={bn} splits={sp} cg={cg}: partials+reduce {out[0]:6.1f} us | cluster reduce {out[1]:6.1f} us", flush=True)
