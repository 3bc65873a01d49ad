import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context
ctx = Context(0)
ctx.set_option("autotune", 0)
for halo in (2, 0):
    ctx.set_option("conv_halo", halo)
    for (bn, sp, cg) in ((80, 1, 2), (80, 1, 1), (160, 1, 2), (160, 2, 2)):
        ctx.set_option("gemm_cg", cg)
        for dbg in (0, 8):
            ctx.set_option("gemm_debug", dbg)
            ms = ctx.bench_conv(1, 64, 64, 320, 320, 3, 1, bn, sp, iters=(20 if dbg == 0 else 1))
            ctx.synchronize()
            if dbg == 0: print(f"halo={halo} bn={bn} sp={sp} cg={cg}: {ms*1e3:.1f} us", flush=True)
