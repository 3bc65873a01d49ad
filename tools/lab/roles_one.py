import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context
bn, cg, dbg, pdl = map(int, sys.argv[1:5])
ctx = Context(0)
ctx.set_option("autotune", 0); ctx.set_option("gemm_cg", cg); ctx.set_option("gemm_debug", dbg); ctx.set_option("pdl", pdl)
ms = ctx.bench_conv(1, 64, 64, 320, 320, 3, 1, bn, 1, iters=20)
print(f"bn={bn} cg={cg} dbg={dbg:03b} pdl={pdl}: {ms*1e3:.1f} us", flush=True)
