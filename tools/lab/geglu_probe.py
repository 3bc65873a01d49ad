"""Where does the time of the GEGLU GEMMs (FF1 of the transformer blocks) go?  Timing only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context  # noqa: E402

ctx = Context(0)
ctx.set_option("autotune", 0)
shapes = [(4096, 2560, 320, 1), (1024, 5120, 640, 1), (256, 10240, 1280, 1), (4096, 2560, 320, 0), (4096, 1280, 320, 0)]
for (m, n, k, geglu) in shapes:
    for cg in (1, 2):
        ctx.set_option("gemm_cg", cg)
        for bn in (256, 160, 128, 64):
            line = []
            for dbg in ((0, 1, 6, 7) if cg == 1 else (0, 1, 6)):  # 7 with pairs: known debug-only barrier timeout
                ctx.set_option("gemm_debug", dbg)
                try:
                    ms = ctx.bench_gemm(m, n, k, 1, geglu, bn, 1, iters=20)
                    line.append(f"{dbg:03b}:{ms * 1e3:6.1f}")
                except Exception as e:  # noqa
                    line.append(f"{dbg:03b}: fail")
            print(f"gemm {m}x{n}x{k} geglu={geglu} cg={cg} bn={bn}  " + "  ".join(line), flush=True)
ctx.set_option("gemm_debug", 8)
for (m, n, k, geglu) in shapes[:2]:
    for cg in (1, 2):
        ctx.set_option("gemm_cg", cg)
        for bn in (256, 128):
            ctx.bench_gemm(m, n, k, 1, geglu, bn, 1, iters=1)
            ctx.synchronize()
            print(f"^^ trace gemm {m}x{n}x{k} geglu={geglu} cg={cg} bn={bn}", flush=True)
ctx.set_option("gemm_debug", 0)
