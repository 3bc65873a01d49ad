"""Lab: the small-M linears of the 16x16 / 8x8 / 32x32 UNet levels (no norm consumer): narrow unsplit tiles (the shipped
plans) vs fat tiles + split-K with the three reduction modes (0 reduce kernel, 1 cluster through L2, 2 cluster through DSMEM)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context  # noqa: E402
ctx = Context(0)
ctx.set_option("autotune", 0)
ctx.set_option("splitk_cluster_max", 16)
shapes = [(256, 1280, 1280), (64, 1280, 1280), (1024, 640, 640), (256, 1280, 5120), (1024, 640, 2560)]
plans = [(32, 1, 2), (64, 1, 2), (64, 1, 1), (64, 2, 1), (64, 4, 1), (128, 4, 1), (160, 4, 1), (160, 8, 1), (160, 4, 2), (80, 4, 2), (128, 8, 1), (64, 8, 1), (256, 8, 1)]
for m, n, k in shapes:
    for bn, sp, cg in plans:
        if n % bn:
            continue
        ctx.set_option("gemm_cg", cg)
        res = []
        for cl in ((0,) if sp == 1 else (0, 1, 2)):
            ctx.set_option("splitk_cluster", cl)
            try:
                ms = ctx.bench_gemm(m, n, k, 1, 0, bn, sp, iters=30)
                res.append(f"mode{cl} {ms * 1e3:6.1f}")
            except Exception as e:  # plan the launcher rejects
                res.append(f"mode{cl}  fail")
        print(f"gemm {m}x{n}x{k} bn={bn:3d} splits={sp} cg={cg}: " + " | ".join(res) + " us", flush=True)
