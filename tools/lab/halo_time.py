"""halo vs implicit-GEMM conv timing on the UNet shapes"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context
ctx = Context(0)
ctx.set_option("autotune", 0)
ctx.set_option("halo_min_w", 8); ctx.set_option("halo_min_h", 8)
shapes = [(64, 64, 320, 320), (64, 64, 640, 320), (64, 64, 960, 320), (32, 32, 320, 640), (32, 32, 640, 640), (32, 32, 1280, 640),
          (16, 16, 1280, 1280), (16, 16, 2560, 1280)]
for (h, w, cin, cout) in shapes:
    fl = 2.0 * h * w * cout * 9 * cin
    res = {}
    for halo in (0, 2):
        ctx.set_option("conv_halo", halo)
        best = None
        for cg in (1, 2):
            ctx.set_option("gemm_cg", cg)
            for bn in (64, 80, 128, 160, 256):
                if cout % bn: continue
                for sp in (1, 2, 3, 4, 5, 6, 8):
                    try:
                        ms = ctx.bench_conv(1, h, w, cin, cout, 3, 1, bn, sp, iters=10)
                    except Exception as e:
                        continue
                    if best is None or ms < best[0]: best = (ms, bn, sp, cg)
        res[halo] = best
    for halo in (0, 2):
        ms, bn, sp, cg = res[halo]
        print(f"conv {h}x{w} {cin}->{cout} halo={halo}: best {ms*1e3:6.1f} us {fl/ms/1e9:6.1f} TF/s (bn={bn} sp={sp} cg={cg})", flush=True)
