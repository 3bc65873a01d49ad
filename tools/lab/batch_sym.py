"""Are the two images of a batch of identical inputs bit-identical?  (lab)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context, Diffusion  # noqa: E402

ctx = Context(0)
m = Diffusion(ctx, 64, 64, max_batch=2)
m.init_random(1234)
rng = np.random.default_rng(0)
x = rng.standard_normal((4, 64, 64), dtype=np.float32)
cx = rng.standard_normal((77, 768), dtype=np.float32)
t = rng.standard_normal(320).astype(np.float32)
for opts in ({}, {"autotune": 0}, {"virtual_concat": 0}, {"defer_reduce": 0, "virtual_concat": 0}):
    for k, v in opts.items():
        ctx.set_option(k, v)
    ys = [m.forward(np.stack([x, x]), cx, t) for _ in range(3)]
    d01 = [float(np.abs(y[0] - y[1]).max()) for y in ys]
    rr = float(np.abs(ys[0] - ys[2]).max())
    print(opts, "max|img0-img1| per run:", d01, " run0 vs run2:", rr, " scale", float(np.abs(ys[0]).max()), flush=True)
    for k in opts:
        ctx.set_option(k, 1 if k != "norm_v2" else 0)
y1 = m.forward(x, cx, t)
y = m.forward(np.stack([x, x]), cx, t)
print("img0 vs single", float(np.abs(y[0] - y1).max()), "img1 vs single", float(np.abs(y[1] - y1).max()), flush=True)
