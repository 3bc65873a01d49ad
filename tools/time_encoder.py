"""Times Encoder.forward (vae.mojo:131-159) at the 512x512 image of BASELINE config 2 (host in/out and
device-resident) and prints the per-family device time of one pass."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context, Encoder  # noqa: E402

ctx = Context(0)
m = Encoder(ctx, 64, 64)
m.init_random(1236)
rng = np.random.default_rng(0)
img = rng.uniform(0, 255, (3, 512, 512)).astype(np.float32)
noise = rng.standard_normal((4, 64, 64), dtype=np.float32)
z = m.forward(img, noise, rescale=True)
for _ in range(3):
    m.forward(img, noise, rescale=True)
t0 = time.perf_counter()
n = 10
for _ in range(n):
    z = m.forward(img, noise, rescale=True)
dt = (time.perf_counter() - t0) / n
print(f"encoder 512x512 -> 64x64x4: {dt * 1e3:.2f} ms per forward (host in/out), params {m.num_params()}, "
      f"finite {bool(np.isfinite(z).all())}, launches/forward {ctx.launch_count()}")
