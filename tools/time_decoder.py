"""Decoder.forward (vae.mojo:221-250) at the 64x64x4 latent of BASELINE config 2 / 4: ms per image, device-resident
batches of 1 and 4 (config 4 asks for 16; 4 keeps the probe short)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context, Decoder  # noqa: E402

ctx = Context(0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = Decoder(ctx, 64, 64, max_batch=nb)
m.init_random(1235)
z = (np.random.default_rng(0).standard_normal((nb, 4, 64, 64)) * 0.18215).astype(np.float32)
for _ in range(3):
    img = m.forward(z, rescale=True)
t0 = time.perf_counter()
n = 5
for _ in range(n):
    img = m.forward(z, rescale=True)
dt = (time.perf_counter() - t0) / n
print(f"decoder batch {nb}: {dt * 1e3:.2f} ms per call (host in/out), {dt * 1e3 / nb:.2f} ms per image, "
      f"{2514.52 * nb / dt / 1e3:.0f} TFLOP/s, finite {bool(np.isfinite(img).all())}")
