"""Launches the norm kernels once at a UNet shape (for `ncu --set full -k regex:norm`)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context
ctx = Context(0)
x = np.random.default_rng(0).standard_normal((320, 64, 64), dtype=np.float32)
ctx.groupnorm(x, 32, 1e-5)
ctx.set_option("norm_v2", 1)
ctx.groupnorm(x, 32, 1e-5)
