"""Lab: TSD_NORM_TRACE=1 python tools/lab/norm2_trace.py - phase timings of the fused norm kernels inside one eager UNet step."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from tsd_b200.api import Context, Diffusion  # noqa: E402
ctx = Context(0)
ctx.set_option("cuda_graph", 0)
m = Diffusion(ctx, 64, 64, max_batch=1)
m.init_random(1234)
rng = np.random.default_rng(0)
x = rng.standard_normal((4, 64, 64), dtype=np.float32)
cx = rng.standard_normal((77, 768), dtype=np.float32)
t = np.concatenate([np.ones(160, np.float32), np.zeros(160, np.float32)])
y = m.forward(x, cx, t)
y = m.forward(x, cx, t)
print("== traced pass", flush=True)
