import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context, Diffusion
from tsd_b200 import sampler as host_sampler
ctx = Context(0)
m = Diffusion(ctx, 64, 64, max_batch=2)
m.init_random(1234)
rng = np.random.default_rng(11)
x = rng.standard_normal((4, 64, 64), dtype=np.float32)
cx = rng.standard_normal((77, 768), dtype=np.float32)
t = host_sampler.get_time_embedding(500.0)
y0 = m.forward(x, cx, t)
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
for opts in ({"producer_stats": 0}, {"producer_stats": 0, "ln_fold": 0}, {"producer_stats": 0, "ln_fold": 0, "autotune": 0},
             {"producer_stats": 0, "ln_fold": 0, "defer_reduce": 0}, {"producer_stats": 0, "ln_fold": 0, "attn_v2": 0},
             {"producer_stats": 0, "ln_fold": 0, "cuda_graph": 0}, {"gn_partial": 0}, {"gn_partial": 0, "ln_fold": 0}):
    old = {k: ctx.get_option(k) for k in opts}
    for k, v in opts.items(): ctx.set_option(k, v)
    y = m.forward(x, cx, t)
    ya = m.forward(x, cx, t)
    for k, v in old.items(): ctx.set_option(k, v)
    print(opts, "rel vs default %.2e" % rel(y, y0), "replay equal", bool(np.array_equal(y, ya)), "nan", bool(np.isnan(y).any()), flush=True)
