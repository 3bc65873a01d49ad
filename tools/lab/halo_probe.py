"""halo conv kernel bring-up: parity of forced-halo convolutions vs torch fp64, both descriptor variants"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
import torch, torch.nn.functional as F
from tsd_b200.api import Context
variant = int(sys.argv[1]); minhw = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = Context(0)
ctx.set_option("autotune", 0); ctx.set_option("conv_halo", 2); ctx.set_option("gemm_debug", 16 if variant == 1 else 0)
if minhw: ctx.set_option("halo_min_w", minhw); ctx.set_option("halo_min_h", minhw)
rng = np.random.default_rng(3)
cases = [(1, 32, 32, 32, 32, 0, 0, 1), (1, 64, 32, 32, 64, 0, 0, 1), (1, 320, 64, 64, 320, 160, 1, 1), (1, 320, 64, 64, 320, 80, 1, 2),
         (1, 320, 64, 64, 320, 160, 2, 2), (1, 640, 32, 32, 640, 128, 2, 2), (2, 96, 40, 24, 48, 0, 0, 0), (1, 36, 20, 28, 80, 0, 0, 0)]
if minhw: cases += [(1, 64, 16, 16, 64, 0, 0, 1), (1, 1280, 16, 16, 1280, 128, 4, 2), (2, 64, 8, 8, 32, 0, 0, 1)]
for (n, cin, h, w, cout, bn, sp, cg) in cases:
    x = rng.standard_normal((n, cin, h, w), dtype=np.float32)
    wt = (rng.standard_normal((cout, cin, 3, 3), dtype=np.float32) / np.sqrt(cin * 9)).astype(np.float32)
    b = rng.standard_normal(cout, dtype=np.float32)
    ctx.set_option("force_bn", bn); ctx.set_option("force_splits", sp); ctx.set_option("gemm_cg", cg)
    try:
        y = ctx.conv2d(x, wt, b, pad=1, stride=1)
    except Exception as e:
        print("FAILED", (n, cin, h, w, cout, bn, sp, cg), e, flush=True); continue
    ref = F.conv2d(torch.from_numpy(x).double(), torch.from_numpy(wt).double(), torch.from_numpy(b).double(), padding=1).numpy()
    e = float(np.abs(y - ref).max() / np.abs(ref).max())
    print(f"variant={variant} n={n} cin={cin} {h}x{w} cout={cout} bn={bn} sp={sp} cg={cg}: rel_linf={e:.2e} {'ok' if e < 5e-3 else 'BAD'}", flush=True)
