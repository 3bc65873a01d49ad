"""Launches each hot kernel at its BASELINE shape once (for `ncu --set full -k regex:...`); `what` selects the group.
  gemm     conv 320->320 @64x64 through tsd_bench_conv (the autotuner may try several tile plans first)
  attn     attention core h=8 T=4096 d=40: attn2_kernel statistics + apply, then the Tk=77 cross-attention (attn_kernel)
  step     two eager UNet steps; the profiler range (cudaProfilerStart/Stop, `ncu --profile-from-start off`) covers the
           SECOND, so launch 0 of a kernel family is its first launch of a warmed-up step with the shipped tile plans:
           gemm_tf32_kernel #0 = ResBlock conv1 320->320 3x3 @64x64; norm_cluster / norm_apply_partial / gemv_multi by -k
  decoder  conv 256->256 @512x512 (VAE decoder l20 shape)"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context, Diffusion  # noqa: E402

what = sys.argv[1]
ctx = Context(0)
if what == "gemm":
    ctx.bench_conv(1, 64, 64, 320, 320, 3, 1, iters=2)
elif what == "decoder":
    ctx.bench_conv(1, 512, 512, 256, 256, 3, 1, iters=2)
elif what == "attn":
    rng = np.random.default_rng(0)
    for tk in (4096, 77):
        q = rng.standard_normal((8, 4096, 40), dtype=np.float32)
        k = rng.standard_normal((8, tk, 40), dtype=np.float32)
        v = rng.standard_normal((8, tk, 40), dtype=np.float32)
        ctx.attention_core(q, k, v)
else:
    ctx.set_option("cuda_graph", 0)
    m = Diffusion(ctx, 64, 64, max_batch=1)
    m.init_random(1234)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 64, 64), dtype=np.float32)
    cx = rng.standard_normal((77, 768), dtype=np.float32)
    t = np.concatenate([np.ones(160, np.float32), np.zeros(160, np.float32)])
    m.forward(x, cx, t)
    ctx.synchronize()
    import torch
    rt = torch.cuda.cudart()
    rt.cudaProfilerStart()
    m.forward(x, cx, t)
    ctx.synchronize()
    rt.cudaProfilerStop()
ctx.synchronize()
