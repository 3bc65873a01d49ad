"""One GEMM shape launched a few times (for an ncu capture): python tools/lab/one_gemm.py M N K geglu bn cg"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context  # noqa: E402

m, n, k, geglu, bn, cg = (int(v) for v in sys.argv[1:7])
ctx = Context(0)
ctx.set_option("autotune", 0)
ctx.set_option("gemm_cg", cg)
print(ctx.bench_gemm(m, n, k, 1, geglu, bn, 1, iters=6) * 1e3, "us")
