"""In-kernel cycle trace of the GEMM epilogue (gemm_debug bit 3) for GEGLU and plain tiles.
Needs a lab build of the library: TSD_LAB_TRACE=1 bash stable-diffusion.mojo_b200/csrc/build.sh (touch gemm_tcgen05.cu first)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200.api import Context  # noqa: E402

ctx = Context(0)
ctx.set_option("autotune", 0)
ctx.set_option("gemm_cg", 1)
for (m, n, k, geglu, bn) in ((4096, 2560, 320, 1, 256), (4096, 2560, 320, 1, 128), (4096, 2560, 320, 0, 256),
                             (4096, 1280, 320, 0, 160), (4096, 320, 320, 0, 160)):
    ctx.set_option("gemm_debug", 0)
    ms = ctx.bench_gemm(m, n, k, 1, geglu, bn, 1, iters=20)
    ctx.set_option("gemm_debug", 32)
    ms_ns = ctx.bench_gemm(m, n, k, 1, geglu, bn, 1, iters=20)
    for dbg in (8, 8 | 32):
        ctx.set_option("gemm_debug", dbg)
        ctx.bench_gemm(m, n, k, 1, geglu, bn, 1, iters=1)
        ctx.synchronize()
        print(f"-- trace block above: debug={dbg}", flush=True)
    print(f"^^ gemm {m}x{n}x{k} geglu={geglu} bn={bn}: {ms * 1e3:.1f} us, without stores {ms_ns * 1e3:.1f} us", flush=True)
