"""GEMM lab: where does the time of the small UNet GEMMs go?
  feed   - operand-feed experiments with the lab-only gemm_debug flags (skip MMA / A loads / B loads)
           and ring-depth overrides; results are timing only (outputs invalid).
  shapes - every distinct GEMM shape of one UNet step x (BN, splits) candidates.
  ncu    - a handful of representative launches, to be wrapped by `ncu --set full`.
Usage: python tools/gemm_lab.py feed|shapes|ncu"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

# (kind, args) : conv = (H, W, cin, cout) 3x3 ; gemm = (M, N, K, geglu)
UNET_SHAPES = [
    ("conv", (64, 64, 320, 320)), ("conv", (64, 64, 640, 320)), ("conv", (64, 64, 960, 320)),
    ("conv", (32, 32, 320, 640)), ("conv", (32, 32, 640, 640)), ("conv", (32, 32, 1280, 640)),
    ("conv", (32, 32, 960, 640)), ("conv", (16, 16, 640, 1280)), ("conv", (16, 16, 1280, 1280)),
    ("conv", (16, 16, 2560, 1280)), ("conv", (16, 16, 1920, 1280)),
    ("gemm", (4096, 320, 320, 0)), ("gemm", (4096, 960, 320, 0)), ("gemm", (4096, 2560, 320, 1)),
    ("gemm", (4096, 320, 1280, 0)), ("gemm", (1024, 640, 640, 0)), ("gemm", (1024, 1920, 640, 0)),
    ("gemm", (1024, 5120, 640, 1)), ("gemm", (1024, 640, 2560, 0)), ("gemm", (256, 1280, 1280, 0)),
    ("gemm", (256, 3840, 1280, 0)), ("gemm", (256, 10240, 1280, 1)), ("gemm", (256, 1280, 5120, 0)),
]


def run(ctx, kind, a, bn=0, sp=0, iters=20):
    if kind == "conv":
        h, w, cin, cout = a
        ms = ctx.bench_conv(1, h, w, cin, cout, 3, 1, bn, sp, iters=iters)
        fl = 2.0 * h * w * cout * 9 * cin
    else:
        m, n, k, geglu = a
        ms = ctx.bench_gemm(m, n, k, 1, geglu, bn, sp, iters=iters)
        fl = 2.0 * m * n * k
    return ms, fl / ms / 1e9


def feed():
    from tsd_b200.api import Context
    ctx = Context(0)
    rows = []
    cases = [("conv", (64, 64, 320, 320), 160, 1), ("conv", (64, 64, 320, 320), 80, 1),
             ("conv", (64, 64, 320, 320), 160, 2), ("conv", (32, 32, 640, 640), 160, 4),
             ("conv", (16, 16, 1280, 1280), 160, 8), ("gemm", (4096, 2560, 320, 1), 160, 1),
             ("gemm", (8192, 8192, 2048, 0), 256, 1)]
    for (kind, a, bn, sp) in cases:
        for stages in (0, 2, 3):
            for dbg in (0, 1, 1 | 2, 1 | 4, 1 | 2 | 4):
                if stages and dbg not in (0, 1):
                    continue
                ctx.set_option("force_stages", stages)
                ctx.set_option("gemm_debug", dbg)
                try:
                    ms, tf = run(ctx, kind, a, bn, sp)
                except Exception as e:  # noqa
                    print("FAILED", kind, a, bn, sp, stages, dbg, e, flush=True)
                    continue
                rows.append(dict(kind=kind, shape=a, bn=bn, splits=sp, stages=stages, debug=dbg, us=ms * 1e3, tflops=tf))
                print(f"{kind} {a} bn={bn} sp={sp} stages={stages or 'max'} dbg={dbg:03b}: {ms * 1e3:8.2f} us"
                      f"  ({tf:6.1f} TF/s equiv)", flush=True)
    ctx.set_option("force_stages", 0)
    ctx.set_option("gemm_debug", 0)
    json.dump(rows, open(os.path.join(OUT, "gemm_lab_feed.json"), "w"), indent=1)


def shapes():
    from tsd_b200.api import Context
    ctx = Context(0)
    rows = []
    for (kind, a) in UNET_SHAPES:
        n = a[3] if kind == "conv" else a[1]
        geglu = 0 if kind == "conv" else a[3]
        npad = (n + 15) // 16 * 16
        cands = [(0, 0)]
        for bn in (64, 80, 96, 128, 160, 192, 256):
            if npad % bn or (geglu and ((n // 2) % (bn // 2) or bn % 32)):
                continue
            for sp in ((1,) if geglu else (1, 2, 3, 4, 6, 8, 12, 16)):
                cands.append((bn, sp))
        best = None
        for (bn, sp) in cands:
            for cg in ((0,) if not bn else (1, 2)):
                ctx.set_option("gemm_cg", cg)
                try:
                    ms, tf = run(ctx, kind, a, bn, sp, iters=10)
                except Exception as e:  # noqa
                    print("FAILED", kind, a, bn, sp, cg, e, flush=True)
                    continue
                rows.append(dict(kind=kind, shape=a, bn=bn, splits=sp, cg=cg, us=ms * 1e3, tflops=tf))
                if bn and (best is None or ms < best[0]):
                    best = (ms, bn, sp, tf, cg)
                if not bn:
                    print(f"{kind} {a} heuristic: {ms * 1e3:8.2f} us {tf:6.1f} TF/s", flush=True)
        print(f"{kind} {a} best: bn={best[1]} sp={best[2]} cg={best[4]} {best[0] * 1e3:8.2f} us {best[3]:6.1f} TF/s", flush=True)
    ctx.set_option("gemm_cg", 0)
    json.dump(rows, open(os.path.join(OUT, "gemm_lab_shapes.json"), "w"), indent=1)


def trace():
    from tsd_b200.api import Context
    ctx = Context(0)
    for (kind, a, bn, sp) in (("conv", (64, 64, 320, 320), 160, 1), ("conv", (64, 64, 320, 320), 160, 2),
                              ("gemm", (4096, 320, 32, 0), 160, 1), ("gemm", (4096, 320, 320, 0), 160, 1),
                              ("gemm", (8192, 8192, 2048, 0), 256, 1)):
        for cg in (1, 2):
            ctx.set_option("gemm_cg", cg)
            for dbg in (8, 8 | 1 | 2 | 4):
                ctx.set_option("gemm_debug", dbg)
                ms, tf = run(ctx, kind, a, bn, sp, iters=2)
                ctx.synchronize()
                print(f"^^ {kind} {a} bn={bn} sp={sp} cg={cg} dbg={dbg:04b}: {ms * 1e3:.2f} us", flush=True)
    ctx.set_option("gemm_debug", 0)
    ctx.set_option("gemm_cg", 0)


def roles():
    """which single-thread role bounds the K loop?  (dbg bit0: no MMAs, bit1: no A loads, bit2: no B loads)"""
    from tsd_b200.api import Context
    ctx = Context(0)
    ctx.set_option("autotune", 0)
    for (kind, a, bn, sp) in (("conv", (64, 64, 320, 320), 80, 1), ("conv", (64, 64, 320, 320), 160, 1),
                              ("conv", (64, 64, 320, 320), 160, 2), ("gemm", (4096, 2560, 320, 1), 160, 1),
                              ("gemm", (256, 1280, 5120, 0), 160, 5)):
        for cg in (1, 2):
            ctx.set_option("gemm_cg", cg)
            line = []
            for dbg in (0, 1, 6, 7):
                ctx.set_option("gemm_debug", dbg)
                ms, tf = run(ctx, kind, a, bn, sp, iters=20)
                line.append(f"dbg={dbg:03b}:{ms * 1e3:6.1f}")
            print(f"{kind} {a} bn={bn} sp={sp} cg={cg}  " + "  ".join(line), flush=True)
    ctx.set_option("gemm_debug", 0)
    ctx.set_option("gemm_cg", 0)


def stats():
    """overhead of the producer-side norm statistics (epilogue / split-K reduce variants)"""
    from tsd_b200.api import Context
    ctx = Context(0)
    for (kind, a, bn, sp) in (("conv", (64, 64, 320, 320), 160, 1), ("conv", (64, 64, 320, 320), 160, 2),
                              ("conv", (32, 32, 640, 640), 128, 3), ("conv", (16, 16, 1280, 1280), 128, 6),
                              ("gemm", (4096, 320, 320, 0), 160, 1), ("gemm", (256, 1280, 1280, 0), 64, 1)):
        for g in (0, 32, 1):
            ctx.set_option("bench_stats_groups", g)
            ms, tf = run(ctx, kind, a, bn, sp, iters=20)
            print(f"{kind} {a} bn={bn} sp={sp} stats_groups={g}: {ms * 1e3:.2f} us", flush=True)
    ctx.set_option("bench_stats_groups", 0)
    ctx.set_option("gemm_debug", 8)
    for g in (0, 32):
        ctx.set_option("bench_stats_groups", g)
        run(ctx, "conv", (64, 64, 320, 320), 160, 1, iters=1)
        ctx.synchronize()
        print("^^ stats_groups", g, flush=True)
    ctx.set_option("gemm_debug", 0)
    ctx.set_option("bench_stats_groups", 0)


def ncu():
    from tsd_b200.api import Context
    ctx = Context(0)
    for (kind, a, bn, sp) in (("conv", (64, 64, 320, 320), 0, 0), ("conv", (32, 32, 640, 640), 0, 0),
                              ("conv", (16, 16, 1280, 1280), 0, 0), ("gemm", (4096, 2560, 320, 1), 0, 0)):
        ms, tf = run(ctx, kind, a, bn, sp, iters=1)
        print(kind, a, ms * 1e3, "us")


if __name__ == "__main__":
    {"feed": feed, "shapes": shapes, "ncu": ncu, "trace": trace, "stats": stats, "roles": roles}[sys.argv[1]]()
