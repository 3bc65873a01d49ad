"""GPU bring-up probe: peaks, op-level parity vs torch-CPU fp64, GEMM/conv timing sweeps.
Each section is meant to run in its own process under `timeout` (see tools/probe_all.sh)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def peaks():
    import torch
    res = {}
    dev = torch.device("cuda:0")
    n = 8192
    for name, dtype, tf32 in (("tf32", torch.float32, True), ("bf16", torch.bfloat16, False),
                              ("fp32", torch.float32, False)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        a = torch.randn(n, n, device=dev, dtype=dtype)
        b = torch.randn(n, n, device=dev, dtype=dtype)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[name + "_tflops_burst"] = 2 * n ** 3 / best / 1e9
        # sustained: back to back for ~3 s
        t0 = time.time()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        cnt = 0
        while time.time() - t0 < 3.0:
            for _ in range(10):
                a @ b
            cnt += 10
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        res[name + "_tflops_sustained"] = 2 * n ** 3 * cnt / e0.elapsed_time(e1) / 1e9
    # smaller tf32 shapes typical of the UNet
    torch.backends.cuda.matmul.allow_tf32 = True
    for (m, nn, k) in ((4096, 320, 2880), (4096, 2560, 320), (1024, 640, 5760), (256, 1280, 11520)):
        a = torch.randn(m, k, device=dev)
        b = torch.randn(k, nn, device=dev)
        for _ in range(3):
            a @ b
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        res[f"cublas_tf32_{m}x{nn}x{k}_tflops"] = 2 * m * nn * k * 20 / e0.elapsed_time(e1) / 1e9
    res["gpu"] = torch.cuda.get_device_name(0)
    print(json.dumps(res, indent=1))
    json.dump(res, open(os.path.join(OUT, "peaks.json"), "w"), indent=1)


def gemm():
    import torch
    from tsd_b200.api import Context
    ctx = Context(0)
    ctx.set_option("gemm_cg", int(os.environ.get("GEMM_CG", "0")))
    rng = np.random.default_rng(0)
    ok = True
    cases = [  # (rows, in, out, bias, force_bn, force_splits)
        (128, 32, 16, True, 0, 0), (128, 64, 64, True, 0, 0), (256, 320, 320, True, 0, 0),
        (200, 320, 960, False, 0, 0), (4096, 320, 320, True, 160, 0), (4096, 320, 320, True, 80, 0),
        (4096, 320, 320, True, 64, 0), (1024, 640, 640, True, 128, 2), (256, 1280, 1280, True, 0, 4),
        (77, 768, 320, False, 0, 0), (4096, 320, 2560, True, 256, 0), (333, 40, 48, True, 0, 0),
        (1024, 2560, 640, True, 0, 0), (64, 1280, 1280, True, 256, 8),
    ]
    for (rows, fin, fout, bias, fbn, fsp) in cases:
        x = rng.standard_normal((1, rows, fin), dtype=np.float32)
        w = (rng.standard_normal((fout, fin), dtype=np.float32) / np.sqrt(fin)).astype(np.float32)
        b = rng.standard_normal(fout, dtype=np.float32) if bias else None
        ctx.set_option("force_bn", fbn)
        ctx.set_option("force_splits", fsp)
        try:
            y = ctx.linear(x, w, b)
        except Exception as e:  # noqa
            print("linear", rows, fin, fout, "FAILED", e)
            ok = False
            continue
        ref = x.astype(np.float64) @ w.astype(np.float64).T + (0 if b is None else b.astype(np.float64))
        e = relerr(y, ref)
        flag = "ok" if e < 5e-3 else "BAD"
        ok &= e < 5e-3
        print(f"linear rows={rows} in={fin} out={fout} bias={bias} bn={fbn} splits={fsp}: rel_linf={e:.2e} {flag}")
    ctx.set_option("force_bn", 0)
    ctx.set_option("force_splits", 0)
    for (c, m, k, n) in ((8, 256, 40, 256), (8, 512, 80, 77), (2, 130, 36, 50), (1, 64, 512, 512)):
        a = rng.standard_normal((c, m, k), dtype=np.float32)
        bb = rng.standard_normal((c, k, n), dtype=np.float32)
        try:
            y = ctx.matmul(a, bb)
        except Exception as e:  # noqa
            print("matmul", c, m, k, n, "FAILED", e)
            ok = False
            continue
        ref = a.astype(np.float64) @ bb.astype(np.float64)
        e = relerr(y, ref)
        ok &= e < 5e-3
        print(f"matmul c={c} m={m} k={k} n={n}: rel_linf={e:.2e} {'ok' if e < 5e-3 else 'BAD'}")
    print("GEMM_PROBE", "PASS" if ok else "FAIL")


def conv():
    import torch
    import torch.nn.functional as F
    from tsd_b200.api import Context
    ctx = Context(0)
    ctx.set_option("gemm_cg", int(os.environ.get("GEMM_CG", "0")))
    rng = np.random.default_rng(1)
    ok = True
    cases = [  # n, cin, h, w, cout, k, pad, stride, bn, splits
        (1, 32, 16, 16, 32, 3, 1, 1, 0, 0), (1, 64, 8, 8, 64, 3, 1, 1, 0, 0),
        (2, 64, 16, 16, 96, 3, 1, 1, 0, 0), (1, 320, 64, 64, 320, 3, 1, 1, 160, 0),
        (1, 320, 64, 64, 320, 3, 1, 1, 160, 2), (1, 640, 32, 32, 640, 3, 1, 1, 0, 0),
        (1, 1280, 16, 16, 1280, 3, 1, 1, 0, 0), (1, 320, 64, 64, 320, 3, 1, 2, 0, 0),
        (1, 4, 64, 64, 320, 3, 1, 1, 0, 0), (1, 320, 64, 64, 4, 3, 1, 1, 0, 0),
        (1, 128, 256, 256, 3, 3, 1, 1, 0, 0), (1, 320, 64, 64, 320, 1, 0, 1, 0, 0),
        (1, 128, 40, 24, 128, 3, 1, 1, 0, 0), (1, 4, 8, 8, 4, 1, 0, 1, 0, 0),
        (1, 96, 20, 12, 48, 3, 1, 1, 0, 0), (1, 256, 128, 128, 256, 3, 1, 1, 0, 0),
    ]
    for (n, cin, h, w, cout, k, pad, stride, fbn, fsp) in cases:
        x = rng.standard_normal((n, cin, h, w), dtype=np.float32)
        wt = (rng.standard_normal((cout, cin, k, k), dtype=np.float32) / np.sqrt(cin * k * k)).astype(np.float32)
        b = rng.standard_normal(cout, dtype=np.float32)
        ctx.set_option("force_bn", fbn)
        ctx.set_option("force_splits", fsp)
        try:
            y = ctx.conv2d(x, wt, b, pad=pad, stride=stride)
        except Exception as e:  # noqa
            print("conv", (n, cin, h, w, cout, k, pad, stride), "FAILED", e)
            ok = False
            continue
        ref = F.conv2d(torch.from_numpy(x).double(), torch.from_numpy(wt).double(),
                       torch.from_numpy(b).double(), stride=stride, padding=pad).numpy()
        e = relerr(y, ref)
        ok &= e < 5e-3
        print(f"conv n={n} cin={cin} {h}x{w} cout={cout} k={k} p={pad} s={stride} bn={fbn} sp={fsp}: "
              f"rel_linf={e:.2e} {'ok' if e < 5e-3 else 'BAD'}")
    print("CONV_PROBE", "PASS" if ok else "FAIL")


def elementwise():
    import torch
    import torch.nn.functional as F
    from tsd_b200.api import Context
    ctx = Context(0)
    rng = np.random.default_rng(2)
    ok = True

    def chk(name, got, ref, tol):
        nonlocal ok
        e = relerr(got, ref)
        ok &= e < tol
        print(f"{name}: rel_linf={e:.2e} {'ok' if e < tol else 'BAD'}")

    for (c, h, w, g, eps) in ((320, 16, 16, 32, 1e-5), (64, 8, 8, 16, 1e-5), (320, 8, 8, 320, 1e-5),
                              (30, 5, 7, 3, 1e-6), (1920, 16, 16, 32, 1e-5), (128, 64, 64, 32, 1e-5)):
        x = (rng.standard_normal((c, h, w)) * 2 + 0.5).astype(np.float32)
        y = ctx.groupnorm(x, g, eps)
        xd = x.astype(np.float64).reshape(g, -1)
        ref = ((xd - xd.mean(1, keepdims=True)) / (xd.std(1, keepdims=True) + eps)).reshape(c, h, w)
        chk(f"groupnorm c={c} {h}x{w} g={g}", y, ref, 2e-5)
    x = rng.standard_normal((320, 256, 1)).astype(np.float32)
    y = ctx.layernorm(x)
    xd = x.astype(np.float64)
    chk("layernorm global", y, (xd - xd.mean()) / (xd.std() + 1e-5), 2e-5)
    x = (rng.standard_normal(100003) * 3).astype(np.float32)
    chk("silu", ctx.silu(x), x.astype(np.float64) / (1 + np.exp(-x.astype(np.float64))), 1e-6)
    chk("gelu", ctx.gelu(x), F.gelu(torch.from_numpy(x).double(), approximate="tanh").numpy(), 1e-6)
    x = rng.standard_normal((12, 5, 7)).astype(np.float32)
    chk("upsample2x c=12", ctx.upsample2x(x), x.repeat(2, 1).repeat(2, 2), 1e-7)
    x = rng.standard_normal((5, 4, 6)).astype(np.float32)
    chk("upsample2x c=5", ctx.upsample2x(x), x.repeat(2, 1).repeat(2, 2), 1e-7)
    x = rng.standard_normal((3, 50, 77)).astype(np.float32)
    e = np.exp(x.astype(np.float64))
    chk("softmax dim=2 (columns over rows)", ctx.softmax(x, 2), e / e.sum(1, keepdims=True), 1e-5)
    chk("softmax dim=1 (rows)", ctx.softmax(x, 1), e / e.sum(2, keepdims=True), 1e-5)
    lat = rng.standard_normal(4 * 64 * 64).astype(np.float32)
    ec = rng.standard_normal(lat.size).astype(np.float32)
    eu = rng.standard_normal(lat.size).astype(np.float32)
    nz = rng.standard_normal(lat.size).astype(np.float32)
    out = ctx.sampler_step(lat, ec, eu, 7.5, nz, 0.9, 0.43, 0.1, 0.88, 0.05)
    e_ = (ec.astype(np.float64) - eu) * 7.5 + eu
    x0 = (lat - e_ * 0.43) / 0.9
    chk("sampler_step", out, x0 * 0.1 + lat * 0.88 + nz * 0.05, 1e-5)
    # unfused attention path
    ctx.set_option("fused_attention", 0)
    for axis in (0, 1):
        ctx.set_option("softmax_axis", axis)
        for (hh, tq, tk, d) in ((8, 256, 256, 40), (8, 128, 77, 80), (2, 64, 64, 160)):
            q = rng.standard_normal((hh, tq, d)).astype(np.float32)
            k = rng.standard_normal((hh, tk, d)).astype(np.float32)
            v = rng.standard_normal((hh, tk, d)).astype(np.float32)
            o = ctx.attention_core(q, k, v)
            s = np.einsum("hid,hjd->hij", q.astype(np.float64), k.astype(np.float64)) / np.sqrt(d)
            p = np.exp(s - s.max())
            p = p / p.sum(1 if axis == 0 else 2, keepdims=True)
            ref = np.einsum("hij,hjd->ihd", p, v.astype(np.float64)).reshape(tq, hh * d)
            chk(f"attention(unfused) axis={axis} h={hh} tq={tq} tk={tk} d={d}", o, ref, 5e-3)
    print("ELEMENTWISE_PROBE", "PASS" if ok else "FAIL")


def timing():
    from tsd_b200.api import Context
    ctx = Context(0)
    rows = []

    def rec(kind, desc, flops, ms, **kw):
        tf = flops / ms / 1e9
        rows.append(dict(kind=kind, desc=desc, ms=ms, tflops=tf, **kw))
        print(f"{kind} {desc} {kw}: {ms * 1e3:.1f} us  {tf:.1f} TFLOP/s", flush=True)

    for (m, n, k) in ((8192, 8192, 8192), (4096, 4096, 4096)):
        for bn in (128, 256):
            ms = ctx.bench_gemm(m, n, k, force_bn=bn, iters=5)
            rec("gemm", f"{m}x{n}x{k}", 2.0 * m * n * k, ms, bn=bn, splits=1)
    for (n_, h, w, cin, cout) in ((1, 64, 64, 320, 320), (1, 32, 32, 640, 640), (1, 16, 16, 1280, 1280),
                                  (1, 16, 16, 2560, 1280), (1, 64, 64, 640, 320), (1, 128, 128, 512, 512),
                                  (1, 512, 512, 128, 128)):
        flops = 2.0 * n_ * h * w * cout * 9 * cin
        cands = [(0, 0)]
        if h * w <= 4096:
            bns = [b for b in (64, 80, 128, 160, 256) if ((cout + 15) // 16 * 16) % b == 0]
            cands += [(b, s) for b in bns for s in (1, 2, 4, 8)]
        for (bn, sp) in cands:
            try:
                ms = ctx.bench_conv(n_, h, w, cin, cout, 3, 1, bn, sp, iters=10)
                rec("conv3x3", f"{h}x{w} {cin}->{cout}", flops, ms, bn=bn, splits=sp)
            except Exception as e:  # noqa
                print("conv timing failed", (h, w, cin, cout, bn, sp), e)
    for (m, n, k, geglu) in ((4096, 320, 320, 0), (4096, 960, 320, 0), (4096, 2560, 320, 1), (4096, 320, 1280, 0),
                             (1024, 5120, 640, 1), (1024, 640, 2560, 0), (256, 10240, 1280, 1),
                             (256, 1280, 5120, 0), (256, 1280, 1280, 0)):
        for (bn, sp) in ((0, 0), (64, 1), (128, 1), (160, 1), (256, 1), (128, 2), (128, 4), (256, 4)):
            np_ = (n + 15) // 16 * 16
            if bn and (np_ % bn or (geglu and (n // 2) % (bn // 2))):
                continue
            if geglu and sp > 1:
                continue
            try:
                ms = ctx.bench_gemm(m, n, k, 1, geglu, bn, sp, iters=10)
                rec("gemm", f"{m}x{n}x{k} geglu={geglu}", 2.0 * m * n * k, ms, bn=bn, splits=sp)
            except Exception as e:  # noqa
                print("gemm timing failed", (m, n, k, bn, sp), e)
    json.dump(rows, open(os.path.join(OUT, "timing_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    {"peaks": peaks, "gemm": gemm, "conv": conv, "elementwise": elementwise, "timing": timing}[sys.argv[1]]()
