"""GPU parity, op level: every op-granular entry point of include/tsd_b200.h called through the
C ABI (host buffers in the reference's layouts) against the fp64 oracle on the same seeded
inputs.  Tolerances (max|a-b|/max|b|):
  TF32 tensor-core ops (conv, linear, matmul, attention): 5e-3   (10-bit mantissa products,
       fp32 accumulate; measured ~6e-4 .. 2e-3)
  CUDA-core fp32 ops (norms, activations, softmax, sampler, degenerate-channel convs): 2e-5
  pure data movement (upsample): exact."""
import numpy as np
import pytest

import tsd_oracle as O
from conftest import relerr
from tsd_b200._lib import TsdError

pytestmark = pytest.mark.gpu

TOL_TF32 = 5e-3
TOL_FP32 = 2e-5


@pytest.fixture(scope="module")
def ops():
    return O.Ops("np", np.float64)


@pytest.mark.parametrize("n,cin,h,w,cout,k,pad,stride", [
    (1, 32, 16, 16, 32, 3, 1, 1),      # smallest tensor-core conv
    (2, 64, 16, 16, 96, 3, 1, 1),      # batch of images
    (1, 320, 32, 32, 320, 3, 1, 1),    # UNet layer-2 shape at a 32x32 latent
    (1, 320, 16, 16, 640, 3, 1, 1),    # channel growth
    (1, 320, 32, 32, 320, 3, 1, 2),    # stride-2 downsample (diffusion.mojo:180)
    (1, 4, 32, 32, 320, 3, 1, 1),      # conv_in: Cin=4 -> CUDA-core direct conv
    (1, 320, 32, 32, 4, 3, 1, 1),      # output layer: Cout=4
    (1, 128, 24, 40, 3, 3, 1, 1),      # VAE l26: Cout=3, ragged W
    (1, 320, 16, 16, 320, 1, 0, 1),    # 1x1 conv
    (1, 96, 12, 20, 48, 3, 1, 1),      # ragged tile edges (W not a divisor of 128)
    (1, 4, 8, 8, 4, 1, 0, 1),          # VAE l1
    (3, 36, 5, 7, 20, 3, 1, 1),        # odd everything
    (1, 320, 64, 64, 320, 3, 1, 1),    # the most frequent UNet convolution at the BASELINE latent (shipped tile plan)
    (1, 960, 64, 64, 320, 3, 1, 1),    # up-block conv1 after the skip concat at 64x64
    (2, 640, 32, 32, 640, 3, 1, 1),    # CFG batch of two at the 32x32 level
    (1, 256, 128, 128, 256, 3, 1, 1),  # VAE decoder level (multi-wave grid)
])
def test_conv2d(ctx, ops, n, cin, h, w, cout, k, pad, stride):
    rng = np.random.default_rng(cin * 1000 + cout + h)
    x = rng.standard_normal((n, cin, h, w), dtype=np.float32)
    wt = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32)
    b = rng.standard_normal(cout, dtype=np.float32)
    y = ctx.conv2d(x, wt, b, pad=pad, stride=stride)
    ref = np.stack([ops.conv2d(x[i], wt, b, pad, stride) for i in range(n)])
    assert y.shape == ref.shape
    tol = TOL_FP32 if cin < 32 else TOL_TF32
    assert relerr(y, ref) < tol


@pytest.mark.parametrize("n,cin,h,w,cout,stride", [
    (1, 128, 32, 32, 128, 2),     # Encoder l4 geometry (vae.mojo:97): two_stride_pad + stride-2, tensor-core path
    (2, 256, 16, 16, 256, 2),     # l7, batched
    (1, 64, 9, 13, 48, 2),        # odd sizes: the last output row/column reads the zero row/column
    (1, 8, 10, 10, 8, 2),         # CUDA-core direct path
    (1, 64, 12, 12, 32, 1),       # stride 1 with one-sided padding
])
def test_conv2d_bottom_right_padding(ctx, ops, n, cin, h, w, cout, stride):
    """Matrix.pad((0,1),(0,1)) + unpadded Conv2D (Encoder.two_stride_pad, vae.mojo:115-116) in one call."""
    rng = np.random.default_rng(cin + h)
    x = rng.standard_normal((n, cin, h, w), dtype=np.float32)
    wt = (rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(cin * 9)).astype(np.float32)
    b = rng.standard_normal(cout, dtype=np.float32)
    y = ctx.conv2d(x, wt, b, pad=0, stride=stride, pad_hi=1)
    ref = np.stack([ops.conv2d(x[i], wt, b, pad=0, stride=stride, pad_hi=1) for i in range(n)])
    assert y.shape == ref.shape == (n, cout, (h + 1 - 3) // stride + 1, (w + 1 - 3) // stride + 1)
    assert relerr(y, ref) < (TOL_FP32 if cin < 32 else TOL_TF32)
    # the explicit pad followed by the plain conv gives the same result
    xp = np.pad(x, ((0, 0), (0, 0), (0, 1), (0, 1)))
    assert relerr(ctx.conv2d(xp, wt, b, pad=0, stride=stride), ref) < (TOL_FP32 if cin < 32 else TOL_TF32)


def test_conv2d_stride2_both_paths(ctx, ops):
    """stride-2 3x3 convolution: the strided tensor-map path and the im2col path against the oracle"""
    rng = np.random.default_rng(5)
    for (n, cin, h, w, cout, pad, pad_hi) in ((1, 64, 16, 16, 64, 1, None), (2, 96, 12, 20, 32, 1, None),
                                              (1, 64, 10, 14, 48, 0, 1), (1, 128, 64, 64, 128, 0, 1)):
        x = rng.standard_normal((n, cin, h, w), dtype=np.float32)
        wt = (rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(cin * 9)).astype(np.float32)
        b = rng.standard_normal(cout, dtype=np.float32)
        ref = np.stack([ops.conv2d(x[i], wt, b, pad=pad, stride=2, pad_hi=pad_hi) for i in range(n)])
        for mode in (1, 0):
            with _Options(ctx, conv_stride_tma=mode):
                y = ctx.conv2d(x, wt, b, pad=pad, stride=2, pad_hi=pad_hi)
            assert y.shape == ref.shape and relerr(y, ref) < TOL_TF32, (mode, n, cin, h, w)


def test_conv2d_no_bias_and_errors(ctx, ops):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((64, 8, 8), dtype=np.float32)
    wt = (rng.standard_normal((64, 64, 3, 3)) / 24).astype(np.float32)
    assert relerr(ctx.conv2d(x, wt, None, pad=1), ops.conv2d(x, wt, None, 1)) < TOL_TF32
    with pytest.raises(TsdError):   # kernel larger than the padded input -> status, never a crash
        ctx.conv2d(x[:, :2, :2], wt, None, pad=0)
    with pytest.raises(TsdError):
        ctx.conv2d(x, wt[:, :32], None, pad=1)


@pytest.mark.parametrize("rows,fin,fout,bias", [
    (128, 32, 16, True), (256, 320, 320, True), (200, 320, 960, False), (1024, 640, 5120, True),
    (77, 768, 320, False), (333, 40, 48, True), (1, 320, 1280, True), (3, 1280, 640, True), (50, 30, 18, True),
])
def test_linear(ctx, ops, rows, fin, fout, bias):
    rng = np.random.default_rng(rows + fin)
    x = rng.standard_normal((1, rows, fin), dtype=np.float32)
    w = (rng.standard_normal((fout, fin)) / np.sqrt(fin)).astype(np.float32)
    b = rng.standard_normal(fout, dtype=np.float32) if bias else None
    y = ctx.linear(x, w, b)
    ref = ops.linear(x[0], w, b)[None]
    tol = TOL_FP32 if (rows <= 4 or fin % 4 or fout % 4) else TOL_TF32   # M<=4 / ragged: CUDA-core GEMV
    assert relerr(y, ref) < tol
    with pytest.raises(TsdError):
        ctx.linear(x, w[:, :-1], b)


@pytest.mark.parametrize("c,m,k,n", [(8, 256, 40, 256), (8, 512, 80, 77), (2, 130, 36, 50), (1, 64, 512, 512)])
def test_matmul(ctx, ops, c, m, k, n):
    rng = np.random.default_rng(m + n)
    a = rng.standard_normal((c, m, k), dtype=np.float32)
    b = rng.standard_normal((c, k, n), dtype=np.float32)
    assert relerr(ctx.matmul(a, b), ops.matmul(a, b)) < TOL_TF32
    with pytest.raises(TsdError):
        ctx.matmul(a, b[:, :-1])


@pytest.mark.parametrize("c,h,w,g,eps", [(320, 16, 16, 32, 1e-5), (64, 8, 8, 16, 1e-5), (320, 8, 8, 320, 1e-5),
                                         (30, 5, 7, 3, 1e-6), (1920, 16, 16, 32, 1e-5), (128, 64, 64, 32, 1e-5)])
def test_groupnorm(ctx, ops, c, h, w, g, eps):
    rng = np.random.default_rng(c + g)
    x = (rng.standard_normal((c, h, w)) * 2 + 0.5).astype(np.float32)
    assert relerr(ctx.groupnorm(x, g, eps), ops.group_norm(x, g, eps)) < TOL_FP32
    with pytest.raises(TsdError):   # reference: "does not evenly divide" -> null Matrix (utils.mojo:1851-1853)
        ctx.groupnorm(x, g + 1 if c % (g + 1) else 7, eps)


def test_layernorm_both_modes(ctx, ops):
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((320, 256, 1)) * 1.5 - 0.3).astype(np.float32)   # reference (C,T,1) layout
    want = ops.layer_norm(x[:, :, 0].T).T[:, :, None]
    assert relerr(ctx.layernorm(x), want) < TOL_FP32
    ctx.set_option("layernorm_mode", 1)
    try:
        tok = O.Ops("np", np.float64, O.Switches(layernorm="token")).layer_norm(x[:, :, 0].T).T[:, :, None]
        assert relerr(ctx.layernorm(x), tok) < TOL_FP32
    finally:
        ctx.set_option("layernorm_mode", 0)


def test_activations_and_upsample(ctx, ops):
    rng = np.random.default_rng(6)
    x = (rng.standard_normal(100003) * 3).astype(np.float32)
    assert relerr(ctx.silu(x), ops.silu(x)) < TOL_FP32
    assert relerr(ctx.gelu(x), ops.gelu(x)) < TOL_FP32
    for c in (12, 5):
        img = rng.standard_normal((c, 5, 7)).astype(np.float32)
        assert np.array_equal(ctx.upsample2x(img), ops.upsample2x(img))


def test_softmax_dims(ctx, ops):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3, 50, 77)).astype(np.float32)
    assert relerr(ctx.softmax(x, 2), ops.softmax(x, 2)) < TOL_FP32
    assert relerr(ctx.softmax(x, 1), ops.softmax(x, 1)) < TOL_FP32
    with pytest.raises(TsdError):   # reference: "Invalid dimension for softmax" (utils.mojo:446-448)
        ctx.softmax(x, 3)


@pytest.mark.parametrize("axis", ["query", "key"])
@pytest.mark.parametrize("h,tq,tk,d", [(8, 256, 256, 40), (8, 128, 77, 80), (2, 64, 64, 160), (1, 256, 256, 512),
                                       (8, 1024, 1024, 80), (8, 200, 77, 40)])
def test_attention_core(ctx, h, tq, tk, d, axis):
    rng = np.random.default_rng(h * tq + d)
    q = rng.standard_normal((h, tq, d), dtype=np.float32)
    k = rng.standard_normal((h, tk, d), dtype=np.float32)
    v = rng.standard_normal((h, tk, d), dtype=np.float32)
    ops = O.Ops("np", np.float64, O.Switches(softmax_axis=axis))
    ctx.set_option("softmax_axis", 0 if axis == "query" else 1)
    try:
        got = ctx.attention_core(q, k, v)
    finally:
        ctx.set_option("softmax_axis", 0)
    assert relerr(got, ops.attention_core(q, k, v)) < TOL_TF32


@pytest.mark.parametrize("axis", ["query", "key"])
@pytest.mark.parametrize("tk", [4096, 77])
@pytest.mark.parametrize("d", [40, 80, 160])
def test_attention_core_sweep_shapes(ctx, d, tk, axis):
    """BASELINE configs[4] shapes: h = 8, Tq = 4096, Tk in {4096, 77}, d in {40, 80, 160}, both softmax axes, against
    the fp64 oracle (helpers/attention.mojo:46-62); the T = 4096 self-attention cases run on the software-pipelined
    attn2_kernel, which is also compared with the first-generation kernel."""
    h, tq = 8, 4096
    rng = np.random.default_rng(d * 7 + tk)
    q = rng.standard_normal((h, tq, d), dtype=np.float32)
    k = rng.standard_normal((h, tk, d), dtype=np.float32)
    v = rng.standard_normal((h, tk, d), dtype=np.float32)
    ops = O.Ops("np", np.float64, O.Switches(softmax_axis=axis))
    ref = ops.attention_core(q, k, v)
    ctx.set_option("softmax_axis", 0 if axis == "query" else 1)
    try:
        got = ctx.attention_core(q, k, v)
        ctx.set_option("attn_v2", 0)
        got_v1 = ctx.attention_core(q, k, v)
    finally:
        ctx.set_option("attn_v2", 1)
        ctx.set_option("softmax_axis", 0)
    e, e1 = relerr(got, ref), relerr(got_v1, ref)
    print(f"attention core h=8 tq=4096 tk={tk} d={d} axis={axis}: rel_linf {e:.2e} (v1 kernel {e1:.2e})")
    assert e < TOL_TF32 and e1 < TOL_TF32
    assert relerr(got, got_v1) < 1e-3    # same arithmetic in another summation order: P may flip a TF32 ulp (2^-11)


def test_attention_core_dev_entry(ctx):
    """tsd_attention_core_dev (device pointers, asynchronous) and tsd_bench_attention."""
    import ctypes as C
    import torch
    rng = np.random.default_rng(3)
    h, t, d = 8, 512, 40
    q, k, v = (rng.standard_normal((h, t, d), dtype=np.float32) for _ in range(3))
    dq, dk, dv = (torch.from_numpy(a).cuda() for a in (q, k, v))
    out = torch.empty((t, h * d), device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    ctx._ck(ctx.L.tsd_attention_core_dev(ctx.h, dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), h, t, t, d, out.data_ptr()))
    ctx.synchronize()
    assert relerr(out.cpu().numpy(), ctx.attention_core(q, k, v)) == 0.0
    ms = C.c_double()
    ctx._ck(ctx.L.tsd_bench_attention(ctx.h, h, t, t, d, 3, C.byref(ms)))
    assert 0.0 < ms.value < 50.0


def test_self_and_cross_attention(ctx, ops):
    rng = np.random.default_rng(8)
    t, c, hd = 256, 320, 8
    x = rng.standard_normal((1, t, c), dtype=np.float32)
    mk = lambda o, i: (rng.standard_normal((o, i)) / np.sqrt(i)).astype(np.float32)  # noqa: E731
    w_in, w_out, b_out = mk(3 * c, c), mk(c, c), rng.standard_normal(c, dtype=np.float32)
    got = ctx.self_attention(x, hd, w_in, None, w_out, b_out)
    assert relerr(got[0], ops.self_attention(x[0], hd, w_in, None, w_out, b_out)) < TOL_TF32
    b_in = rng.standard_normal(3 * c, dtype=np.float32)   # VAE flavour: in_proj bias on
    got = ctx.self_attention(x, 1, w_in, b_in, w_out, b_out)
    assert relerr(got[0], ops.self_attention(x[0], 1, w_in, b_in, w_out, b_out)) < TOL_TF32
    cx = rng.standard_normal((1, 77, 768), dtype=np.float32)
    wq, wk, wv, wo = mk(c, c), mk(c, 768), mk(c, 768), mk(c, c)
    got = ctx.cross_attention(x, cx, hd, wq, None, wk, None, wv, None, wo, b_out)
    assert relerr(got[0], ops.cross_attention(x[0], cx[0], hd, wq, None, wk, None, wv, None, wo, b_out)) < TOL_TF32


def test_sampler_step(ctx):
    rng = np.random.default_rng(9)
    n = 4 * 64 * 64
    lat, ec, eu, nz = (rng.standard_normal(n).astype(np.float32) for _ in range(4))
    sm = O.DDPMSampler()
    sm.set_inference_timesteps(20)
    for t in (950, 500, 0):
        co = sm.coefficients(t)
        eps = O.cfg_combine(ec.astype(np.float64), eu, 7.5)
        want = sm.step(t, lat.astype(np.float64), eps, nz)
        got = ctx.sampler_step(lat, ec, eu, 7.5, nz if t > 0 else None, *[float(v) for v in co])
        assert relerr(got, want) < TOL_FP32
    got = ctx.sampler_step(lat, ec, None, 1.0, None, *[float(v) for v in sm.coefficients(500)])
    assert relerr(got, sm.step(500, lat.astype(np.float64), ec.astype(np.float64), None)) < TOL_FP32


def test_sampler_add_noise(ctx):
    """DDPMSampler.add_noise, sampler.mojo:111-124."""
    rng = np.random.default_rng(10)
    x, nz = (rng.standard_normal((4, 16, 16)).astype(np.float32) for _ in range(2))
    sm = O.DDPMSampler()
    for t in (999, 400, 0):
        ab = sm.alphas_cumprod[t]
        got = ctx.sampler_add_noise(x, nz, np.sqrt(ab), np.sqrt(1 - ab))
        assert got.shape == x.shape and relerr(got, sm.add_noise(x.astype(np.float64), t, nz)) < TOL_FP32
    with pytest.raises(TsdError):
        ctx.sampler_add_noise(x, nz[:2], 1.0, 0.0)


# ------------------------------------------------------------------------------------------------
# execution variants of the tensor-core GEMM / convolution: every variant must meet the same
# tolerance against the oracle (the autotuner may pick any of them)
# ------------------------------------------------------------------------------------------------
class _Options:
    """set tsd options for the duration of a with-block (restored afterwards)"""

    def __init__(self, ctx, **kv):
        self.ctx, self.kv, self.old = ctx, kv, {}

    def __enter__(self):
        for k, v in self.kv.items():
            self.old[k] = self.ctx.get_option(k)
            self.ctx.set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            self.ctx.set_option(k, v)


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("bn,splits", [(0, 0), (64, 1), (160, 1), (80, 2), (160, 3)])
def test_conv2d_tile_variants(ctx, ops, cg, bn, splits):
    """single CTAs vs CTA pairs (tcgen05 cta_group::2), forced tile widths and split-K factors"""
    rng = np.random.default_rng(100 + bn + splits)
    x = rng.standard_normal((320, 32, 32), dtype=np.float32)
    wt = (rng.standard_normal((320, 320, 3, 3)) / np.sqrt(2880)).astype(np.float32)
    b = rng.standard_normal(320, dtype=np.float32)
    ref = ops.conv2d(x, wt, b, 1, 1)
    with _Options(ctx, autotune=0, gemm_cg=cg, force_bn=bn, force_splits=splits):
        y = ctx.conv2d(x, wt, b, pad=1)
    assert relerr(y, ref) < TOL_TF32


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("cg,bn,splits", [(1, 80, 2), (1, 160, 4), (2, 80, 3), (1, 64, 8), (2, 160, 4), (1, 256, 5), (2, 32, 4)])
def test_splitk_inside_a_cluster(ctx, ops, mode, cg, bn, splits):
    """split-K reduced inside the GEMM by the thread-block cluster of a tile's splits: partial tiles through L2
    (splitk_cluster = 1) or kept in shared memory and read across the cluster with ld.shared::cluster (= 2);
    a 3x3 convolution (ragged last chunk: 300 channels) and linears with ragged M / N"""
    rng = np.random.default_rng(300 + bn + splits + mode)
    x = rng.standard_normal((300, 32, 32), dtype=np.float32)
    wt = (rng.standard_normal((320, 300, 3, 3)) / np.sqrt(2700)).astype(np.float32)
    b = rng.standard_normal(320, dtype=np.float32)
    xs = rng.standard_normal((200, 1280), dtype=np.float32)
    ws = (rng.standard_normal((328, 1280)) / np.sqrt(1280)).astype(np.float32)
    bs = rng.standard_normal(328, dtype=np.float32)
    with _Options(ctx, autotune=0, gemm_cg=cg, force_bn=bn, force_splits=splits, splitk_cluster=mode, splitk_cluster_max=16):
        y = ctx.conv2d(x, wt, b, pad=1)
        ys = ctx.linear(xs, ws, bs)
        y2 = ctx.conv2d(x, wt, b, pad=1)
    assert relerr(y, ops.conv2d(x, wt, b, 1, 1)) < TOL_TF32
    assert relerr(ys, ops.linear(xs, ws, bs)) < TOL_TF32
    assert np.array_equal(y, y2)   # fixed summation order


@pytest.mark.parametrize("n,cin,h,w,cout,bn,splits,cg", [
    (1, 32, 32, 32, 32, 0, 0, 1), (1, 320, 32, 32, 320, 160, 1, 2), (1, 320, 32, 32, 320, 80, 2, 2),
    (2, 96, 40, 24, 48, 0, 0, 0),      # ragged edges, two images
    (1, 36, 20, 28, 80, 0, 0, 0),      # partial channel chunk
    (1, 64, 16, 16, 64, 0, 0, 1),      # TMA box taller than the image
    (2, 64, 8, 8, 32, 0, 0, 1),        # box larger than the image in both directions
])
def test_conv2d_halo_kernel(ctx, ops, n, cin, h, w, cout, bn, splits, cg):
    """3x3 convolution with the activation halo held in shared memory (nine taps = nine shifted
    UMMA descriptor views of one TMA tile): zero padding, ragged tiles, CTA pairs, split-K"""
    rng = np.random.default_rng(7 * cin + cout)
    x = rng.standard_normal((n, cin, h, w), dtype=np.float32)
    wt = (rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(cin * 9)).astype(np.float32)
    b = rng.standard_normal(cout, dtype=np.float32)
    ref = np.stack([ops.conv2d(x[i], wt, b, 1, 1) for i in range(n)])
    with _Options(ctx, autotune=0, conv_halo=2, halo_min_w=8, halo_min_h=8, gemm_cg=cg, force_bn=bn, force_splits=splits):
        y = ctx.conv2d(x, wt, b, pad=1)
    assert relerr(y, ref) < TOL_TF32


def test_linear_autotune_is_stable(ctx, ops):
    """the autotuner's choice is cached per context: repeated calls are bit-identical"""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1, 1024, 640), dtype=np.float32)
    w = (rng.standard_normal((640, 640)) / np.sqrt(640)).astype(np.float32)
    with _Options(ctx, autotune=1):
        y0 = ctx.linear(x, w, None)
        y1 = ctx.linear(x, w, None)
    assert np.array_equal(y0, y1)
    assert relerr(y0, ops.linear(x[0], w, None)[None]) < TOL_TF32


@pytest.mark.parametrize("path", ["fused", "v2", "cluster"])
@pytest.mark.parametrize("c,h,w,g,eps", [(320, 16, 16, 32, 1e-5), (1920, 16, 16, 32, 1e-5), (320, 8, 8, 320, 1e-5),
                                         (640, 32, 32, 1, 1e-5), (128, 64, 64, 32, 1e-6), (960, 64, 64, 32, 1e-5),
                                         (320, 64, 64, 32, 1e-5), (1280, 8, 8, 32, 1e-5), (96, 5, 7, 16, 1e-5)])
def test_groupnorm_fused_kernels(ctx, ops, path, c, h, w, g, eps):
    """the single-launch norm kernels: grid barrier v1, register-resident two-level barrier v2, and the
    cluster-per-(image, group) kernel (shapes it does not take - odd channels per group, one group - fall through)"""
    rng = np.random.default_rng(c + h)
    x = (rng.standard_normal((c, h, w)) * 2 + 0.5).astype(np.float32)
    with _Options(ctx, norm_v2=int(path == "v2"), norm_cluster=int(path == "cluster")):
        y = ctx.groupnorm(x, g, eps)
    assert relerr(y, ops.group_norm(x, g, eps)) < TOL_FP32
