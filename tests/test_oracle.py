"""CPU tests of the oracle itself: the two evaluators agree, the 'intended' switches match an
independent implementation (torch.nn.functional), and the committed goldens are reproduced.
PARITY UNPINNED by the reference (it has no tests): these are the pins this repo adds."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synth
import tsd_oracle as O
from conftest import GOLDEN, relerr


@pytest.fixture(scope="module")
def ops64():
    return O.Ops("np", np.float64)


@pytest.fixture(scope="module")
def opsc(oracle_lib):
    return O.Ops("c32")


def test_param_inventory_counts():
    # 299.74 M (UNet + time embed + final) and 49.47 M (decoder)
    assert synth.num_params(synth.diffusion_specs()) == 299_742_724
    assert synth.num_params(synth.decoder_specs()) == 49_467_159


def test_synth_generator_pinned(golden_small):
    got = synth.synth_tensor(1234, 3, 64, np.float32(0.5))
    assert np.array_equal(got, golden_small["synth_probe"])
    assert np.abs(got).max() < 0.5
    t = synth.synth_tensor(5, 0, 1 << 16, np.float32(1.0))
    assert abs(t.mean()) < 0.02 and abs(t.std() - 1 / np.sqrt(3)) < 0.01


def test_round_tf32():
    x = np.array([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -10, -1.0 - 2 ** -11, 3.14159265], np.float32)
    r = synth.round_tf32(x)
    assert r[0] == 1.0 and r[1] == np.float32(1.0 + 2 ** -10) and r[2] == np.float32(1.0 + 2 ** -10)
    assert r[3] == np.float32(-1.0 - 2 ** -10)
    assert abs(r[4] - x[4]) <= 2 ** -10


@pytest.mark.parametrize("stride,pad,k", [(1, 1, 3), (2, 1, 3), (1, 0, 1)])
def test_conv_vs_torch(ops64, opsc, stride, pad, k):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 9, 8))
    w = rng.standard_normal((7, 5, k, k))
    b = rng.standard_normal(7)
    ref = F.conv2d(torch.from_numpy(x)[None], torch.from_numpy(w), torch.from_numpy(b), stride=stride, padding=pad)[0].numpy()
    assert relerr(ops64.conv2d(x, w, b, pad, stride), ref) < 1e-12
    assert relerr(opsc.conv2d(x, w, b, pad, stride), ref) < 1e-5


def test_conv_reads_only_in_channels(ops64):
    # Q9: a conv built for 4 input channels applied to a 6-channel tensor reads the first 4
    rng = np.random.default_rng(1)
    x = rng.standard_normal((6, 5, 5))
    w = rng.standard_normal((3, 4, 3, 3))
    assert np.array_equal(ops64.conv2d(x, w, None, 1), ops64.conv2d(x[:4], w, None, 1))


def test_groupnorm_formula(ops64, opsc):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((12, 4, 5)) * 3 + 1
    # independent: torch group_norm divides by sqrt(var+eps); with eps=0 both agree exactly
    ref0 = F.group_norm(torch.from_numpy(x)[None], 3, eps=0.0)[0].numpy()
    assert relerr(ops64.group_norm(x, 3, 0.0), ref0) < 1e-12
    # reference formula: eps is added to the std (utils.mojo:1868-1870)
    g = x.reshape(3, -1)
    want = ((g - g.mean(1, keepdims=True)) / (g.std(1, keepdims=True) + 1e-5)).reshape(x.shape)
    assert relerr(ops64.group_norm(x, 3, 1e-5), want) < 1e-13
    assert relerr(opsc.group_norm(x, 3, 1e-5), want) < 1e-5
    with pytest.raises(ValueError):
        ops64.group_norm(x, 5)


def test_layernorm_modes(ops64):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((10, 6)) * 2 + 0.5   # (T,C)
    glob = ops64.layer_norm(x)
    assert relerr(glob, (x - x.mean()) / (x.std() + 1e-5)) < 1e-13          # Q5: one mean/std for the tensor
    tok = O.Ops("np", np.float64, O.Switches(layernorm="token")).layer_norm(x)
    ref = F.layer_norm(torch.from_numpy(x), (6,), eps=0.0).numpy()
    assert relerr(tok, ref) < 1e-4


def test_activations_vs_torch(ops64, opsc):
    x = np.linspace(-6, 6, 1001)
    assert relerr(ops64.silu(x), F.silu(torch.from_numpy(x)).numpy()) < 1e-13
    assert relerr(ops64.gelu(x), F.gelu(torch.from_numpy(x), approximate="tanh").numpy()) < 1e-13
    assert relerr(opsc.silu(x), ops64.silu(x)) < 1e-6
    assert relerr(opsc.gelu(x), ops64.gelu(x)) < 1e-6
    u = ops64.upsample2x(np.arange(12.0).reshape(1, 3, 4))
    assert np.array_equal(u, F.interpolate(torch.arange(12.0).reshape(1, 1, 3, 4), scale_factor=2, mode="nearest")[0].numpy())


def test_softmax_axes(ops64, opsc):
    rng = np.random.default_rng(4)
    s = rng.standard_normal((3, 6, 5))
    d2 = ops64.softmax(s, 2)
    assert np.allclose(d2.sum(axis=1), 1.0)       # Q3: dim=2 normalises every column over the rows
    assert np.allclose(ops64.softmax(s, 1).sum(axis=2), 1.0)
    assert relerr(ops64.softmax(s, 1), torch.softmax(torch.from_numpy(s), dim=2).numpy()) < 1e-13
    assert relerr(opsc.softmax(s, 2), d2) < 1e-6   # no max-subtraction in the C loops (utils.mojo:413)
    with pytest.raises(ValueError):
        ops64.softmax(s, 3)


def test_attention_key_axis_vs_sdpa():
    rng = np.random.default_rng(5)
    h, t, d, c = 4, 9, 8, 32
    x = rng.standard_normal((t, c))
    w_in = rng.standard_normal((3 * c, c)) / np.sqrt(c)
    w_out = rng.standard_normal((c, c)) / np.sqrt(c)
    b_out = rng.standard_normal(c)
    ops = O.Ops("np", np.float64, O.Switches(softmax_axis="key"))
    got = ops.self_attention(x, h, w_in, None, w_out, b_out)
    # independent restatement with the raw-reshape head split (Q4) done by torch views
    qkv = torch.from_numpy(x) @ torch.from_numpy(w_in).T
    q, k, v = (qkv[:, i * c:(i + 1) * c].contiguous().view(h, t, d) for i in range(3))
    o = F.scaled_dot_product_attention(q[None], k[None], v[None])[0]          # (h,t,d)
    o = o.transpose(0, 1).reshape(t, c)
    ref = (o @ torch.from_numpy(w_out).T + torch.from_numpy(b_out)).numpy()
    assert relerr(got, ref) < 1e-12


def test_attention_query_axis_definition(ops64, opsc):
    rng = np.random.default_rng(6)
    q = rng.standard_normal((2, 7, 4))
    k = rng.standard_normal((2, 5, 4))
    v = rng.standard_normal((2, 5, 4))
    s = np.einsum("hid,hjd->hij", q, k) / 2.0
    p = np.exp(s) / np.exp(s).sum(axis=1, keepdims=True)      # normalised over the query index
    want = np.einsum("hij,hjd->ihd", p, v).reshape(7, 8)
    assert relerr(ops64.attention_core(q, k, v), want) < 1e-13
    assert relerr(opsc.attention_core(q, k, v), want) < 1e-5


def test_sampler_against_closed_form():
    sm = O.DDPMSampler()
    sm.set_inference_timesteps(20)
    assert list(sm.timesteps[:3]) == [950, 900, 850] and sm.timesteps[-1] == 0
    ab = sm.alphas_cumprod
    t, p = 500, 450
    s_ab, s_1mab, c0, c1, sigma = sm.coefficients(t)
    # posterior mean/variance of q(x_{p} | x_t, x_0) (Ho et al. eq. 7) with alpha_t = ab_t/ab_p
    a_t = ab[t] / ab[p]
    assert np.isclose(c0, np.sqrt(ab[p]) * (1 - a_t) / (1 - ab[t]))
    assert np.isclose(c1, np.sqrt(a_t) * (1 - ab[p]) / (1 - ab[t]))
    assert np.isclose(sigma ** 2, (1 - ab[p]) / (1 - ab[t]) * (1 - a_t))
    assert sm.coefficients(0)[4] == 0.0 and np.isclose(sm.coefficients(0)[2], 1.0)   # t=0: x_prev = x0_hat
    x = np.ones((4, 2, 2))
    assert np.allclose(sm.step(0, x, np.zeros_like(x)), x / sm.coefficients(0)[0])


def test_time_embedding_variants():
    e = O.get_time_embedding(999)
    assert e.shape == (320,) and np.isclose(e[0], np.cos(999.0), atol=1e-6) and np.isclose(e[160], np.sin(999.0), atol=1e-6)
    w = O.get_time_embedding(999, as_written=True)   # Q12: freqs ~ 0 except index 0 -> [1]*160 + [0]*160
    assert np.allclose(w[1:159], 1.0) and np.allclose(w[161:319], 0.0)


def test_goldens_small_reproduced(golden_small, ops64, opsc):
    g = golden_small
    assert relerr(ops64.conv2d(g["conv_x"], g["conv_w"], g["conv_b"], 1), g["conv_y"]) < 1e-13
    assert relerr(ops64.conv2d(g["conv_x"], g["conv_w"], g["conv_b"], 1, 2), g["conv_y_s2"]) < 1e-13
    assert relerr(opsc.conv2d(g["conv_x"], g["conv_w"], g["conv_b"], 1), g["conv_y"]) < 1e-5
    assert relerr(ops64.group_norm(g["conv_x"], 3), g["gn_y"]) < 1e-13
    assert relerr(ops64.softmax(g["sm_x"], 2), g["sm_dim2"]) < 1e-13
    assert relerr(ops64.attention_core(g["at_q"], g["at_k"], g["at_v"]), g["at_query"]) < 1e-13
    assert relerr(opsc.attention_core(g["at_q"], g["at_k"], g["at_v"]), g["at_query"]) < 1e-5
    sm = O.DDPMSampler()
    sm.set_inference_timesteps(20)
    assert np.array_equal(sm.timesteps, g["sched_t"])
    assert np.allclose(np.stack([sm.coefficients(int(t)) for t in sm.timesteps]), g["sched_coef"], rtol=1e-12)


def test_unet8_c_loops_match_fp64_golden(golden_small, opsc, unet_weights):
    """The reference's fp32 scalar loop order (C) against the fp64 golden: plumbing config 1
    (one DDPM step forward, random weights) at an 8x8 latent."""
    g = golden_small
    y = O.diffusion_forward(opsc, unet_weights, g["unet8_x"], g["unet8_ctx"], g["unet8_t"])
    assert y.shape == (4, 8, 8)
    assert relerr(y, g["unet8_y"]) < 1e-4
    # the semantic switches change the answer (they are not no-ops)
    assert relerr(g["unet8_y_intended"], g["unet8_y"]) > 1e-3
    assert relerr(g["unet8_y_alias"], g["unet8_y"]) > 1e-4


def test_decoder8_c_loops_match_fp64_golden(golden_small, opsc, decoder_weights):
    g = golden_small
    y = O.decoder_forward(opsc, decoder_weights, g["dec8_z"])
    assert y.shape == (3, 64, 64)
    assert relerr(y, g["dec8_y"]) < 1e-4
    img = O.rescale_image(y)
    assert img.min() >= 0 and img.max() <= 255


def test_full_size_goldens_present():
    for name in ("unet64.npz", "decoder64.npz"):
        assert os.path.exists(os.path.join(GOLDEN, name)), f"run tools/make_golden.py {name[:-4]}"
    u = np.load(os.path.join(GOLDEN, "unet64.npz"))
    assert u["y"].shape == (4, 64, 64) and np.isfinite(u["y"]).all()


# ---- CLIP text encoder ------------------------------------------------
def test_clip_param_inventory():
    # clip.mojo:71-83: 49408 x 768 token table, 77 x 768 positions, 12 layers of 7 084 800 parameters
    assert synth.num_params(synth.clip_specs()) == 49408 * 768 + 77 * 768 + 12 * 7_084_800 == 123_022_080


def test_clip_intended_switches_vs_torch():
    """With the 'intended' switches (key-axis softmax, per-token LayerNorm) the oracle's CLIP layer is the
    standard pre-LN transformer layer with a causal mask and quick-GELU: cross-checked against an
    independent torch restatement (raw-reshape head split kept, Q4)."""
    n_vocab, n_layers, c, t, h = 50, 2, 768, 77, 12
    specs = synth.clip_specs(n_vocab, n_layers)
    blob = synth.random_blob(specs, 3)
    W = synth.BlobWeights(specs, blob)
    tokens = np.random.default_rng(1).integers(0, n_vocab, 11)
    ops = O.Ops("np", np.float64, O.Switches(softmax_axis="key", layernorm="token"))
    got = O.clip_forward(ops, W, tokens, n_layers=n_layers)

    tw = lambda name: torch.from_numpy(np.asarray(W[name], np.float64))  # noqa: E731
    tok = np.zeros(t, np.int64)
    tok[:tokens.size] = tokens
    x = tw("embedding.token_embedding.weight").view(-1, c)[torch.from_numpy(tok)] + tw("embedding.position_embedding").view(t, c)

    def ln(v):  # (x - mean) / (std + eps), biased std: helpers/utils.mojo:1868-1870 (not sqrt(var + eps))
        m = v.mean(-1, keepdim=True)
        s = ((v - m) ** 2).mean(-1, keepdim=True).sqrt()
        return (v - m) / (s + 1e-5)

    for l in range(1, n_layers + 1):
        b = f"player{l}"
        r = x
        qkv = ln(x) @ tw(b + ".layer2.in_proj.weight").T + tw(b + ".layer2.in_proj.bias")
        q, k, v = (qkv[:, i * c:(i + 1) * c].contiguous().view(h, t, c // h) for i in range(3))
        o = F.scaled_dot_product_attention(q[None], k[None], v[None], is_causal=True)[0]
        o = o.transpose(0, 1).reshape(t, c)
        x = o @ tw(b + ".layer2.out_proj.weight").T + tw(b + ".layer2.out_proj.bias") + r
        r = x
        y = ln(x) @ tw(b + ".layer4.weight").T + tw(b + ".layer4.bias")
        y = y * torch.sigmoid(1.702 * y)
        x = y @ tw(b + ".layer5.weight").T + tw(b + ".layer5.bias") + r
    ref = ln(x).numpy()
    assert got.shape == (77, 768)
    assert relerr(got, ref) < 1e-10


def test_clip_default_switches_are_the_reference_deviations():
    """Default switches keep the reference's deterministic deviations: column softmax over the (unmasked)
    queries (Q3) and one mean/std over the whole (C, T) tensor (Q5)."""
    n_vocab, n_layers = 40, 1
    specs = synth.clip_specs(n_vocab, n_layers)
    W = synth.BlobWeights(specs, synth.random_blob(specs, 5))
    tokens = np.arange(20) % n_vocab
    a = O.clip_forward(O.Ops("np", np.float64), W, tokens, n_layers=n_layers)
    b = O.clip_forward(O.Ops("np", np.float64, O.Switches(softmax_axis="key", layernorm="token")), W, tokens, n_layers=n_layers)
    assert np.isfinite(a).all() and relerr(a, b) > 1e-3   # genuinely different semantics
    # global LayerNorm: the output as a whole has zero mean and unit (biased) std
    assert abs(a.mean()) < 1e-9 and abs(a.std() - 1.0) < 1e-4


def test_clip_golden_reproduced():
    """tests/golden/clip_small.npz (tools/make_golden.py clip) pins the CLIP restatement."""
    g = np.load(os.path.join(GOLDEN, "clip_small.npz"))
    W = synth.SynthWeights(synth.clip_specs(1000, 3), 77)
    y = O.clip_forward(O.Ops("np", np.float64), W, g["tokens"], n_layers=3)
    assert relerr(y, g["y_reference_switches"]) < 1e-12


# ---- VAE Encoder / img2img ---------------------------------------------
def test_encoder_param_inventory():
    # vae.mojo:94-112, GroupNorm owns no tensor (utils.mojo:1825-1872)
    def conv(ci, co, k):
        return co * ci * k * k + co

    def res(ci, co):
        return conv(ci, co, 3) + conv(co, co, 3) + (conv(ci, co, 1) if ci != co else 0)

    want = (conv(3, 128, 3) + 2 * res(128, 128) + conv(128, 128, 3) + res(128, 256) + res(256, 256) + conv(256, 256, 3)
            + res(256, 512) + res(512, 512) + conv(512, 512, 3) + 4 * res(512, 512) + (512 * 1536 + 1536)
            + (512 * 512 + 512) + conv(512, 8, 3) + conv(8, 8, 1))
    # = the published SD-v1 VAE encoder (34 163 592) - its GroupNorm affine tensors (16 640) + quant_conv 8->8 (72)
    assert synth.num_params(synth.encoder_specs()) == want == 34_163_592 - 16_640 + 72


def test_asymmetric_pad_conv_vs_torch(ops64, opsc):
    """two_stride_pad + stride-2 conv (vae.mojo:97,115-116) = F.pad(x, (0,1,0,1)) + conv2d(stride 2)."""
    rng = np.random.default_rng(3)
    for (h, w) in ((8, 8), (7, 10), (2, 2)):
        x = rng.standard_normal((5, h, w))
        wt = rng.standard_normal((6, 5, 3, 3)) * 0.2
        b = rng.standard_normal(6)
        ref = F.conv2d(F.pad(torch.from_numpy(x)[None], (0, 1, 0, 1)), torch.from_numpy(wt), torch.from_numpy(b),
                       stride=2)[0].numpy()
        got = ops64.conv2d(x, wt, b, pad=0, stride=2, pad_hi=1)
        assert got.shape == ref.shape == (6, (h + 1 - 3) // 2 + 1, (w + 1 - 3) // 2 + 1)
        assert relerr(got, ref) < 1e-12
        assert relerr(opsc.conv2d(x, wt, b, pad=0, stride=2, pad_hi=1), ref) < 1e-5


def test_latent_from_moments_definition(ops64):
    """metrics_evals, vae.mojo:118-129, including the (-30, 20) clamp of the log-variance."""
    m = np.zeros((8, 1, 3))
    m[:4, 0] = [[1.0, -2.0, 0.5]] * 4
    m[4:, 0] = [[0.0, 100.0, -100.0]] * 4          # -> std 1, e^10, e^-15
    noise = np.full((4, 1, 3), 2.0)
    got = O.latent_from_moments(ops64, m, noise)
    want = (np.array([1.0, -2.0, 0.5]) + 2.0 * np.array([1.0, np.exp(10.0), np.exp(-15.0)])) * 0.18215
    assert np.allclose(got[2, 0], want, rtol=1e-14)


def test_resize_and_rescale_input():
    img = np.arange(2 * 4 * 6, dtype=np.float64).reshape(2, 4, 6)
    assert O.resize_image(img, 4, 6) is img
    up = O.resize_image(img, 8, 12)            # nearest neighbour: every source pixel doubled
    assert np.array_equal(up, img.repeat(2, 1).repeat(2, 2))
    down = O.resize_image(img, 2, 3)
    assert np.array_equal(down, img[:, ::2, ::2])
    assert np.allclose(O.rescale_input(np.array([0.0, 127.5, 255.0])), [-1.0, 0.0, 1.0])


def test_set_strength_slices_the_schedule():
    sm = O.DDPMSampler()
    sm.set_inference_timesteps(20)
    full = sm.timesteps.copy()
    sm.set_strength(0.8)                       # start_step = 20 - int(16.0) = 4 (sampler.mojo:67-73)
    assert sm.start_step == 4 and np.array_equal(sm.timesteps, full[4:])
    x, nz = np.ones(5), np.full(5, 2.0)
    ab = sm.alphas_cumprod[int(sm.timesteps[0])]
    assert np.allclose(sm.add_noise(x, sm.timesteps[0], nz), np.sqrt(ab) + 2.0 * np.sqrt(1 - ab))


def test_encoder_intended_switches_vs_torch():
    """Independent torch restatement of Encoder.forward (key-axis softmax; GroupNorm keeps the reference's
    (x - mean) / (std + eps) formula) at a 32x32 image."""
    specs = synth.encoder_specs()
    W = synth.BlobWeights(specs, synth.random_blob(specs, 9))
    rng = np.random.default_rng(2)
    x = rng.uniform(-1, 1, (3, 32, 32))
    noise = rng.standard_normal((4, 4, 4))
    got = O.encoder_forward(O.Ops("np", np.float64, O.Switches(softmax_axis="key", layernorm="token")), W, x, noise)

    tw = lambda name: torch.from_numpy(np.asarray(W[name], np.float64))  # noqa: E731

    def gn(v, g):
        c, h, w = v.shape
        r = v.reshape(g, -1)
        m = r.mean(1, keepdim=True)
        s = ((r - m) ** 2).mean(1, keepdim=True).sqrt()
        return ((r - m) / (s + 1e-5)).reshape(c, h, w)

    def conv(v, name, pad=0, stride=1):
        return F.conv2d(v[None], tw(name + ".weight"), tw(name + ".bias"), padding=pad, stride=stride)[0]

    def res(v, name, ci, co):
        o = conv(F.silu(gn(v, 16)), name + ".conv1", 1)
        o = conv(F.silu(gn(o, 16)), name + ".conv2", 1)
        return o + (conv(v, name + ".res_conv_layer") if ci != co else v)

    def down(v, name):
        return conv(F.pad(v, (0, 1, 0, 1)), name, 0, 2)

    v = conv(torch.from_numpy(x), "l1", 1)
    v = res(res(v, "l2", 128, 128), "l3", 128, 128)
    v = down(v, "l4")
    v = res(res(v, "l5", 128, 256), "l6", 256, 256)
    v = down(v, "l7")
    v = res(res(v, "l8", 256, 512), "l9", 512, 512)
    v = down(v, "l10")
    for n in ("l11", "l12", "l13"):
        v = res(v, n, 512, 512)
    c, h, w = v.shape
    seq = gn(v, 32).reshape(c, h * w).T
    qkv = seq @ tw("l14.attention.in_proj.weight").T + tw("l14.attention.in_proj.bias")
    q, k, vv = (qkv[:, i * c:(i + 1) * c] for i in range(3))
    o = F.scaled_dot_product_attention(q[None, None], k[None, None], vv[None, None])[0, 0]
    o = o @ tw("l14.attention.out_proj.weight").T + tw("l14.attention.out_proj.bias")
    v = o.T.reshape(c, h, w) + v
    v = res(v, "l15", 512, 512)
    v = conv(F.silu(gn(v, 32)), "l18", 1)
    v = conv(v, "l19")
    mean, logvar = v[:4], v[4:].clamp(-30, 20)
    ref = ((mean + torch.from_numpy(noise) * (0.5 * logvar).exp()) * 0.18215).numpy()
    assert got.shape == (4, 4, 4)
    assert relerr(got, ref) < 1e-10


def test_encoder_golden_reproduced(opsc):
    """tests/golden/encoder_small.npz (tools/make_golden.py encoder) pins the Encoder restatement; the C
    loops (fp32) agree with the fp64 evaluation."""
    g = np.load(os.path.join(GOLDEN, "encoder_small.npz"))
    W = synth.SynthWeights(synth.encoder_specs(), 1236)
    x = O.rescale_input(g["img4"].astype(np.float64))
    z = O.encoder_forward(O.Ops("np", np.float64), W, x, g["noise4"])
    assert relerr(z, g["z4"]) < 1e-12
    assert relerr(O.encoder_forward(opsc, W, x.astype(np.float32), g["noise4"]), g["z4"]) < 2e-4
    assert os.path.exists(os.path.join(GOLDEN, "encoder64.npz")), "run tools/make_golden.py encoder64"
    sm = O.DDPMSampler()
    sm.set_inference_timesteps(5)
    sm.set_strength(0.6)
    assert np.array_equal(sm.timesteps, g["i2i_timesteps"])
    assert relerr(sm.add_noise(g["z16"], sm.timesteps[0], g["i2i_start_noise"].astype(np.float64)), g["i2i_start"]) < 1e-14


def test_affine_norms_and_eps_mode_match_torch():
    """Row f2: per-channel norm weights and the standard eps placement (norm_eps_inside) of the oracle against
    torch.nn.functional - what a real checkpoint needs on top of the reference's scalar-gamma (x - mean) / (std + eps)."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(3)
    ops = O.Ops("np", np.float64, O.Switches(layernorm="token", norm_eps_inside=True))
    x = rng.standard_normal((32, 6, 5))
    g, b = rng.standard_normal(32), rng.standard_normal(32)
    want = F.group_norm(torch.from_numpy(x)[None], 8, torch.from_numpy(g), torch.from_numpy(b), 1e-5)[0].numpy()
    assert np.abs(ops.group_norm(x, 8, 1e-5, g, b) - want).max() < 1e-12
    seq = rng.standard_normal((11, 32))
    want = F.layer_norm(torch.from_numpy(seq), (32,), torch.from_numpy(g), torch.from_numpy(b), 1e-5).numpy()
    assert np.abs(ops.layer_norm(seq, g, b) - want).max() < 1e-12
    ref = O.Ops("np", np.float64)                       # reference placement differs, by construction
    assert np.abs(ref.group_norm(x, 8, 1e-1) - ops.group_norm(x, 8, 1e-1)).max() > 1e-3
