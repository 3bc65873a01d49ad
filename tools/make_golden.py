"""Generates tests/golden/*.npz with the fp64 oracle (oracle/tsd_oracle.py, backend "np").

PARITY UNPINNED: the reference has no golden vectors and cannot run here; these files pin the
repo's own restatement so that (a) the oracle cannot drift silently and (b) the GPU path is
checked at the full BASELINE sizes, where running the oracle inside a test would take minutes.

  python tools/make_golden.py small      # seconds-to-a-minute cases (ops, 8x8 UNet, 8x8 decoder, loop)
  python tools/make_golden.py unet64     # one UNet step at the 64x64 latent of BASELINE config 2
  python tools/make_golden.py decoder64  # one VAE decode 64x64x4 -> 512x512x3 (stored subsampled)
  python tools/make_golden.py clip       # CLIP text encoder, 1000-token vocabulary, 3 layers, both switch sets
  python tools/make_golden.py encoder    # VAE Encoder at 32x32 / 128x128 images + the img2img start latents
  python tools/make_golden.py encoder64  # VAE Encoder 512x512x3 -> 64x64x4
  python tools/make_golden.py tokenizer  # synthetic tokenizer .bin + known-answer token ids
  python tools/make_golden.py loop64     # 20-step denoising loop at 64x64 (BASELINE config 2) with the latents after
                                         # steps 1, 2, 5, 10, 20, and a 4-step CFG 7.5 loop (config 3 per-GPU shape)
  python tools/make_golden.py decoder64_full  # two distinct 512x512 decodes kept whole (fp16 x scale): batch-16 test
  python tools/make_golden.py clip12     # the 12-layer, 49408-token CLIP of clip.mojo:71-83
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth  # noqa: E402
import tsd_oracle as O  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
UNET_SEED, DEC_SEED = 1234, 1235


def inputs(seed, side, n_ctx=1):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((4, side, side), dtype=np.float32)
    ctx = rng.standard_normal((n_ctx, 77, 768), dtype=np.float32)
    return x, ctx


def small():
    ops = O.Ops("np", np.float64)
    rng = np.random.default_rng(7)
    out = {}
    # op level known answers
    x = rng.standard_normal((6, 5, 7), dtype=np.float32)
    w = rng.standard_normal((4, 6, 3, 3), dtype=np.float32) * 0.2
    b = rng.standard_normal(4, dtype=np.float32)
    out.update(conv_x=x, conv_w=w, conv_b=b, conv_y=ops.conv2d(x, w, b, pad=1), conv_y_s2=ops.conv2d(x, w, b, pad=1, stride=2))
    out.update(gn_y=ops.group_norm(x, 3, 1e-5), silu_y=ops.silu(x), gelu_y=ops.gelu(x), up_y=ops.upsample2x(x))
    s = rng.standard_normal((2, 5, 7), dtype=np.float32)
    out.update(sm_x=s, sm_dim2=ops.softmax(s, 2), sm_dim1=ops.softmax(s, 1))
    q = rng.standard_normal((2, 6, 4), dtype=np.float32)
    k = rng.standard_normal((2, 5, 4), dtype=np.float32)
    v = rng.standard_normal((2, 5, 4), dtype=np.float32)
    out.update(at_q=q, at_k=k, at_v=v, at_query=ops.attention_core(q, k, v),
               at_key=O.Ops("np", np.float64, O.Switches(softmax_axis="key")).attention_core(q, k, v))
    sm = O.DDPMSampler()
    sm.set_inference_timesteps(20)
    out.update(sched_t=sm.timesteps, sched_coef=np.stack([sm.coefficients(int(t)) for t in sm.timesteps]),
               temb_999=O.get_time_embedding(999), temb_999_as_written=O.get_time_embedding(999, True))
    out["synth_probe"] = synth.synth_tensor(UNET_SEED, 3, 64, np.float32(0.5))
    # UNet step, 8x8 latent, default (reference-faithful) switches and the intended ones
    W = synth.SynthWeights(synth.diffusion_specs(), UNET_SEED)
    x8, ctx = inputs(11, 8)
    t = O.get_time_embedding(999)
    out.update(unet8_x=x8, unet8_ctx=ctx[0], unet8_t=t)
    out["unet8_y"] = O.diffusion_forward(ops, W, x8, ctx[0], t)
    out["unet8_y_intended"] = O.diffusion_forward(
        O.Ops("np", np.float64, O.Switches(softmax_axis="key", layernorm="token")), W, x8, ctx[0], t)
    out["unet8_y_alias"] = O.diffusion_forward(O.Ops("np", np.float64, O.Switches(mojo_alias_time=True)), W, x8, ctx[0], t)
    # 3-step loop with CFG at 8x8
    rng2 = np.random.default_rng(12)
    noise = rng2.standard_normal((3, 4, 8, 8), dtype=np.float32)
    uctx = rng2.standard_normal((77, 768), dtype=np.float32)
    out.update(loop8_noise=noise, loop8_uctx=uctx)
    out["loop8_lat"] = O.generate_latents(ops, W, x8, ctx[0], 3, noise, cfg_context=uctx, cfg_scale=7.5)
    # decoder at 8x8 latent
    Wd = synth.SynthWeights(synth.decoder_specs(), DEC_SEED)
    z = (np.random.default_rng(13).standard_normal((4, 8, 8)) * 0.18215).astype(np.float32)
    out.update(dec8_z=z, dec8_y=O.decoder_forward(ops, Wd, z))
    np.savez_compressed(os.path.join(G, "small.npz"), **{k: np.asarray(v) for k, v in out.items()})


def unet64():
    ops = O.Ops("np", np.float64)
    W = synth.SynthWeights(synth.diffusion_specs(), UNET_SEED)
    x, ctx = inputs(21, 64)
    t = O.get_time_embedding(999)
    t0 = time.time()
    y = O.diffusion_forward(ops, W, x, ctx[0], t)
    print("unet64 oracle seconds", time.time() - t0, flush=True)
    np.savez_compressed(os.path.join(G, "unet64.npz"), x=x, ctx=ctx[0], t=t, y=y.astype(np.float64))


def decoder64():
    ops = O.Ops("np", np.float64)
    Wd = synth.SynthWeights(synth.decoder_specs(), DEC_SEED)
    z = (np.random.default_rng(31).standard_normal((4, 64, 64)) * 0.18215).astype(np.float32)
    t0 = time.time()
    y = O.decoder_forward(ops, Wd, z)
    print("decoder64 oracle seconds", time.time() - t0, flush=True)
    # full image is 3 MB: keep a strided subsample plus global moments
    np.savez_compressed(os.path.join(G, "decoder64.npz"), z=z, y_sub=y[:, 3::8, 5::8].astype(np.float64),
                        y_mean=y.mean(axis=(1, 2)), y_std=y.std(axis=(1, 2)), y_absmax=np.abs(y).max())


CLIP_SEED, CLIP_VOCAB, CLIP_LAYERS = 77, 1000, 3


def clip():
    W = synth.SynthWeights(synth.clip_specs(CLIP_VOCAB, CLIP_LAYERS), CLIP_SEED)
    tokens = np.random.default_rng(4).integers(0, CLIP_VOCAB, 23)
    t0 = time.time()
    y_ref = O.clip_forward(O.Ops("np", np.float64), W, tokens, n_layers=CLIP_LAYERS)
    y_int = O.clip_forward(O.Ops("np", np.float64, O.Switches(softmax_axis="key", layernorm="token")), W, tokens,
                           n_layers=CLIP_LAYERS)
    print(f"clip oracle x2: {time.time() - t0:.1f} s")
    np.savez_compressed(os.path.join(G, "clip_small.npz"), tokens=tokens, y_reference_switches=y_ref,
                        y_intended_switches=y_int)


ENC_SEED = 1236


def encoder_inputs(seed, side):
    """image in (0,255) of (3, 8*side, 8*side) and the reparameterisation noise (4, side, side)"""
    rng = np.random.default_rng(seed)
    img = rng.uniform(0.0, 255.0, (3, 8 * side, 8 * side)).astype(np.float32)
    noise = rng.standard_normal((4, side, side), dtype=np.float32)
    return img, noise


def encoder():
    """VAE Encoder (row f3) at 32x32 and 128x128 images (latents 4x4 and 16x16), both switch sets."""
    We = synth.SynthWeights(synth.encoder_specs(), ENC_SEED)
    out = {}
    t0 = time.time()
    for side in (4, 16):
        img, noise = encoder_inputs(40 + side, side)
        out[f"img{side}"], out[f"noise{side}"] = img, noise
        out[f"z{side}"] = O.encoder_forward(O.Ops("np", np.float64), We, O.rescale_input(img.astype(np.float64)), noise)
        out[f"z{side}_intended"] = O.encoder_forward(
            O.Ops("np", np.float64, O.Switches(softmax_axis="key", layernorm="token")), We,
            O.rescale_input(img.astype(np.float64)), noise)
    # img2img start (pipeline.mojo:66-79) at the 16x16 latent: strength 0.6 of 5 steps
    sm = O.DDPMSampler()
    sm.set_inference_timesteps(5)
    sm.set_strength(0.6)
    start_noise = np.random.default_rng(77).standard_normal((4, 16, 16), dtype=np.float32)
    out.update(i2i_timesteps=sm.timesteps, i2i_start_noise=start_noise,
               i2i_start=sm.add_noise(out["z16"], sm.timesteps[0], start_noise.astype(np.float64)))
    print(f"encoder oracle: {time.time() - t0:.1f} s")
    np.savez_compressed(os.path.join(G, "encoder_small.npz"), **out)


def encoder64():
    We = synth.SynthWeights(synth.encoder_specs(), ENC_SEED)
    img, noise = encoder_inputs(51, 64)
    t0 = time.time()
    z = O.encoder_forward(O.Ops("np", np.float64), We, O.rescale_input(img.astype(np.float64)), noise)
    print("encoder64 oracle seconds", time.time() - t0, flush=True)
    # the 512x512 image is regenerated from its seed in the test; only the latent is stored
    np.savez_compressed(os.path.join(G, "encoder64.npz"), noise=noise, z=z.astype(np.float64))


TOK_PROMPTS = ["a cat flying a spaceship", "photo of an astronaut riding a horse on mars", "", "x",
               "it's a \"quoted\" prompt", "tab\there", "caf\u00e9 at night", "the  double  space",
               "highly detailed oil painting of an old castle at sunset"]


def tokenizer():
    """Row f4: a synthetic vocabulary in tokenizer_clip.bin layout (written by the restatement of
    tokenizer_creation.py) and the ids the restated bpe_encode gives, for both str_concat readings."""
    import json
    import tokenizer_oracle as T
    keys, merges = T.synthetic_vocab()
    blob = T.tokenizer_bin(keys, merges)
    with open(os.path.join(G, "tokenizer_small.bin"), "wb") as f:
        f.write(blob)
    tok = T.Tokenizer(len(keys), blob)
    cases = []
    for p in TOK_PROMPTS:
        text = T.preprocess_prompt(p)
        a, ok_a = T.bpe_encode(text, tok, True)
        b, ok_b = T.bpe_encode(text, tok, False)
        cases.append(dict(prompt=p, as_written=a, intended=b, complete=ok_a and ok_b))
    with open(os.path.join(G, "tokenizer_small.json"), "w") as f:
        json.dump(dict(vocab_size=len(keys), n_merges=len(merges), cases=cases), f)
    print("tokenizer golden:", len(keys), "tokens,", len(blob), "bytes")


def _loop_trace(ops, W, lat0, context, steps, noise, keep, cfg_context=None, cfg_scale=7.5):
    """O.generate_latents (pipeline.mojo:86-122) with the latents after the steps listed in `keep` recorded."""
    sm = O.DDPMSampler()
    sm.set_inference_timesteps(steps)
    lat = ops.arr(lat0)
    out = {}
    for i, t in enumerate(sm.timesteps):
        t0 = time.time()
        temb = O.get_time_embedding(float(t))
        eps = O.diffusion_forward(ops, W, lat, context, temb)
        if cfg_context is not None:
            eps = O.cfg_combine(eps, O.diffusion_forward(ops, W, lat, cfg_context, temb), cfg_scale)
        lat = sm.step(int(t), lat, eps, ops.arr(noise[i]))
        if i + 1 in keep:
            out[i + 1] = np.asarray(lat, np.float64).copy()
        print(f"  step {i + 1}/{steps}: {time.time() - t0:.1f} s", flush=True)
    return out


def loop64():
    """BASELINE config 2 (20 steps, no CFG) and config 3's per-GPU shape (CFG 7.5; 4 steps of its 50) at the 64x64
    latent, fp64.  Inputs are regenerated from their seeds in the test; only the latents are stored."""
    ops = O.Ops("np", np.float64)
    W = synth.SynthWeights(synth.diffusion_specs(), UNET_SEED)
    x, ctx = inputs(61, 64, n_ctx=2)
    rng = np.random.default_rng(62)
    noise = rng.standard_normal((20, 4, 64, 64), dtype=np.float32)
    keep = (1, 2, 5, 10, 20)
    tr = _loop_trace(ops, W, x, ctx[0], 20, noise, keep)
    out = {f"lat_step{k}": v for k, v in tr.items()}
    trc = _loop_trace(ops, W, x, ctx[0], 4, noise[:4], (1, 2, 4), cfg_context=ctx[1], cfg_scale=7.5)
    out.update({f"cfg_lat_step{k}": v for k, v in trc.items()})
    np.savez_compressed(os.path.join(G, "loop64.npz"), **out)


def decoder64_full():
    ops = O.Ops("np", np.float64)
    Wd = synth.SynthWeights(synth.decoder_specs(), DEC_SEED)
    out = {}
    for i, seed in enumerate((31, 32)):
        z = (np.random.default_rng(seed).standard_normal((4, 64, 64)) * 0.18215).astype(np.float32)
        t0 = time.time()
        y = O.decoder_forward(ops, Wd, z)
        print("decoder64 oracle seconds", time.time() - t0, flush=True)
        scale = float(np.abs(y).max())
        out[f"y{i}_f16"] = (y / scale).astype(np.float16)   # 2^-11 relative to the image maximum: 40x below the tolerance
        out[f"y{i}_scale"] = scale
    np.savez_compressed(os.path.join(G, "decoder64_full.npz"), **out)


CLIP12_SEED = 78


def clip12():
    W = synth.SynthWeights(synth.clip_specs(49408, 12), CLIP12_SEED)
    tokens = np.random.default_rng(5).integers(0, 49408, 77)
    t0 = time.time()
    y_ref = O.clip_forward(O.Ops("np", np.float64), W, tokens, n_layers=12)
    y_int = O.clip_forward(O.Ops("np", np.float64, O.Switches(softmax_axis="key", layernorm="token")), W, tokens,
                           n_layers=12)
    print(f"clip12 oracle x2: {time.time() - t0:.1f} s")
    np.savez_compressed(os.path.join(G, "clip12.npz"), tokens=tokens, y_reference_switches=y_ref.astype(np.float32),
                        y_intended_switches=y_int.astype(np.float32))


if __name__ == "__main__":
    {"tokenizer": tokenizer, "small": small, "unet64": unet64, "decoder64": decoder64, "clip": clip, "encoder": encoder,
     "encoder64": encoder64, "loop64": loop64, "decoder64_full": decoder64_full, "clip12": clip12}[sys.argv[1]]()
