"""numpy-facing wrapper over the C ABI: one method per entry point of include/tsd_b200.h.

Arrays go in and come out in the reference's layouts (Matrix (dim0,dim1,dim2) row-major:
images (C,H,W), sequences (1,T,C); conv kernels OIHW; linear weights [out][in]).
Status codes become TsdError; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import TsdError

MODEL_NORM_AFFINE = 1  # TSD_MODEL_NORM_AFFINE (include/tsd_b200.h)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data


class Context:
    """tsd_ctx: one device, one stream, one workspace (tsd_init / tsd_shutdown)."""

    def __init__(self, device: int = 0):
        self.L = _lib.lib()
        h = C.c_void_p()
        rc = self.L.tsd_init(device, C.byref(h))
        if rc != 0:
            raise TsdError(rc, self.L.tsd_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.tsd_shutdown(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc != 0:
            raise TsdError(rc, self.L.tsd_last_error(self.h).decode())

    def set_option(self, name: str, value: int):
        self._ck(self.L.tsd_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        v = C.c_int32()
        self._ck(self.L.tsd_get_option(self.h, name.encode(), C.byref(v)))
        return v.value

    def synchronize(self):
        self._ck(self.L.tsd_synchronize(self.h))

    def launch_count(self) -> int:
        return int(self.L.tsd_launch_count(self.h))

    def timer_start(self):
        self._ck(self.L.tsd_timer_start(self.h))

    def timer_stop(self) -> float:
        """Device milliseconds since timer_start (CUDA events on the context's stream)."""
        ms = C.c_double()
        self._ck(self.L.tsd_timer_stop(self.h, C.byref(ms)))
        return ms.value

    # -- ops ----------------------------------------------------------------------------------
    def conv2d(self, x, weight, bias=None, pad=0, stride=1, pad_hi=None):
        """Conv2D.forward; pad_hi != None: Matrix.pad((pad,pad_hi),(pad,pad_hi)) first (vae.mojo:115-116)."""
        x = _f32(x)
        squeeze = x.ndim == 3
        if squeeze:
            x = x[None]
        weight = _f32(weight)
        n, cin, h, w = x.shape
        cout, cin2, k, k2 = weight.shape
        if cin2 != cin or k != k2:
            raise TsdError(1, "conv2d: weight shape does not match input channels")
        ph = pad if pad_hi is None else pad_hi
        ho, wo = (h + pad + ph - k) // stride + 1, (w + pad + ph - k) // stride + 1
        out = np.empty((n, cout, max(ho, 0), max(wo, 0)), np.float32)
        b = None if bias is None else _f32(bias)
        self._ck(self.L.tsd_conv2d_pad(self.h, _p(x), n, cin, h, w, _p(weight), _p(b), cout, k, pad, ph, stride,
                                       _p(out)))
        return out[0] if squeeze else out

    def linear(self, x, weight, bias=None):
        x = _f32(x)
        weight = _f32(weight)
        shp = x.shape
        in_f = shp[-1]
        rows = int(np.prod(shp[:-1]))
        out_f = weight.shape[0]
        if weight.shape[1] != in_f:
            raise TsdError(1, "linear: invalid input dimensions")
        out = np.empty(shp[:-1] + (out_f,), np.float32)
        b = None if bias is None else _f32(bias)
        self._ck(self.L.tsd_linear(self.h, _p(x), 1, rows, in_f, _p(weight), _p(b), out_f, _p(out)))
        return out

    def matmul(self, a, b):
        a, b = _f32(a), _f32(b)
        c, m, k = a.shape
        c2, k2, n = b.shape
        if c != c2 or k != k2:
            raise TsdError(1, "matmul: non-matching dimensions")
        out = np.empty((c, m, n), np.float32)
        self._ck(self.L.tsd_matmul(self.h, _p(a), _p(b), c, m, k, n, _p(out)))
        return out

    def groupnorm(self, x, groups, eps=1e-5, gamma=None, beta=None):
        x = _f32(x)
        squeeze = x.ndim == 3
        if squeeze:
            x = x[None]
        n, c, h, w = x.shape
        out = np.empty_like(x)
        g = None if gamma is None else _f32(gamma)
        b = None if beta is None else _f32(beta)
        self._ck(self.L.tsd_groupnorm(self.h, _p(x), n, c, h, w, groups, eps, _p(g), _p(b), _p(out)))
        return out[0] if squeeze else out

    def layernorm(self, x):
        x = _f32(x)  # (C, T, 1)
        c, t = x.shape[0], x.shape[1]
        out = np.empty_like(x)
        self._ck(self.L.tsd_layernorm(self.h, _p(x), c, t, _p(out)))
        return out

    def silu(self, x):
        x = _f32(x)
        out = np.empty_like(x)
        self._ck(self.L.tsd_silu(self.h, _p(x), x.size, _p(out)))
        return out

    def gelu(self, x):
        x = _f32(x)
        out = np.empty_like(x)
        self._ck(self.L.tsd_gelu(self.h, _p(x), x.size, _p(out)))
        return out

    def upsample2x(self, x):
        x = _f32(x)
        c, h, w = x.shape
        out = np.empty((c, 2 * h, 2 * w), np.float32)
        self._ck(self.L.tsd_upsample2x(self.h, _p(x), c, h, w, _p(out)))
        return out

    def softmax(self, x, dim=2):
        x = _f32(x)
        c, r, cc = x.shape
        out = np.empty_like(x)
        self._ck(self.L.tsd_softmax(self.h, _p(x), c, r, cc, dim, _p(out)))
        return out

    def attention_core(self, q, k, v):
        q, k, v = _f32(q), _f32(k), _f32(v)
        h, tq, d = q.shape
        tk = k.shape[1]
        out = np.empty((tq, h * d), np.float32)
        self._ck(self.L.tsd_attention_core(self.h, _p(q), _p(k), _p(v), h, tq, tk, d, _p(out)))
        return out

    def self_attention(self, x, n_heads, w_in, b_in, w_out, b_out):
        x = _f32(x)
        t, c = x.shape[-2], x.shape[-1]
        w_in, w_out = _f32(w_in), _f32(w_out)
        b_in = None if b_in is None else _f32(b_in)
        b_out = None if b_out is None else _f32(b_out)
        out = np.empty_like(x)
        self._ck(self.L.tsd_self_attention(self.h, _p(x), t, c, n_heads, _p(w_in), _p(b_in), _p(w_out),
                                           _p(b_out), _p(out)))
        return out

    def cross_attention(self, x, context, n_heads, wq, bq, wk, bk, wv, bv, wo, bo):
        x, context = _f32(x), _f32(context)
        t, c = x.shape[-2], x.shape[-1]
        tk, dc = context.shape[-2], context.shape[-1]
        arrs = [None if a is None else _f32(a) for a in (wq, bq, wk, bk, wv, bv, wo, bo)]
        out = np.empty_like(x)
        self._ck(self.L.tsd_cross_attention(self.h, _p(x), t, c, _p(context), tk, dc, n_heads,
                                            *[_p(a) for a in arrs], _p(out)))
        return out

    def sampler_step(self, latents, eps_cond, eps_uncond, cfg_scale, noise, sqrt_ab, sqrt_1mab, c0, c1,
                     sigma):
        latents, eps_cond = _f32(latents), _f32(eps_cond)
        eu = None if eps_uncond is None else _f32(eps_uncond)
        nz = None if noise is None else _f32(noise)
        out = np.empty_like(latents)
        self._ck(self.L.tsd_sampler_step(self.h, _p(latents), _p(eps_cond), _p(eu), cfg_scale, _p(nz),
                                         sqrt_ab, sqrt_1mab, c0, c1, sigma, latents.size, _p(out)))
        return out

    def sampler_add_noise(self, x, noise, sqrt_ab, sqrt_1mab):
        """DDPMSampler.add_noise (sampler.mojo:111-124) with the host sampler's add_noise_coefficients(t)."""
        x, noise = _f32(x), _f32(noise)
        if x.shape != noise.shape:
            raise TsdError(1, "add_noise: shapes differ")
        out = np.empty_like(x)
        self._ck(self.L.tsd_sampler_add_noise(self.h, _p(x), _p(noise), float(sqrt_ab), float(sqrt_1mab), x.size,
                                              _p(out)))
        return out

    def sampler_step_dev(self, latents_dev, eps_dev, eps_uncond_dev, cfg_scale, noise_dev, coef, n, out_dev):
        """tsd_sampler_step_dev on raw device addresses (ints); coef = the 5 schedule scalars."""
        self._ck(self.L.tsd_sampler_step_dev(self.h, latents_dev, eps_dev, eps_uncond_dev, cfg_scale, noise_dev,
                                             *[float(v) for v in coef], n, out_dev))

    # -- tuning probes --------------------------------------------------------------------------
    def bench_gemm(self, m, n, k, batch=1, geglu=0, force_bn=0, force_splits=0, iters=20) -> float:
        ms = C.c_double()
        self._ck(self.L.tsd_bench_gemm(self.h, m, n, k, batch, geglu, force_bn, force_splits, iters,
                                       C.byref(ms)))
        return ms.value

    def bench_conv(self, n, h, w, cin, cout, k=3, stride=1, force_bn=0, force_splits=0, iters=20) -> float:
        ms = C.c_double()
        self._ck(self.L.tsd_bench_conv(self.h, n, h, w, cin, cout, k, stride, force_bn, force_splits,
                                       iters, C.byref(ms)))
        return ms.value


class _Model:
    """Shared plumbing of the two model handles (parameter blob I/O)."""
    _prefix = ""

    def _fn(self, name):
        return getattr(self.ctx.L, f"tsd_{self._prefix}_{name}")

    def num_params(self) -> int:
        return int(self._fn("num_params")(self.m))

    def param_table(self):
        """[(name, offset, numel)] in blob order (reference struct-declaration order)."""
        out = []
        for i in range(self._fn("param_count")(self.m)):
            off, n = C.c_int64(), C.c_int64()
            name = self._fn("param_name")(self.m, i, C.byref(off), C.byref(n))
            out.append((name.decode(), off.value, n.value))
        return out

    def load_weights(self, blob):
        blob = _f32(blob).reshape(-1)
        self.ctx._ck(self._fn("load_weights")(self.m, _p(blob), blob.size))

    def init_random(self, seed: int):
        self.ctx._ck(self._fn("init_random")(self.m, C.c_uint64(seed)))

    def get_param(self, i: int) -> np.ndarray:
        _, _, n = self.param_table()[i]
        out = np.empty(n, np.float32)
        self.ctx._ck(self._fn("get_param")(self.m, i, _p(out)))
        return out

    def close(self):
        if getattr(self, "m", None) and getattr(self.ctx, "h", None):
            self._fn("destroy")(self.m)
        self.m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Diffusion(_Model):
    """tsd_diffusion: Diffusion (diffusion.mojo:294-318) = Time_Embedding + UNet + output layer."""
    _prefix = "diffusion"

    def __init__(self, ctx: Context, latent_h=64, latent_w=64, max_batch=1, context_len=77, context_dim=768,
                 mojo_alias_time=False, norm_affine=False):
        """norm_affine: every GroupNorm / LayerNorm carries a per-channel weight and bias (what a real checkpoint needs;
        the reference's norm structs own no tensors, so the default is off)."""
        self.ctx = ctx
        self.cfg = _lib.DiffusionConfig(latent_h, latent_w, max_batch, context_len, context_dim,
                                        int(mojo_alias_time), int(norm_affine))
        m = C.c_void_p()
        ctx._ck(ctx.L.tsd_diffusion_create(ctx.h, C.byref(self.cfg), C.byref(m)))
        self.m = m

    def forward(self, x, context, time):
        """Diffusion.forward(x, context, time): x (n,4,H,W) or (4,H,W); context (n_ctx,77,768) or
        (77,768); time (n_time,320) or (320,) -> same shape as x."""
        x = _f32(x)
        squeeze = x.ndim == 3
        if squeeze:
            x = x[None]
        context = _f32(context)
        if context.ndim == 2:
            context = context[None]
        time = _f32(time).reshape(-1, 320)
        out = np.empty_like(x)
        self.ctx._ck(self.ctx.L.tsd_diffusion_forward(self.m, _p(x), _p(context), context.shape[0], _p(time),
                                                      time.shape[0], x.shape[0], _p(out)))
        return out[0] if squeeze else out

    def step(self, latents, context, time, coef, noise=None, uncond_context=None, cfg_scale=7.5):
        """tsd_diffusion_step: one iteration of the reference loop (pipeline.mojo:107-121) in one call - forward
        (x2 with CFG), combine, DDPMSampler.step.  latents (n,4,H,W); context (77,768) or (n,77,768); coef = the five
        schedule scalars of the step; noise like latents or None; returns the new latents."""
        latents = _f32(latents)
        squeeze = latents.ndim == 3
        if squeeze:
            latents = latents[None]
            noise = None if noise is None else _f32(noise)[None]
        context = _f32(context)
        if context.ndim == 2:
            context = context[None]
        cfg = uncond_context is not None
        if cfg:
            u = _f32(uncond_context)
            context = np.ascontiguousarray(np.concatenate([context, u[None] if u.ndim == 2 else u], axis=0))
        nz = None if noise is None else _f32(noise)
        out = np.empty_like(latents)
        self.ctx._ck(self.ctx.L.tsd_diffusion_step(self.m, _p(latents), _p(context), context.shape[0],
                                                   _p(_f32(time).reshape(320)), _p(nz), int(cfg), float(cfg_scale),
                                                   *[float(v) for v in coef], latents.shape[0], _p(out)))
        return out[0] if squeeze else out

    def forward_dev(self, x_dev, ctx_dev, n_ctx, time_dev, n_time, n, out_dev):
        """tsd_diffusion_forward_dev on raw device addresses (ints), asynchronous on the context's
        stream.  ctx_dev = None reuses the context (and its hoisted K/V projections) of the last call."""
        self.ctx._ck(self.ctx.L.tsd_diffusion_forward_dev(self.m, x_dev, ctx_dev, n_ctx, time_dev, n_time, n,
                                                          out_dev))

    def profile(self, x_dev, ctx_dev, n_ctx, time_dev, n_time, n, out_dev):
        """Eager forward on device pointers with per-launch CUDA events; returns per-family
        (ms, flops, launches) for families gemm/conv, attention, norm, other."""
        ms = (C.c_double * 4)()
        fl = (C.c_double * 4)()
        ln = (C.c_int64 * 4)()
        self.ctx._ck(self.ctx.L.tsd_diffusion_profile(self.m, x_dev, ctx_dev, n_ctx, time_dev, n_time, n, out_dev,
                                                      ms, fl, ln))
        fam = ("gemm", "attention", "norm", "other")
        return {f: dict(ms=ms[i], flops=fl[i], launches=ln[i]) for i, f in enumerate(fam)}

    def generate_latents(self, latents, context, timesteps, time_emb, coef, noise=None, cfg=False,
                         cfg_scale=7.5):
        """Whole denoising loop (pipeline.mojo:86-122) on the device.  latents (n,4,H,W);
        context rows: cond first, then (cfg) uncond; coef (steps,5); noise (steps,n,4,H,W)."""
        latents = _f32(latents)
        n = latents.shape[0]
        context = _f32(context)
        timesteps = np.ascontiguousarray(timesteps, np.int32)
        time_emb = _f32(time_emb)
        coef = _f32(coef)
        nz = None if noise is None else _f32(noise)
        lp = _lib.LoopParams(len(timesteps), int(cfg), float(cfg_scale),
                             timesteps.ctypes.data_as(_lib.c_i32_p), time_emb.ctypes.data_as(_lib.c_float_p),
                             coef.ctypes.data_as(_lib.c_float_p),
                             None if nz is None else nz.ctypes.data_as(_lib.c_float_p))
        out = np.empty_like(latents)
        self.ctx._ck(self.ctx.L.tsd_generate_latents(self.m, C.byref(lp), _p(latents), _p(context),
                                                     context.shape[0], n, _p(out)))
        return out


class Clip(_Model):
    """tsd_clip: CLIP text encoder (clip.mojo:56-109): token ids -> the (77, 768) context."""
    _prefix = "clip"

    def __init__(self, ctx: Context, n_vocab: int = 0, n_layers: int = 0, norm_affine=False):
        self.ctx = ctx
        m = C.c_void_p()
        ctx._ck(ctx.L.tsd_clip_create_ex(ctx.h, int(n_vocab), int(n_layers), MODEL_NORM_AFFINE if norm_affine else 0,
                                         C.byref(m)))
        self.m = m

    def forward(self, tokens) -> np.ndarray:
        """CLIP.forward (clip.mojo:88-109): up to 77 token ids (zero-padded) -> (77, 768)."""
        tok = np.ascontiguousarray(np.asarray(tokens).reshape(-1), np.int32)
        out = np.empty((77, 768), np.float32)
        self.ctx._ck(self.ctx.L.tsd_clip_forward(self.m, tok.ctypes.data_as(C.c_void_p), tok.size, _p(out)))
        return out


class Decoder(_Model):
    """tsd_decoder: VAE Decoder (vae.mojo:162-250)."""
    _prefix = "decoder"

    def __init__(self, ctx: Context, latent_h=64, latent_w=64, max_batch=1, norm_affine=False):
        self.ctx = ctx
        self.shape = (latent_h, latent_w)
        m = C.c_void_p()
        ctx._ck(ctx.L.tsd_decoder_create_ex(ctx.h, latent_h, latent_w, max_batch, MODEL_NORM_AFFINE if norm_affine else 0,
                                            C.byref(m)))
        self.m = m

    def forward(self, z, rescale=False):
        z = _f32(z)
        squeeze = z.ndim == 3
        if squeeze:
            z = z[None]
        n, _, h, w = z.shape
        img = np.empty((n, 3, 8 * h, 8 * w), np.float32)
        self.ctx._ck(self.ctx.L.tsd_decoder_forward(self.m, _p(z), n, int(rescale), _p(img)))
        return img[0] if squeeze else img

    def forward_dev(self, z_dev, n, rescale, img_dev):
        self.ctx._ck(self.ctx.L.tsd_decoder_forward_dev(self.m, z_dev, n, int(rescale), img_dev))


class Encoder(_Model):
    """tsd_encoder: VAE Encoder (vae.mojo:70-159).  forward(x, noise) as Encoder.forward (:131-159):
    x (3,8h,8w) [or (n,3,8h,8w)] in (-1,1) - or in (0,255) with rescale=True, the pipeline's
    rescale((0,255),(-1,1)) (pipeline.mojo:71) - and noise (4,h,w) -> latent (4,h,w)."""
    _prefix = "encoder"

    def __init__(self, ctx: Context, latent_h=64, latent_w=64, max_batch=1, norm_affine=False):
        self.ctx = ctx
        self.shape = (latent_h, latent_w)
        m = C.c_void_p()
        ctx._ck(ctx.L.tsd_encoder_create_ex(ctx.h, latent_h, latent_w, max_batch, MODEL_NORM_AFFINE if norm_affine else 0,
                                            C.byref(m)))
        self.m = m

    def forward(self, x, noise, rescale=False):
        x, noise = _f32(x), _f32(noise)
        squeeze = x.ndim == 3
        if squeeze:
            x, noise = x[None], noise[None]
        n, ch, h, w = x.shape
        lh, lw = self.shape
        if ch != 3 or (h, w) != (8 * lh, 8 * lw) or noise.shape != (n, 4, lh, lw):
            raise TsdError(1, "encoder: image must be (3,8h,8w) and noise (4,h,w) for the latent size of the model")
        z = np.empty((n, 4, lh, lw), np.float32)
        self.ctx._ck(self.ctx.L.tsd_encoder_forward(self.m, _p(x), _p(noise), n, int(rescale), _p(z)))
        return z[0] if squeeze else z

    def forward_dev(self, img_dev, noise_dev, n, rescale, z_dev):
        self.ctx._ck(self.ctx.L.tsd_encoder_forward_dev(self.m, img_dev, noise_dev, n, int(rescale), z_dev))
