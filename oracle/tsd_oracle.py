"""oracle/tsd_oracle.py - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the denoising hot path of lrmantovani10/Stable-Diffusion.mojo under the
semantics contract of SURVEY.md section 0.  PARITY UNPINNED: the reference ships no tests,
golden vectors or fixtures, cannot be compiled in this image (Mojo 24.x, no toolchain, no
network) and as written cannot run at the 64x64-latent shapes at all; the
"intended" switches are cross-checked against torch.nn.functional in tests/test_oracle.py so the
restatement is at least validated by an independent implementation.

Two evaluators of the same algorithm:
  * backend "np"  : numpy, float64 by default (the truth the GPU and the fp32 loops are both
                    measured against), vectorised (im2col + matmul);
  * backend "c32" : oracle/ref_loops.c through ctypes - the reference's own scalar loop nests
                    and fp32 accumulation order (also the timed CPU baseline of bench.py).

Layouts are the reference's: images (C,H,W); token sequences (T,C) (the leading 1 of the
reference's (1,T,C) Matrix is dropped); conv kernels OIHW; linear weights [out][in].
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Reference citations are file:line relative to the reference root.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from synth import ATTN_C, ATTN_LAYER, RES_IN, RES_LAYER, RES_OUT  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB_PATH = os.path.join(_HERE, "_build", "libref_loops.so")


@dataclass
class Switches:
    """SURVEY section 0 switches; defaults are reference-faithful (class D kept)."""
    softmax_axis: str = "query"      # Q3: "query" (reference Softmax(dim=2)) | "key" (standard)
    layernorm: str = "global"        # Q5: "global" (GroupNorm(1,C) over the whole tensor) | "token"
    mojo_alias_time: bool = False    # Q2: SiLU^k(t_emb) aliasing, off by default
    norm_eps_inside: bool = False    # Q6: False = (x-mean)/(std+eps) as the reference; True = (x-mean)/sqrt(var+eps)


# ---------------------------------------------------------------------------------------------
# C backend
# ---------------------------------------------------------------------------------------------
_clib = None


def clib():
    global _clib
    if _clib is None:
        if not os.path.exists(_CLIB_PATH):
            raise RuntimeError(f"{_CLIB_PATH} missing: run `make -C oracle` (or __graft_entry__.build())")
        L = C.CDLL(_CLIB_PATH)
        fp, i, f = C.c_void_p, C.c_int, C.c_float
        L.ref_num_threads.restype = i
        L.ref_set_num_threads.argtypes = [i]
        L.ref_set_num_threads.restype = i
        L.ref_conv2d.argtypes = [fp, i, i, i, fp, fp, i, i, i, i, fp]
        L.ref_matmul.argtypes = [fp, fp, i, i, i, i, fp]
        L.ref_linear.argtypes = [fp, i, i, fp, fp, i, fp]
        L.ref_softmax.argtypes = [fp, i, i, i, i]
        L.ref_groupnorm.argtypes = [fp, i, i, i, f, fp]
        L.ref_silu.argtypes = [fp, C.c_size_t, fp]
        L.ref_gelu.argtypes = [fp, C.c_size_t, fp]
        L.ref_upsample2x.argtypes = [fp, i, i, i, fp]
        L.ref_attention_core.argtypes = [fp, fp, fp, i, i, i, i, i, fp]
        for n in ("ref_conv2d", "ref_matmul", "ref_linear", "ref_softmax", "ref_groupnorm", "ref_silu",
                  "ref_gelu", "ref_upsample2x", "ref_attention_core"):
            getattr(L, n).restype = None
        _clib = L
    return _clib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data


class Ops:
    """Op-level restatement.  backend "np" (dtype float64/float32) or "c32"."""

    def __init__(self, backend: str = "np", dtype=np.float64, sw: Switches | None = None):
        assert backend in ("np", "c32")
        self.backend = backend
        self.dtype = np.float32 if backend == "c32" else dtype
        self.sw = sw or Switches()

    def arr(self, a):
        return np.ascontiguousarray(a, dtype=self.dtype)

    # Conv2D.forward, helpers/utils.mojo:1738-1811 (+pad :1383-1413); Appendix C formula
    def conv2d(self, x, w, b=None, pad=0, stride=1, pad_hi=None):
        cout, cin, k, _ = w.shape
        x = x[:cin]  # the loop reads only the first in_channels planes (utils.mojo:1771, Q9)
        if pad_hi is not None and pad_hi != pad:
            # Matrix.pad((top,bottom),(left,right)), utils.mojo:1383-1413, then an unpadded Conv2D
            x = np.pad(self.arr(x), ((0, 0), (pad, pad_hi), (pad, pad_hi)))
            pad = 0
        _, h, wd = x.shape
        ho, wo = (h + 2 * pad - k) // stride + 1, (wd + 2 * pad - k) // stride + 1
        if self.backend == "c32":
            x, w = _f32(x), _f32(w)
            b32 = None if b is None else _f32(b)
            out = np.empty((cout, ho, wo), np.float32)
            clib().ref_conv2d(_p(x), cin, h, wd, _p(w), _p(b32), cout, k, pad, stride, _p(out))
            return out
        x = self.arr(x)
        xp = np.pad(x, ((0, 0), (pad, pad), (pad, pad))) if pad else x
        win = np.lib.stride_tricks.sliding_window_view(xp, (k, k), axis=(1, 2))[:, ::stride, ::stride]
        # win (cin, ho, wo, k, k) -> (ho*wo, cin*k*k), processed in row chunks to bound memory
        wm = self.arr(w).reshape(cout, cin * k * k)
        out = np.empty((cout, ho * wo), self.dtype)
        rows = max(1, (1 << 26) // max(1, cin * k * k * wo))
        for r0 in range(0, ho, rows):
            r1 = min(ho, r0 + rows)
            col = np.ascontiguousarray(win[:, r0:r1].transpose(1, 2, 0, 3, 4)).reshape((r1 - r0) * wo, cin * k * k)
            out[:, r0 * wo:r1 * wo] = wm @ col.T
        if b is not None:
            out += self.arr(b)[:, None]
        return out.reshape(cout, ho, wo)

    # Linear.forward, helpers/utils.mojo:1954-1976 (bias on every column, Q7); x (T,in)
    def linear(self, x, w, b=None):
        if self.backend == "c32":
            x, w = _f32(x), _f32(w)
            b32 = None if b is None else _f32(b)
            out = np.empty((x.shape[0], w.shape[0]), np.float32)
            clib().ref_linear(_p(x), x.shape[0], x.shape[1], _p(w), _p(b32), w.shape[0], _p(out))
            return out
        y = self.arr(x) @ self.arr(w).T
        if b is not None:
            y = y + self.arr(b)
        return y

    # Matrix.matmul, helpers/utils.mojo:1549-1569
    def matmul(self, a, b):
        if self.backend == "c32":
            a, b = _f32(a), _f32(b)
            c, m, k = a.shape
            n = b.shape[2]
            out = np.empty((c, m, n), np.float32)
            clib().ref_matmul(_p(a), _p(b), c, m, k, n, _p(out))
            return out
        return self.arr(a) @ self.arr(b)

    # GroupNorm.forward, helpers/utils.mojo:1845-1885; (x-mean)/(std+eps), biased std, gamma=1 (Q6)
    def group_norm(self, x, groups, eps=1e-5, gamma=None, beta=None):
        """gamma / beta: optional per-channel vectors (the reference has a scalar gamma = 1 and never adds beta)."""
        c = x.shape[0]
        if c % groups:
            raise ValueError("Number of channels does not evenly divide the number of groups")
        if self.backend == "c32" and gamma is None and beta is None and not self.sw.norm_eps_inside:
            x = _f32(x)
            out = np.empty_like(x)
            clib().ref_groupnorm(_p(x), c, int(np.prod(x.shape[1:])), groups, eps, _p(out))
            return out
        x = self.arr(x)
        g = x.reshape(groups, -1)
        mean = g.mean(axis=1, keepdims=True)
        var = ((g - mean) ** 2).mean(axis=1, keepdims=True)
        den = np.sqrt(var + self.dtype(eps)) if self.sw.norm_eps_inside else np.sqrt(var) + self.dtype(eps)
        out = ((g - mean) / den).reshape(x.shape)
        bshape = (c,) + (1,) * (x.ndim - 1)
        if gamma is not None:
            out = out * self.arr(gamma).reshape(bshape)
        if beta is not None:
            out = out + self.arr(beta).reshape(bshape)
        return out

    # LayerNorm.forward = GroupNorm(1, C) on the (C,T,1) Matrix, helpers/utils.mojo:2052-2061 (Q5).
    # x is token-major (T,C); statistics are layout independent in the default mode.
    def layer_norm(self, x, gamma=None, beta=None):
        if self.sw.layernorm == "global":
            t, c = x.shape
            return self.group_norm(np.ascontiguousarray(x.T).reshape(c, t, 1), 1, 1e-5, gamma, beta).reshape(c, t).T
        x = self.arr(x)
        mean = x.mean(axis=1, keepdims=True)
        var = ((x - mean) ** 2).mean(axis=1, keepdims=True)
        den = np.sqrt(var + self.dtype(1e-5)) if self.sw.norm_eps_inside else np.sqrt(var) + self.dtype(1e-5)
        out = (x - mean) / den
        if gamma is not None:
            out = out * self.arr(gamma)[None, :]
        if beta is not None:
            out = out + self.arr(beta)[None, :]
        return out

    def silu(self, x):  # SiLU.forward, helpers/utils.mojo:1892-1902
        if self.backend == "c32":
            x = _f32(x)
            out = np.empty_like(x)
            clib().ref_silu(_p(x), x.size, _p(out))
            return out
        x = self.arr(x)
        return x / (1 + np.exp(-x))

    def gelu(self, x):  # Gelu.forward, helpers/utils.mojo:1908-1919 (tanh form)
        if self.backend == "c32":
            x = _f32(x)
            out = np.empty_like(x)
            clib().ref_gelu(_p(x), x.size, _p(out))
            return out
        x = self.arr(x)
        return x * 0.5 * (1 + np.tanh(np.sqrt(2 / np.pi) * (x + 0.044715 * x ** 3)))

    def upsample2x(self, x):  # Upsample.forward under Q8: y[c,i,j] = x[c,i//2,j//2]
        return np.repeat(np.repeat(x, 2, axis=1), 2, axis=2)

    # Softmax, helpers/utils.mojo:411-448; reference dim numbering on a (c,r,cols) tensor
    def softmax(self, m, dim=2):
        if dim not in (1, 2):
            raise ValueError("Invalid dimension for softmax")
        if self.backend == "c32":
            m = _f32(m).copy()
            clib().ref_softmax(_p(m), m.shape[0], m.shape[1], m.shape[2], dim)
            return m
        m = self.arr(m)
        axis = 1 if dim == 2 else 2
        e = np.exp(m - m.max(axis=axis, keepdims=True))  # max-subtraction: identical when nothing overflows
        return e / e.sum(axis=axis, keepdims=True)

    # attention core, helpers/attention.mojo:46-62 / 105-115; q (h,Tq,d), k,v (h,Tk,d) -> (Tq, h*d)
    def attention_core(self, q, k, v, causal=False):
        h, tq, d = q.shape
        tk = k.shape[1]
        key_axis = self.sw.softmax_axis == "key"
        if causal and self.backend == "c32":
            raise NotImplementedError("the C restatement covers the per-step path only (no causal mask)")
        if self.backend == "c32":
            q, k, v = _f32(q), _f32(k), _f32(v)
            out = np.empty((tq, h * d), np.float32)
            clib().ref_attention_core(_p(q), _p(k), _p(v), h, tq, tk, d, int(key_axis), _p(out))
            return out
        out = np.empty((tq, h, d), self.dtype)
        for i in range(h):  # one head at a time bounds the T x T score plane
            s = self.arr(q[i]) @ self.arr(k[i]).T
            if causal:
                # Self_Attention.forward, helpers/attention.mojo:48-56: masked_fill(triu(1), -inf) before the
                # 1/sqrt(d) scaling; triu(1) taken as the standard strict upper triangle (class G)
                s = np.where(np.triu(np.ones((tq, tk), bool), 1), -np.inf, s)
            s = s / np.sqrt(self.dtype(d))
            p = self.softmax(s[None], dim=1 if key_axis else 2)[0]
            out[:, i, :] = p @ self.arr(v[i])
        return out.reshape(tq, h * d)

    # Self_Attention.forward, helpers/attention.mojo:26-65; x (T,C)
    def self_attention(self, x, n_heads, w_in, b_in, w_out, b_out, causal=False):
        t, c = x.shape
        qkv = self.linear(x, w_in, b_in)                       # (T,3C)
        q, k, v = (np.ascontiguousarray(qkv[:, i * c:(i + 1) * c]) for i in range(3))   # chunk(2,3) :29
        d = c // n_heads
        q, k, v = (a.reshape(n_heads, t, d) for a in (q, k, v))  # raw reshape, no transpose (Q4) :30-44
        o = self.attention_core(q, k, v, causal=causal)
        return self.linear(o, w_out, b_out)

    def quick_gelu(self, x):  # ClipPlayer.forward, clip.mojo:49-50 (intent: x * sigmoid(1.702 x))
        x = self.arr(x)
        return x / (1.0 + np.exp(-1.702 * x))

    # Cross_Attention.forward, helpers/attention.mojo:96-118; x (T,C), context (Tk,Dc)
    def cross_attention(self, x, context, n_heads, wq, bq, wk, bk, wv, bv, wo, bo):
        t, c = x.shape
        d = c // n_heads
        q = np.ascontiguousarray(self.linear(x, wq, bq)).reshape(n_heads, t, d)
        k = np.ascontiguousarray(self.linear(context, wk, bk)).reshape(n_heads, context.shape[0], d)
        v = np.ascontiguousarray(self.linear(context, wv, bv)).reshape(n_heads, context.shape[0], d)
        o = self.attention_core(q, k, v)
        return self.linear(o, wo, bo)


# ---------------------------------------------------------------------------------------------
# blocks and models
# ---------------------------------------------------------------------------------------------
def clip_forward(ops: Ops, W, tokens, n_embed=768, n_tokens=77, n_heads=12, n_layers=12):
    """CLIP.forward, clip.mojo:88-109 (ClipEmbedding :17-20, ClipPlayer :36-53): token ids -> (77, 768).
    LayerNorm here is ops.layer_norm, i.e. the reference's GroupNorm(1, C) over the whole tensor unless the
    per-token switch is set; the softmax axis follows ops.sw (Q3); the causal mask is the standard one."""
    tok = np.zeros(n_tokens, np.int64)          # reshaped_tokens *= 0 ; set_items(...)  :90-92
    t_in = np.asarray(tokens).reshape(-1)
    tok[:t_in.size] = t_in
    table = W["embedding.token_embedding.weight"].reshape(-1, n_embed)
    pos = W["embedding.position_embedding"].reshape(n_tokens, n_embed)
    x = ops.arr(table[tok]) + ops.arr(pos)                                   # :17-20
    for l in range(1, n_layers + 1):
        b = f"player{l}"
        residue = x
        h = ops.layer_norm(x, *_affine(W, b + ".layer1"))                    # :38-41
        h = ops.self_attention(h, n_heads, W[b + ".layer2.in_proj.weight"], W[b + ".layer2.in_proj.bias"],
                               W[b + ".layer2.out_proj.weight"], W[b + ".layer2.out_proj.bias"], causal=True)  # :42
        x = h + residue                                                      # :43
        residue = x
        h = ops.layer_norm(x, *_affine(W, b + ".layer3"))                    # :45-47
        h = ops.linear(h, W[b + ".layer4.weight"], W[b + ".layer4.bias"])    # :48
        h = ops.quick_gelu(h)                                                # :49-50
        h = ops.linear(h, W[b + ".layer5.weight"], W[b + ".layer5.bias"])    # :51
        x = h + residue                                                      # :52
    return ops.layer_norm(x, *_affine(W, "layernorm"))                       # :106-108



def time_embedding_mlp(ops: Ops, W, t):
    """Time_Embedding.forward, diffusion.mojo:17-21; t (320,) -> (1280,)"""
    h = ops.linear(ops.arr(t)[None, :], W["time_embed.layer1.weight"], W["time_embed.layer1.bias"])
    h = ops.silu(h)
    return ops.linear(h, W["time_embed.layer2.weight"], W["time_embed.layer2.bias"])[0]


def _affine(W, name):
    """(weight, bias) of a norm when the weight set carries them (norm_affine models), else (None, None)."""
    if (name + ".weight") in W:
        return W[name + ".weight"], W[name + ".bias"]
    return None, None


def unet_res_block(ops: Ops, W, base, x, time_act, cin, cout):
    """Unet_Residual_Block.forward, diffusion.mojo:54-72.  time_act = the vector fed to layer3
    (SiLU(t_emb), or SiLU^k under mojo_alias_time)."""
    x = x[:cin]
    out = ops.group_norm(x, 32, 1e-5, *_affine(W, base + ".layer1"))
    out = ops.silu(out)
    out = ops.conv2d(out, W[base + ".layer2.weight"], W[base + ".layer2.bias"], pad=1)
    tb = ops.linear(ops.arr(time_act)[None, :], W[base + ".layer3.weight"], W[base + ".layer3.bias"])[0]
    merged = out + ops.arr(tb)[:, None, None]
    merged = ops.group_norm(merged, 32, 1e-5, *_affine(W, base + ".layer4"))
    merged = ops.silu(merged)
    merged = ops.conv2d(merged, W[base + ".layer5.weight"], W[base + ".layer5.bias"], pad=1)
    if cin != cout:
        return merged + ops.conv2d(x, W[base + ".layer6.weight"], W[base + ".layer6.bias"])
    return merged + ops.arr(x)


def unet_attn_block(ops: Ops, W, base, x, context, n_heads=8):
    """Unet_Attention_Block.forward, diffusion.mojo:112-147; x (C,H,W), context (77,768)."""
    c, h, w = x.shape
    residue_long = ops.arr(x)
    out = ops.group_norm(x, 32, 1e-6, *_affine(W, base + ".layer1"))
    out = ops.conv2d(out, W[base + ".layer2.weight"], W[base + ".layer2.bias"])
    seq = np.ascontiguousarray(out.reshape(c, h * w).T)         # (T,C) token-major view of :118-123
    rs = seq
    seq = ops.layer_norm(seq, *_affine(W, base + ".layer3"))
    seq = ops.self_attention(seq, n_heads, W[base + ".layer4.in_proj.weight"], None,
                             W[base + ".layer4.out_proj.weight"], W[base + ".layer4.out_proj.bias"])
    seq = seq + rs
    rs = seq
    seq = ops.layer_norm(seq, *_affine(W, base + ".layer5"))
    seq = ops.cross_attention(seq, context, n_heads, W[base + ".layer6.q_proj.weight"], None,
                              W[base + ".layer6.k_proj.weight"], None, W[base + ".layer6.v_proj.weight"], None,
                              W[base + ".layer6.out_proj.weight"], W[base + ".layer6.out_proj.bias"])
    seq = seq + rs
    rs = seq
    seq = ops.layer_norm(seq, *_affine(W, base + ".layer7"))
    hcat = ops.linear(seq, W[base + ".layer8.weight"], W[base + ".layer8.bias"])     # (T,8C)
    half = hcat.shape[1] // 2
    seq = hcat[:, :half] * ops.gelu(np.ascontiguousarray(hcat[:, half:]))            # chunk(2,2) :138-141
    seq = ops.linear(seq, W[base + ".layer9.weight"], W[base + ".layer9.bias"])
    seq = seq + rs
    img = np.ascontiguousarray(seq.T).reshape(c, h, w)
    return ops.conv2d(img, W[base + ".layer10.weight"], W[base + ".layer10.bias"]) + residue_long


def diffusion_forward(ops: Ops, W, x, context, time):
    """Diffusion.forward, diffusion.mojo:309-318 (UNet.forward :228-273, output layer :287-291).
    x (4,H,W); context (77,768) [a (2,77,768) CFG stack must be split by the caller, Q13];
    time (320,).  Returns (4,H,W)."""
    temb = time_embedding_mlp(ops, W, time)
    state = {"t": temb}

    def time_act():
        if ops.sw.mojo_alias_time:   # Q2: SiLU().forward(time) overwrites the shared embedding
            state["t"] = ops.silu(state["t"])
            return state["t"]
        return ops.silu(temb)

    def res(i, v):
        return unet_res_block(ops, W, f"unet.layer{RES_LAYER[i]}", v, time_act(), RES_IN[i], RES_OUT[i])

    def att(i, v):
        return unet_attn_block(ops, W, f"unet.layer{ATTN_LAYER[i]}", v, context)

    cat = lambda a, b: np.concatenate([a, b], axis=0)  # noqa: E731  Matrix.concat dim 0, utils.mojo:605-722
    out = ops.conv2d(x, W["unet.layer1.weight"], W["unet.layer1.bias"], pad=1)
    skip1 = out
    out = att(0, res(0, out))
    skip2 = out
    out = ops.conv2d(out, W["unet.layer4.weight"], W["unet.layer4.bias"], pad=1, stride=2)
    skip3 = out
    out = att(1, res(1, out))
    skip4 = out
    out = ops.conv2d(out, W["unet.layer7.weight"], W["unet.layer7.bias"], pad=1, stride=2)
    skip5 = out
    out = att(2, res(2, out))
    skip6 = out
    out = att(3, res(3, cat(out, skip6)))           # Q10: [x;x]
    out = att(4, res(4, cat(out, skip5)))
    out = ops.upsample2x(out)                       # layer14 (Q8)
    out = att(5, res(5, out))                       # layer15 reads only its first 1280 channels: skip4 dead (Q9)
    out = att(6, res(6, cat(out, skip3)))
    out = ops.upsample2x(out)                       # layer19
    out = att(7, res(7, out))                       # skip2 dead (Q9)
    out = att(8, res(8, cat(out, skip1)))
    del skip2, skip4
    out = ops.group_norm(out, 320, 1e-5, *_affine(W, "final.layer1"))   # GroupNorm(320, 320): Q11
    out = ops.silu(out)
    return ops.conv2d(out, W["final.layer2.weight"], W["final.layer2.bias"], pad=1)


def vae_res_block(ops: Ops, W, base, x, cin, cout):
    """Res_Block.forward, vae.mojo:57-67 (GroupNorm(16,.), Q17)."""
    out = ops.group_norm(x, 16, 1e-5, *_affine(W, base + ".groupnorm1"))
    out = ops.silu(out)
    out = ops.conv2d(out, W[base + ".conv1.weight"], W[base + ".conv1.bias"], pad=1)
    out = ops.group_norm(out, 16, 1e-5, *_affine(W, base + ".groupnorm2"))
    out = ops.silu(out)
    out = ops.conv2d(out, W[base + ".conv2.weight"], W[base + ".conv2.bias"], pad=1)
    res = ops.arr(x)
    if cin != cout:
        res = ops.conv2d(x, W[base + ".res_conv_layer.weight"], W[base + ".res_conv_layer.bias"])
    return out + res


def vae_attn_block(ops: Ops, W, x, name="l4"):
    """Attention_Block.forward, vae.mojo:17-27: GroupNorm(32) -> 1-head self-attention -> + residue."""
    c, h, w = x.shape
    out = ops.group_norm(x, 32, 1e-5, *_affine(W, name + ".groupnorm"))
    seq = np.ascontiguousarray(out.reshape(c, h * w).T)
    seq = ops.self_attention(seq, 1, W[name + ".attention.in_proj.weight"], W[name + ".attention.in_proj.bias"],
                             W[name + ".attention.out_proj.weight"], W[name + ".attention.out_proj.bias"])
    return np.ascontiguousarray(seq.T).reshape(c, h, w) + ops.arr(x)


def encoder_forward(ops: Ops, W, x, noise):
    """Encoder.forward, vae.mojo:131-159; x (3,8h,8w) in (-1,1), noise (4,h,w) -> latent (4,h,w).
    two_stride_pad (:115-116) = one zero row below and one zero column right before each stride-2 conv."""
    out = ops.conv2d(ops.arr(x), W["l1.weight"], W["l1.bias"], pad=1)
    out = vae_res_block(ops, W, "l2", out, 128, 128)
    out = vae_res_block(ops, W, "l3", out, 128, 128)
    out = ops.conv2d(out, W["l4.weight"], W["l4.bias"], pad=0, stride=2, pad_hi=1)
    out = vae_res_block(ops, W, "l5", out, 128, 256)
    out = vae_res_block(ops, W, "l6", out, 256, 256)
    out = ops.conv2d(out, W["l7.weight"], W["l7.bias"], pad=0, stride=2, pad_hi=1)
    out = vae_res_block(ops, W, "l8", out, 256, 512)
    out = vae_res_block(ops, W, "l9", out, 512, 512)
    out = ops.conv2d(out, W["l10.weight"], W["l10.bias"], pad=0, stride=2, pad_hi=1)
    for n in ("l11", "l12", "l13"):
        out = vae_res_block(ops, W, n, out, 512, 512)
    out = vae_attn_block(ops, W, out, "l14")
    out = vae_res_block(ops, W, "l15", out, 512, 512)
    out = ops.group_norm(out, 32, 1e-5, *_affine(W, "l16"))
    out = ops.silu(out)
    out = ops.conv2d(out, W["l18.weight"], W["l18.bias"], pad=1)
    out = ops.conv2d(out, W["l19.weight"], W["l19.bias"])
    return latent_from_moments(ops, out, noise)


def latent_from_moments(ops: Ops, moments, noise):
    """Encoder.metrics_evals, vae.mojo:118-129: chunk(0,2) -> mean, log-variance; clamp(-30,20);
    std = sqrt(exp(.)); (mean + noise*std) * 0.18215."""
    m = ops.arr(moments)
    half = m.shape[0] // 2
    mean, logvar = m[:half], np.clip(m[half:], -30.0, 20.0)
    std = np.sqrt(np.exp(logvar))
    return (mean + ops.arr(noise) * std) * ops.dtype(0.18215)


def rescale_input(img):
    """Matrix.rescale((0,255),(-1,1)), helpers/utils.mojo:577-597, pipeline.mojo:71."""
    return np.asarray(img) * 2.0 / 255.0 - 1.0


def resize_image(img, new_h, new_w):
    """resize_image, helpers/utils.mojo:372-402: nearest neighbour, index int(row*old/new).  The
    reference names dim1 "width" and dim2 "height" (:376-377) and so scales rows by dim2/new_h; for
    the square images of the pipeline (image_size x image_size) both readings coincide."""
    img = np.asarray(img)
    c, h, w = img.shape
    if h == new_h and w == new_w:
        return img
    ys = (np.arange(new_h) * (h / new_h)).astype(np.int64)
    xs = (np.arange(new_w) * (w / new_w)).astype(np.int64)
    return img[:, ys][:, :, xs]


def rescale_image(img):
    """Matrix.rescale((-1,1),(0,255),clamp=True), helpers/utils.mojo:577-597, pipeline.mojo:127."""
    return np.clip((img + 1.0) * 255.0 / 2.0, 0.0, 255.0)


def decoder_forward(ops: Ops, W, z, rescale=False):
    """Decoder.forward, vae.mojo:221-250; z (4,h,w) -> (3,8h,8w)."""
    out = ops.arr(z) / ops.dtype(0.18215)
    out = ops.conv2d(out, W["l1.weight"], W["l1.bias"])
    out = ops.conv2d(out, W["l2.weight"], W["l2.bias"], pad=1)
    out = vae_res_block(ops, W, "l3", out, 512, 512)
    out = vae_attn_block(ops, W, out)
    for n in ("l5", "l6", "l7", "l8"):
        out = vae_res_block(ops, W, n, out, 512, 512)
    out = ops.upsample2x(out)
    out = ops.conv2d(out, W["l10.weight"], W["l10.bias"], pad=1)
    for n in ("l11", "l12", "l13"):
        out = vae_res_block(ops, W, n, out, 512, 512)
    out = ops.upsample2x(out)
    out = ops.conv2d(out, W["l15.weight"], W["l15.bias"], pad=1)
    out = vae_res_block(ops, W, "l16", out, 512, 256)
    out = vae_res_block(ops, W, "l17", out, 256, 256)
    out = vae_res_block(ops, W, "l18", out, 256, 256)
    out = ops.upsample2x(out)
    out = ops.conv2d(out, W["l20.weight"], W["l20.bias"], pad=1)
    out = vae_res_block(ops, W, "l21", out, 256, 128)
    out = vae_res_block(ops, W, "l22", out, 128, 128)
    out = vae_res_block(ops, W, "l23", out, 128, 128)
    out = ops.group_norm(out, 32, 1e-5, *_affine(W, "l24"))
    out = ops.silu(out)
    out = ops.conv2d(out, W["l26.weight"], W["l26.bias"], pad=1)
    return rescale_image(out) if rescale else out


# ---------------------------------------------------------------------------------------------
# sampler and loop (host-side scalar math; fp64 here, fp32 scalars reach the kernels)
# ---------------------------------------------------------------------------------------------
def get_time_embedding(timestep: float, as_written: bool = False) -> np.ndarray:
    """get_time_embedding, helpers/utils.mojo:353-370.  Intended: f_i = 10000^(-i/160);
    as written (Q12): f_i = (-i/160)^10000."""
    i = np.arange(160, dtype=np.float64)
    with np.errstate(over="ignore"):
        freqs = np.power(-i / 160.0, 10000.0) if as_written else np.power(10000.0, -i / 160.0)
    x = freqs * float(timestep)
    return np.concatenate([np.cos(x), np.sin(x)]).astype(np.float32)


class DDPMSampler:
    """DDPMSampler, sampler.mojo:5-124, with T = 1000 training steps (Q14)."""

    def __init__(self, num_training_steps=1000, beta_start=0.00085, beta_end=0.0120):
        self.T = num_training_steps
        self.betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_training_steps, dtype=np.float64) ** 2  # :28-30
        self.alphas_cumprod = np.cumprod(1.0 - self.betas)                                                        # :31-32
        self.set_inference_timesteps(1)

    def set_inference_timesteps(self, n):  # :35-44
        self.n = n
        ratio = self.T // n
        self.timesteps = np.round(np.arange(n - 1, -1, -1) * ratio).astype(np.int64)

    def coefficients(self, t):
        """Scalars of step() (:75-109): sqrt_ab, sqrt_1mab, c0 (x0 coef), c1 (x_t coef), sigma."""
        prev = t - self.T // self.n                                   # :46-51
        ab = self.alphas_cumprod[t]
        ab_prev = self.alphas_cumprod[prev] if prev >= 0 else 1.0
        cur_alpha = ab / ab_prev
        cur_beta = 1 - cur_alpha
        c0 = (ab_prev ** 0.5 * cur_beta) / (1 - ab)
        c1 = cur_alpha ** 0.5 * (1 - ab_prev) / (1 - ab)
        sigma = 0.0
        if t > 0:                                                      # :101-108, variance :53-65
            sigma = max((1 - ab_prev) / (1 - ab) * cur_beta, 1e-20) ** 0.5
        return np.array([ab ** 0.5, (1 - ab) ** 0.5, c0, c1, sigma], dtype=np.float64)

    def step(self, t, latents, model_output, noise=None):
        s_ab, s_1mab, c0, c1, sigma = self.coefficients(int(t))
        x0 = (latents - model_output * s_1mab) / s_ab
        out = x0 * c0 + latents * c1
        if t > 0 and noise is not None:
            out = out + noise * sigma
        return out

    def set_strength(self, strength):  # :67-73 (intended slice timesteps[start:])
        self.start_step = self.n - int(self.n * strength)
        self.timesteps = self.timesteps[self.start_step:]

    def add_noise(self, x, t, noise):  # :111-124
        ab = self.alphas_cumprod[int(t)]
        return x * ab ** 0.5 + noise * (1 - ab) ** 0.5


def cfg_combine(cond, uncond, scale):
    """pipeline.mojo:117-119"""
    return (cond - uncond) * scale + uncond


def generate_latents(ops: Ops, W, latents, context, steps, noise=None, cfg_context=None, cfg_scale=7.5,
                     time_as_written=False):
    """Denoising loop, pipeline.mojo:86-122 (intended CFG: two evaluations, Q13).
    latents (4,H,W); context (77,768); cfg_context = uncond context or None; noise (steps,4,H,W)."""
    sm = DDPMSampler()
    sm.set_inference_timesteps(steps)
    lat = ops.arr(latents)
    for i, t in enumerate(sm.timesteps):
        temb = get_time_embedding(float(t), time_as_written)
        eps = diffusion_forward(ops, W, lat, context, temb)
        if cfg_context is not None:
            eps = cfg_combine(eps, diffusion_forward(ops, W, lat, cfg_context, temb), cfg_scale)
        lat = sm.step(int(t), lat, eps, None if noise is None else ops.arr(noise[i]))
    return lat
