"""oracle/synth.py - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Parameter inventories of Diffusion / Decoder in the reference's struct-declaration order
(reference diffusion.mojo:295-297, 151-173, 25-30, 76-85 and
vae.mojo:163-188, 33-37, 6-7) and the CPU twin of the device's synthetic-weight generator
(csrc/models.cu synth_param_kernel): counter-based splitmix64 on the reference-layout element
index, so the oracle and the GPU see bit-identical fp32 weights without moving 1.2 GB around.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _splitmix64_scalar(x: int) -> int:
    return int(_splitmix64(np.array([x & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0])


def synth_tensor(seed: int, index: int, numel: int, scale: np.float32) -> np.ndarray:
    """value(j) = float32(int(z >> 40) - 2^23) * float32(scale * 2^-23), z = splitmix64(pseed + j)."""
    if scale is None:
        return np.ones(numel, np.float32)
    if float(scale) == 0.0:
        return np.zeros(numel, np.float32)
    pseed = _splitmix64_scalar(seed ^ _splitmix64_scalar(index + 1))
    out = np.empty(numel, np.float32)
    step = np.float32(scale) * np.float32(1.0 / 8388608.0)
    chunk = 1 << 24
    for s in range(0, numel, chunk):
        n = min(chunk, numel - s)
        with np.errstate(over="ignore"):
            ctr = (np.uint64(pseed) + np.arange(s, s + n, dtype=np.uint64)) & _M64
        z = _splitmix64(ctr)
        u = (z >> np.uint64(40)).astype(np.int64) - 8388608
        out[s:s + n] = u.astype(np.float32) * step
    return out


def round_tf32(x: np.ndarray) -> np.ndarray:
    """cvt.rna.tf32.f32: round to 10 mantissa bits, ties away from zero (finite inputs)."""
    b = np.ascontiguousarray(x, np.float32).view(np.uint32)
    b = (b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return b.view(np.float32)


# ---- inventories ---------------------------------------------------------------------------
RES_IN = [320, 320, 640, 2560, 1920, 1280, 960, 640, 640]
RES_OUT = [320, 640, 1280, 1280, 1280, 640, 640, 320, 320]
RES_LAYER = [2, 5, 8, 10, 12, 15, 17, 20, 22]
ATTN_C = [320, 640, 1280, 1280, 1280, 640, 640, 320, 320]
ATTN_LAYER = [3, 6, 9, 11, 13, 16, 18, 21, 23]


def _conv(specs, name, cin, cout, k):
    s = np.float32(1.0) / np.sqrt(np.float32(cin * k * k))
    specs.append((name + ".weight", (cout, cin, k, k), s))
    specs.append((name + ".bias", (cout,), np.float32(0.0)))


def _linear(specs, name, fin, fout, bias):
    s = np.float32(1.0) / np.sqrt(np.float32(fin))
    specs.append((name + ".weight", (fout, fin), s))
    if bias:
        specs.append((name + ".bias", (fout,), s))


def _norm(specs, name, c):
    # init scale None = the constant 1 (norm weight); the bias is 0
    specs.append((name + ".weight", (c,), None))
    specs.append((name + ".bias", (c,), np.float32(0.0)))


def diffusion_specs(context_dim: int = 768, norm_affine: bool = False):
    """[(name, reference-layout shape, init scale)] in blob order.  norm_affine: the per-channel weight / bias of every
    GroupNorm / LayerNorm, in struct order (tsd_diffusion_config.norm_affine)."""
    sp = []
    _linear(sp, "time_embed.layer1", 320, 1280, True)
    _linear(sp, "time_embed.layer2", 1280, 1280, True)
    for layer in range(1, 24):
        base = f"unet.layer{layer}"
        if layer == 1:
            _conv(sp, base, 4, 320, 3)
        elif layer == 4:
            _conv(sp, base, 320, 320, 3)
        elif layer == 7:
            _conv(sp, base, 640, 640, 3)
        elif layer in (14, 19):
            pass
        elif layer in RES_LAYER:
            i = RES_LAYER.index(layer)
            cin, cout = RES_IN[i], RES_OUT[i]
            if norm_affine:
                _norm(sp, base + ".layer1", cin)
            _conv(sp, base + ".layer2", cin, cout, 3)
            _linear(sp, base + ".layer3", 1280, cout, True)
            if norm_affine:
                _norm(sp, base + ".layer4", cout)
            _conv(sp, base + ".layer5", cout, cout, 3)
            if cin != cout:
                _conv(sp, base + ".layer6", cin, cout, 1)
        else:
            c = ATTN_C[ATTN_LAYER.index(layer)]
            if norm_affine:
                _norm(sp, base + ".layer1", c)
            _conv(sp, base + ".layer2", c, c, 1)
            if norm_affine:
                _norm(sp, base + ".layer3", c)
            _linear(sp, base + ".layer4.in_proj", c, 3 * c, False)
            _linear(sp, base + ".layer4.out_proj", c, c, True)
            if norm_affine:
                _norm(sp, base + ".layer5", c)
            _linear(sp, base + ".layer6.q_proj", c, c, False)
            _linear(sp, base + ".layer6.k_proj", context_dim, c, False)
            _linear(sp, base + ".layer6.v_proj", context_dim, c, False)
            _linear(sp, base + ".layer6.out_proj", c, c, True)
            if norm_affine:
                _norm(sp, base + ".layer7", c)
            _linear(sp, base + ".layer8", c, 8 * c, True)
            _linear(sp, base + ".layer9", 4 * c, c, True)
            _conv(sp, base + ".layer10", c, c, 1)
    if norm_affine:
        _norm(sp, "final.layer1", 320)
    _conv(sp, "final.layer2", 320, 4, 3)
    return sp


def decoder_specs(norm_affine: bool = False):
    """norm_affine: per-channel weight / bias of every GroupNorm (TSD_MODEL_NORM_AFFINE order, include/tsd_b200.h)."""
    sp = []

    def res(name, cin, cout):
        _conv(sp, name + ".conv1", cin, cout, 3)
        _conv(sp, name + ".conv2", cout, cout, 3)
        if cin != cout:
            _conv(sp, name + ".res_conv_layer", cin, cout, 1)
        if norm_affine:
            _norm(sp, name + ".groupnorm1", cin)
            _norm(sp, name + ".groupnorm2", cout)

    _conv(sp, "l1", 4, 4, 1)
    _conv(sp, "l2", 4, 512, 3)
    res("l3", 512, 512)
    _linear(sp, "l4.attention.in_proj", 512, 1536, True)
    _linear(sp, "l4.attention.out_proj", 512, 512, True)
    if norm_affine:
        _norm(sp, "l4.groupnorm", 512)
    for n in ("l5", "l6", "l7", "l8"):
        res(n, 512, 512)
    _conv(sp, "l10", 512, 512, 3)
    for n in ("l11", "l12", "l13"):
        res(n, 512, 512)
    _conv(sp, "l15", 512, 512, 3)
    res("l16", 512, 256)
    res("l17", 256, 256)
    res("l18", 256, 256)
    _conv(sp, "l20", 256, 256, 3)
    res("l21", 256, 128)
    res("l22", 128, 128)
    res("l23", 128, 128)
    if norm_affine:
        _norm(sp, "l24", 128)
    _conv(sp, "l26", 128, 3, 3)
    return sp


def encoder_specs(norm_affine: bool = False):
    """VAE Encoder (vae.mojo:71-112) in struct-declaration order."""
    sp = []

    def res(name, cin, cout):
        _conv(sp, name + ".conv1", cin, cout, 3)
        _conv(sp, name + ".conv2", cout, cout, 3)
        if cin != cout:
            _conv(sp, name + ".res_conv_layer", cin, cout, 1)
        if norm_affine:
            _norm(sp, name + ".groupnorm1", cin)
            _norm(sp, name + ".groupnorm2", cout)

    _conv(sp, "l1", 3, 128, 3)
    res("l2", 128, 128)
    res("l3", 128, 128)
    _conv(sp, "l4", 128, 128, 3)
    res("l5", 128, 256)
    res("l6", 256, 256)
    _conv(sp, "l7", 256, 256, 3)
    res("l8", 256, 512)
    res("l9", 512, 512)
    _conv(sp, "l10", 512, 512, 3)
    for n in ("l11", "l12", "l13"):
        res(n, 512, 512)
    _linear(sp, "l14.attention.in_proj", 512, 1536, True)
    _linear(sp, "l14.attention.out_proj", 512, 512, True)
    if norm_affine:
        _norm(sp, "l14.groupnorm", 512)
    res("l15", 512, 512)
    if norm_affine:
        _norm(sp, "l16", 512)
    _conv(sp, "l18", 512, 8, 3)
    _conv(sp, "l19", 8, 8, 1)
    return sp


def clip_specs(n_vocab: int = 49408, n_layers: int = 12, n_embed: int = 768, n_tokens: int = 77,
               norm_affine: bool = False):
    """CLIP text encoder (clip.mojo:5-15, 23-34, 56-87) in struct-declaration order.  The embedding
    table and the position embedding are stored flat (device kind P_VEC); LayerNorm owns no tensor."""
    sp = [("embedding.token_embedding.weight", (n_vocab * n_embed,), np.float32(1.0)),
          ("embedding.position_embedding", (n_tokens * n_embed,), np.float32(0.0))]
    for l in range(1, n_layers + 1):
        b = f"player{l}"
        _linear(sp, b + ".layer2.in_proj", n_embed, 3 * n_embed, True)
        _linear(sp, b + ".layer2.out_proj", n_embed, n_embed, True)
        _linear(sp, b + ".layer4", n_embed, 4 * n_embed, True)
        _linear(sp, b + ".layer5", 4 * n_embed, n_embed, True)
        if norm_affine:
            _norm(sp, b + ".layer1", n_embed)
            _norm(sp, b + ".layer3", n_embed)
    if norm_affine:
        _norm(sp, "layernorm", n_embed)
    return sp


def num_params(specs) -> int:
    return int(sum(int(np.prod(s)) for _, s, _ in specs))


class SynthWeights:
    """Lazy name -> fp32 array (reference layout) view of tsd_*_init_random(seed)."""

    def __init__(self, specs, seed: int, bias_seed: int | None = None):
        self.specs = specs
        self.seed = seed
        self.index = {name: i for i, (name, _, _) in enumerate(specs)}
        self._cache = {}

    def __contains__(self, name):
        return name in self.index

    def __getitem__(self, name) -> np.ndarray:
        if name not in self._cache:
            i = self.index[name]
            _, shape, scale = self.specs[i]
            self._cache[name] = synth_tensor(self.seed, i, int(np.prod(shape)), scale).reshape(shape)
        return self._cache[name]

    def drop(self, name):
        self._cache.pop(name, None)

    def blob(self) -> np.ndarray:
        out = np.empty(num_params(self.specs), np.float32)
        off = 0
        for i, (name, shape, scale) in enumerate(self.specs):
            n = int(np.prod(shape))
            out[off:off + n] = synth_tensor(self.seed, i, n, scale)
            off += n
        return out


class BlobWeights:
    """name -> array view over a flat reference-order blob (tsd_*_load_weights layout)."""

    def __init__(self, specs, blob: np.ndarray):
        self.specs = specs
        self.map = {}
        off = 0
        for name, shape, _ in specs:
            n = int(np.prod(shape))
            self.map[name] = blob[off:off + n].reshape(shape)
            off += n
        assert off == blob.size

    def __contains__(self, name):
        return name in self.map

    def __getitem__(self, name):
        return self.map[name]

    def drop(self, name):
        pass


def random_blob(specs, seed: int, bias_scale: float = 1.0) -> np.ndarray:
    """A dense random blob for loader tests: like the synthetic init but with non-zero conv
    biases so that every bias path is exercised."""
    rng = np.random.default_rng(seed)
    parts = []
    for name, shape, scale in specs:
        n = int(np.prod(shape))
        if scale is None:   # norm weight: around 1
            parts.append((1.0 + 0.3 * (rng.random(n, dtype=np.float32) * 2 - 1)).astype(np.float32))
            continue
        s = float(scale)
        if s == 0.0:
            s = 0.05 * bias_scale
        parts.append(((rng.random(n, dtype=np.float32) * 2 - 1) * np.float32(s)).astype(np.float32))
    return np.concatenate(parts)
