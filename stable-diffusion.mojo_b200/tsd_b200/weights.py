"""Checkpoint I/O (SURVEY section 8 row f2 - the reference's own TODO, README.md:44,55: its weights are random).

* `SafeTensors`: reader over the C ABI (csrc/host_io.cu): names, shapes, tensors as fp32 (F32 / F16 / BF16 / F64).
* `write_safetensors`: pure-Python writer (tests, export).
* `build_blob`: the flat fp32 blob `load_weights` expects, assembled in the model's parameter order from a
  checkpoint and a name map {parameter name of this library: tensor name in the file}.
* `export_model` / `import_model`: a model's parameters to / from a safetensors file under this library's names.

* `tiny_sd_unet_name_map`: this library's UNet parameter names -> diffusers-convention tensor names (the layout
  `segmind/tiny-sd` ships in: `down_blocks.N.resnets.M.conv1.weight`, `...attentions.M.transformer_blocks.0.attn1.to_q.weight`,
  ...), for a model created with `norm_affine=True` (a real checkpoint's GroupNorm / LayerNorm weights need a home; the
  reference's norm structs own no tensors, helpers/utils.mojo:1825-1872, 2052-2061).  A fused tensor of this library
  (self-attention `in_proj` = to_q | to_k | to_v) maps to a LIST of file tensors concatenated along their first axis.
  The map follows the reference's 23-layer topology (diffusion.mojo:175-201); no checkpoint exists offline, so it is
  tested against a synthetic file with those names, not against the published weights - `missing` / `unused` in the
  report of `build_blob` show what a real file does not cover (e.g. the upsamplers' convolutions, which the reference's
  `Upsample` does not have).  Run such a model with softmax_axis = 1, layernorm_mode = 1, norm_eps_mode = 1."""
from __future__ import annotations

import ctypes as C
import json
import struct

import numpy as np

from . import _lib
from ._lib import TsdError


class SafeTensors:
    def __init__(self, source):
        self.L = _lib.lib()
        h = C.c_void_p()
        if isinstance(source, (bytes, bytearray, memoryview)):
            buf = bytes(source)
            rc = self.L.tsd_safetensors_from_memory(buf, len(buf), C.byref(h))
        else:
            rc = self.L.tsd_safetensors_open(str(source).encode(), C.byref(h))
        if rc:
            raise TsdError(rc, "safetensors: missing or malformed file")
        self.h = h
        self._index = {self.L.tsd_safetensors_name(h, i).decode(): i for i in range(self.L.tsd_safetensors_count(h))}

    def names(self):
        return list(self._index)

    def __contains__(self, name):
        return name in self._index

    def info(self, name):
        """(dtype string, shape tuple)"""
        i = self._index[name]
        dt = C.create_string_buffer(8)
        rank, numel = C.c_int32(), C.c_int64()
        shape = (C.c_int64 * 8)()
        rc = self.L.tsd_safetensors_info(self.h, i, dt, C.byref(rank), shape, C.byref(numel))
        if rc:
            raise TsdError(rc, f"safetensors: bad tensor entry {name}")
        return dt.value.decode(), tuple(shape[:rank.value])

    def read(self, name) -> np.ndarray:
        """The tensor as fp32 in its stored shape (F16 / BF16 widen exactly)."""
        dtype, shape = self.info(name)
        out = np.empty(shape, np.float32)
        rc = self.L.tsd_safetensors_read_f32(self.h, self._index[name], out.ctypes.data, out.size)
        if rc:
            raise TsdError(rc, f"safetensors: tensor {name} has dtype {dtype}, not a floating-point parameter")
        return out

    def close(self):
        if getattr(self, "h", None):
            self.L.tsd_safetensors_close(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _bf16_bits(a: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit patterns, round to nearest even."""
    u = np.asarray(a, np.float32).reshape(-1).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def write_safetensors(path, tensors: dict, dtype: str = "F32", metadata: dict | None = None) -> None:
    """tensors: {name: array}; dtype F32 / F16 / BF16 for every tensor."""
    header, chunks, off = {}, [], 0
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    for name, a in tensors.items():
        a = np.asarray(a, np.float32)          # (ascontiguousarray would turn a 0-d tensor into shape (1,))
        raw = {"F32": lambda: a.tobytes(), "F16": lambda: a.astype(np.float16).tobytes(),
               "BF16": lambda: _bf16_bits(a).tobytes()}[dtype]()
        header[name] = {"dtype": dtype, "shape": list(a.shape), "data_offsets": [off, off + len(raw)]}
        chunks.append(raw)
        off += len(raw)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for ch in chunks:
            f.write(ch)


def build_blob(param_table, source: SafeTensors, name_map: dict | None = None, strict: bool = True):
    """param_table: [(name, offset, numel)] of a model handle (`model.param_table()`).  Returns (blob, report) with
    report = {"missing": [parameters without a tensor], "unused": [file tensors nobody asked for]}.  A tensor whose
    element count differs from the parameter's is an error; `strict` also makes missing parameters one."""
    total = max(o + n for _, o, n in param_table)
    blob = np.zeros(total, np.float32)
    used, missing = set(), []
    for name, off, numel in param_table:
        src = (name_map or {}).get(name, name)
        parts = list(src) if isinstance(src, (list, tuple)) else [src]   # a list: tensors concatenated along axis 0
        if any(p not in source for p in parts):
            missing.append(name)
            continue
        t = np.concatenate([source.read(p).reshape(-1) for p in parts])
        if t.size != numel:
            raise TsdError(1, f"checkpoint tensor {src} has {t.size} elements, parameter {name} needs {numel}")
        blob[off:off + numel] = t
        used.update(parts)
    if strict and missing:
        raise TsdError(1, f"checkpoint lacks {len(missing)} parameters, first: {missing[0]}")
    return blob, {"missing": missing, "unused": [n for n in source.names() if n not in used]}


def export_model(model, path, dtype: str = "F32") -> None:
    """Every parameter of a model handle (Diffusion / Decoder / Encoder / Clip) under its struct-order name, in the
    reference layouts (conv OIHW, linear [out][in])."""
    table = model.param_table()
    write_safetensors(path, {name: model.get_param(i) for i, (name, _, _) in enumerate(table)}, dtype,
                      {"format": "tsd_b200", "params": len(table)})


def import_model(model, path, name_map: dict | None = None, strict: bool = True):
    st = SafeTensors(path)
    try:
        blob, report = build_blob(model.param_table(), st, name_map, strict)
    finally:
        st.close()
    model.load_weights(blob)
    return report


# reference layer number (diffusion.mojo:175-201) -> diffusers block path, for the 23-layer topology of the reference
_RES_PATH = {2: "down_blocks.0.resnets.0", 5: "down_blocks.1.resnets.0", 8: "down_blocks.2.resnets.0",
             10: "up_blocks.0.resnets.0", 12: "up_blocks.0.resnets.1", 15: "up_blocks.1.resnets.0",
             17: "up_blocks.1.resnets.1", 20: "up_blocks.2.resnets.0", 22: "up_blocks.2.resnets.1"}
_ATTN_PATH = {3: "down_blocks.0.attentions.0", 6: "down_blocks.1.attentions.0", 9: "down_blocks.2.attentions.0",
              11: "up_blocks.0.attentions.0", 13: "up_blocks.0.attentions.1", 16: "up_blocks.1.attentions.0",
              18: "up_blocks.1.attentions.1", 21: "up_blocks.2.attentions.0", 23: "up_blocks.2.attentions.1"}
_RES_FIELD = {"layer1": "norm1", "layer2": "conv1", "layer3": "time_emb_proj", "layer4": "norm2", "layer5": "conv2",
              "layer6": "conv_shortcut"}
_ATTN_FIELD = {"layer1": "norm", "layer2": "proj_in", "layer3": "transformer_blocks.0.norm1",
               "layer4.out_proj": "transformer_blocks.0.attn1.to_out.0", "layer5": "transformer_blocks.0.norm2",
               "layer6.q_proj": "transformer_blocks.0.attn2.to_q", "layer6.k_proj": "transformer_blocks.0.attn2.to_k",
               "layer6.v_proj": "transformer_blocks.0.attn2.to_v", "layer6.out_proj": "transformer_blocks.0.attn2.to_out.0",
               "layer7": "transformer_blocks.0.norm3", "layer8": "transformer_blocks.0.ff.net.0.proj",
               "layer9": "transformer_blocks.0.ff.net.2", "layer10": "proj_out"}


def tiny_sd_unet_name_map(param_names) -> dict:
    """{parameter name of a norm_affine Diffusion: diffusers tensor name, or a list of names to concatenate}."""
    out = {}
    for name in param_names:
        stem, kind = name.rsplit(".", 1)             # ".weight" / ".bias"
        if stem.startswith("time_embed.layer"):
            out[name] = f"time_embedding.linear_{stem[-1]}.{kind}"
        elif stem == "unet.layer1":
            out[name] = f"conv_in.{kind}"
        elif stem in ("unet.layer4", "unet.layer7"):
            out[name] = f"down_blocks.{0 if stem.endswith('4') else 1}.downsamplers.0.conv.{kind}"
        elif stem == "final.layer1":
            out[name] = f"conv_norm_out.{kind}"
        elif stem == "final.layer2":
            out[name] = f"conv_out.{kind}"
        elif stem.startswith("unet.layer"):
            layer, field = stem[len("unet.layer"):].split(".", 1)
            layer = int(layer)
            if layer in _RES_PATH:
                out[name] = f"{_RES_PATH[layer]}.{_RES_FIELD[field]}.{kind}"
            elif field == "layer4.in_proj":          # to_q | to_k | to_v stacked along the output axis (attention.mojo:29)
                base = f"{_ATTN_PATH[layer]}.transformer_blocks.0.attn1"
                out[name] = [f"{base}.to_q.{kind}", f"{base}.to_k.{kind}", f"{base}.to_v.{kind}"]
            else:
                out[name] = f"{_ATTN_PATH[layer]}.{_ATTN_FIELD[field]}.{kind}"
        else:
            raise KeyError(name)
    return out


# ---- VAE (diffusers AutoencoderKL names) and CLIP text encoder (transformers CLIPTextModel names) ----
_VAE_RES_FIELD = {"conv1": "conv1", "conv2": "conv2", "res_conv_layer": "conv_shortcut", "groupnorm1": "norm1",
                  "groupnorm2": "norm2"}
# reference struct field (vae.mojo:163-188 decoder, :71-112 encoder) -> diffusers module path
_VAE_DECODER_PATH = {"l1": "post_quant_conv", "l2": "decoder.conv_in", "l3": "decoder.mid_block.resnets.0",
                     "l4": "decoder.mid_block.attentions.0", "l5": "decoder.mid_block.resnets.1",
                     "l6": "decoder.up_blocks.0.resnets.0", "l7": "decoder.up_blocks.0.resnets.1",
                     "l8": "decoder.up_blocks.0.resnets.2", "l10": "decoder.up_blocks.0.upsamplers.0.conv",
                     "l11": "decoder.up_blocks.1.resnets.0", "l12": "decoder.up_blocks.1.resnets.1",
                     "l13": "decoder.up_blocks.1.resnets.2", "l15": "decoder.up_blocks.1.upsamplers.0.conv",
                     "l16": "decoder.up_blocks.2.resnets.0", "l17": "decoder.up_blocks.2.resnets.1",
                     "l18": "decoder.up_blocks.2.resnets.2", "l20": "decoder.up_blocks.2.upsamplers.0.conv",
                     "l21": "decoder.up_blocks.3.resnets.0", "l22": "decoder.up_blocks.3.resnets.1",
                     "l23": "decoder.up_blocks.3.resnets.2", "l24": "decoder.conv_norm_out", "l26": "decoder.conv_out"}
_VAE_ENCODER_PATH = {"l1": "encoder.conv_in", "l2": "encoder.down_blocks.0.resnets.0", "l3": "encoder.down_blocks.0.resnets.1",
                     "l4": "encoder.down_blocks.0.downsamplers.0.conv", "l5": "encoder.down_blocks.1.resnets.0",
                     "l6": "encoder.down_blocks.1.resnets.1", "l7": "encoder.down_blocks.1.downsamplers.0.conv",
                     "l8": "encoder.down_blocks.2.resnets.0", "l9": "encoder.down_blocks.2.resnets.1",
                     "l10": "encoder.down_blocks.2.downsamplers.0.conv", "l11": "encoder.down_blocks.3.resnets.0",
                     "l12": "encoder.down_blocks.3.resnets.1", "l13": "encoder.mid_block.resnets.0",
                     "l14": "encoder.mid_block.attentions.0", "l15": "encoder.mid_block.resnets.1",
                     "l16": "encoder.conv_norm_out", "l18": "encoder.conv_out", "l19": "quant_conv"}


def _vae_name_map(param_names, paths) -> dict:
    out = {}
    for name in param_names:
        stem, kind = name.rsplit(".", 1)
        layer, _, field = stem.partition(".")
        base = paths[layer]
        if not field:
            out[name] = f"{base}.{kind}"
        elif field == "attention.in_proj":      # to_q | to_k | to_v stacked along the output axis (attention.mojo:29)
            out[name] = [f"{base}.to_q.{kind}", f"{base}.to_k.{kind}", f"{base}.to_v.{kind}"]
        elif field == "attention.out_proj":
            out[name] = f"{base}.to_out.0.{kind}"
        elif field == "groupnorm":
            out[name] = f"{base}.group_norm.{kind}"
        else:
            out[name] = f"{base}.{_VAE_RES_FIELD[field]}.{kind}"
    return out


def tiny_sd_vae_decoder_name_map(param_names) -> dict:
    """{parameter name of a TSD_MODEL_NORM_AFFINE Decoder: AutoencoderKL tensor name (or names to concatenate)}."""
    return _vae_name_map(param_names, _VAE_DECODER_PATH)


def tiny_sd_vae_encoder_name_map(param_names) -> dict:
    """{parameter name of a TSD_MODEL_NORM_AFFINE Encoder: AutoencoderKL tensor name (or names to concatenate)}."""
    return _vae_name_map(param_names, _VAE_ENCODER_PATH)


def tiny_sd_clip_name_map(param_names) -> dict:
    """{parameter name of a TSD_MODEL_NORM_AFFINE Clip: CLIPTextModel tensor name (or names to concatenate)}.
    The token table and the position embedding are stored flat on our side: build_blob only needs equal sizes."""
    field = {"layer2.out_proj": "self_attn.out_proj", "layer4": "mlp.fc1", "layer5": "mlp.fc2", "layer1": "layer_norm1",
             "layer3": "layer_norm2"}
    out = {}
    for name in param_names:
        if name == "embedding.token_embedding.weight":
            out[name] = "text_model.embeddings.token_embedding.weight"
        elif name == "embedding.position_embedding":
            out[name] = "text_model.embeddings.position_embedding.weight"
        elif name.startswith("layernorm."):
            out[name] = "text_model.final_layer_norm." + name.rsplit(".", 1)[1]
        else:
            stem, kind = name.rsplit(".", 1)
            layer, f = stem.split(".", 1)
            base = f"text_model.encoder.layers.{int(layer[len('player'):]) - 1}"
            if f == "layer2.in_proj":
                out[name] = [f"{base}.self_attn.{p}_proj.{kind}" for p in "qkv"]
            else:
                out[name] = f"{base}.{field[f]}.{kind}"
    return out
