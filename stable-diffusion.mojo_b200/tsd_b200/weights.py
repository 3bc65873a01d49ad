"""Checkpoint I/O (SURVEY section 8 row f2 - the reference's own TODO, README.md:44,55: its weights are random).

* `SafeTensors`: reader over the C ABI (csrc/host_io.cu): names, shapes, tensors as fp32 (F32 / F16 / BF16 / F64).
* `write_safetensors`: pure-Python writer (tests, export).
* `build_blob`: the flat fp32 blob `load_weights` expects, assembled in the model's parameter order from a
  checkpoint and a name map {parameter name of this library: tensor name in the file}.
* `export_model` / `import_model`: a model's parameters to / from a safetensors file under this library's names.

What is NOT here: a verified name map for `segmind/tiny-sd`.  No checkpoint exists offline to check one against, and
the reference's structs own no GroupNorm / LayerNorm affine tensors (helpers/utils.mojo:1825-1872, 2052-2061), so a
real checkpoint's norm weights have nowhere to go without extending the model; `missing` / `unused` in the report of
`build_blob` make both visible instead of hiding them."""
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
        if src not in source:
            missing.append(name)
            continue
        t = source.read(src)
        if t.size != numel:
            raise TsdError(1, f"checkpoint tensor {src} has {t.size} elements, parameter {name} needs {numel}")
        blob[off:off + numel] = t.reshape(-1)
        used.add(src)
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
