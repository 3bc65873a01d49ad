"""Host side of the prompt path over the C ABI: Tokenizer + bpe_encode
(reference helpers/utils.mojo:228-327), the tokenizer_clip.bin writer of tokenizer_creation.py:20-48 and
the prompt preprocessing of pipeline.mojo:39-40.  The lookups and merges run in libtsd_b200.so
(csrc/host_io.cu); nothing here needs a GPU."""
from __future__ import annotations

import ctypes as C
import json
import struct

import numpy as np

from . import _lib
from ._lib import TsdError

START_ID, END_ID = "<|startoftext|>", "<|endoftext|>"        # tokenizer_creation.py:17-18


class Tokenizer:
    """Tokenizer(vocab_size, buf), utils.mojo:228-292.  `source` is a path or the bytes of a .bin file."""

    def __init__(self, source, vocab_size: int = 49408):            # pipeline.mojo:36-37
        self.L = _lib.lib()
        h = C.c_void_p()
        if isinstance(source, (bytes, bytearray, memoryview)):
            buf = bytes(source)
            rc = self.L.tsd_tokenizer_from_memory(buf, len(buf), int(vocab_size), C.byref(h))
        else:
            rc = self.L.tsd_tokenizer_load(str(source).encode(), int(vocab_size), C.byref(h))
        if rc:
            raise TsdError(rc, "tokenizer: missing, truncated or malformed vocabulary file")
        self.h = h

    @property
    def vocab_size(self) -> int:
        return self.L.tsd_tokenizer_vocab_size(self.h)

    @property
    def max_token_length(self) -> int:
        return self.L.tsd_tokenizer_max_token_length(self.h)

    def token(self, i: int):
        """(bytes, score) of token id i."""
        n, s = C.c_int32(), C.c_float()
        p = self.L.tsd_tokenizer_token(self.h, int(i), C.byref(n), C.byref(s))
        if not p:
            raise IndexError(i)
        return C.string_at(p, n.value), s.value

    def find(self, s) -> int:                                        # utils.mojo:276-292
        b = s.encode() if isinstance(s, str) else bytes(s)
        return self.L.tsd_tokenizer_find(self.h, b, len(b))

    def encode(self, text, concat_as_written: bool = True, strict: bool = False) -> list[int]:
        """bpe_encode(text, tok), utils.mojo:294-327.  A byte without a token ends the encoding with the ids
        collected so far, as the reference does (strict=True raises instead)."""
        b = text.encode() if isinstance(text, str) else bytes(text)
        ids = (C.c_int32 * max(1, len(b)))()
        n = C.c_int32()
        rc = self.L.tsd_tokenizer_encode(self.h, b, len(b), 0 if concat_as_written else 1, ids, len(b), C.byref(n))
        if rc and (strict or rc != 1):
            raise TsdError(rc, "Not a good prompt token")
        return list(ids[:n.value])

    def close(self):
        if getattr(self, "h", None):
            self.L.tsd_tokenizer_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bpe_encode(text, tok: Tokenizer, concat_as_written: bool = True) -> list[int]:
    return tok.encode(text, concat_as_written)


def preprocess_prompt(prompt: str) -> str:
    """pipeline.mojo:39-40"""
    return prompt.replace(" ", "</w>")


def prompt_tokens(prompt: str, tok: Tokenizer, concat_as_written: bool = True) -> np.ndarray:
    """pipeline.mojo:39-53: the token ids clip.forward receives (zero padding to 77 happens in CLIP.forward)."""
    return np.asarray(tok.encode(preprocess_prompt(prompt), concat_as_written), np.int32)


def merge_scores(vocab_keys, merges) -> list[float]:
    """tokenizer_creation.py:36-39: score(key) = sum over merges of merge.count(key).  `merges` entries are
    either "a b" strings (str.count: non-overlapping substring occurrences) or [a, b] lists (list.count: equal
    elements), whichever the tokenizer.json at hand holds.  Same result as the reference's double loop, via
    one pass over the substrings of every merge."""
    counts: dict = {}
    empty = 0.0                                     # "ab".count("") == len + 1; ["a", "b"].count("") == 0
    for m in merges:
        if isinstance(m, str):
            empty += len(m) + 1
            for s in {m[i:j] for i in range(len(m)) for j in range(i + 1, len(m) + 1)}:
                counts[s] = counts.get(s, 0) + m.count(s)
        else:
            for s in set(m):
                counts[s] = counts.get(s, 0) + list(m).count(s)
    return [float(empty if k == "" and "" not in counts else counts.get(k, 0)) for k in vocab_keys]


def tokenizer_bin(vocab_keys, scores) -> bytes:
    """tokenizer_creation.py:26-48: start/end markers renamed, uint32 max length, (float32, uint32, bytes)*."""
    toks = []
    for k in vocab_keys:
        if k == START_ID:
            k = "\n<s>\n"
        elif k == END_ID:
            k = "\n</s>\n"
        toks.append(k.encode("utf-8"))
    out = [struct.pack("I", max(len(t) for t in toks))]
    for t, s in zip(toks, scores):
        out.append(struct.pack("fI", s, len(t)))
        out.append(t)
    return b"".join(out)


def tokenizer_bin_from_json(tokenizer_json: str, out_file: str = "tokenizer_clip.bin") -> int:
    """The __main__ of tokenizer_creation.py for an already downloaded clip_tokenizer/tokenizer.json."""
    with open(tokenizer_json, "r") as f:
        data = json.load(f)
    merges = data["model"]["merges"]
    keys = list(data["model"]["vocab"].keys())
    blob = tokenizer_bin(keys, merge_scores(keys, merges))
    with open(out_file, "wb") as f:
        f.write(blob)
    return len(keys)
