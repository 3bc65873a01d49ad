"""SURVEY section 8 row f4 - the data formats either side of the device path, through the C ABI
(csrc/host_io.cu; no GPU involved): tokenizer_clip.bin reader + bpe_encode against the line-by-line
restatement of helpers/utils.mojo:143-327 (oracle/tokenizer_oracle.py), bit-exact, and the PNG writer
against zlib / PIL.  PARITY UNPINNED by the reference (no tests, no asset offline): the committed
tests/golden/tokenizer_small.* pin the restatement."""
import io
import json
import os
import random
import struct

import numpy as np
import pytest

import tokenizer_oracle as T
from conftest import GOLDEN
from tsd_b200 import tokenizer as P
from tsd_b200._lib import TsdError
from tsd_b200.image import encode_png, save_png


@pytest.fixture(scope="module")
def vocab():
    keys, merges = T.synthetic_vocab()
    blob = T.tokenizer_bin(keys, merges)
    return keys, merges, blob


@pytest.fixture(scope="module")
def toks(vocab):
    keys, _, blob = vocab
    return P.Tokenizer(blob, len(keys)), T.Tokenizer(len(keys), blob)


def test_golden_bin_is_reproduced_and_loads_from_disk(vocab):
    keys, merges, blob = vocab
    path = os.path.join(GOLDEN, "tokenizer_small.bin")
    assert open(path, "rb").read() == blob
    meta = json.load(open(os.path.join(GOLDEN, "tokenizer_small.json")))
    assert meta["vocab_size"] == len(keys) and meta["n_merges"] == len(merges)
    tok = P.Tokenizer(path, len(keys))                       # read_file + Tokenizer(n, buf), pipeline.mojo:32-37
    assert tok.vocab_size == len(keys) and tok.max_token_length == struct.unpack_from("<I", blob)[0]
    for case in meta["cases"]:
        text = P.preprocess_prompt(case["prompt"])
        assert tok.encode(text, concat_as_written=True) == case["as_written"], case["prompt"]
        assert tok.encode(text, concat_as_written=False) == case["intended"], case["prompt"]
        if not case["complete"]:
            with pytest.raises(TsdError):
                tok.encode(text, strict=True)
    # the two readings of str_concat differ once a merge involves a multi-byte token
    c0 = meta["cases"][0]
    assert c0["as_written"] != c0["intended"] and len(c0["intended"]) < len(c0["as_written"])


def test_vocabulary_table_and_find(toks, vocab):
    prod, ref = toks
    keys = vocab[0]
    for i in range(len(keys)):
        b, s = prod.token(i)
        assert b == ref.vocab[i] and s == ref.vocab_scores[i]
        # wrap() sends the two quote characters to their <0xXX> spelling, absent here (utils.mojo:197-206)
        assert prod.find(b) == ref.find(b) == (-1 if b in (b"'", b'"') else i)
    assert prod.token(len(keys) - 2)[0] == b"\n<s>\n" and prod.token(len(keys) - 1)[0] == b"\n</s>\n"
    for s in (b"", b"zzzzzz", b"\xff", b"a\0b", b"'", b'"', b"\\n", b"\\t", b"<0x27>", b"th", b"</w>"):
        assert prod.find(s) == ref.find(s.split(b"\0")[0]), s
    with pytest.raises(IndexError):
        prod.token(len(keys))


def test_bpe_encode_matches_restatement_on_random_prompts(toks):
    prod, ref = toks
    rnd = random.Random(3)
    words = "a cat flying spaceship the of an astronaut riding horse mars castle sunset oil painting xyzzy qq".split()
    alphabet = "abcdefghijklmnopqrstuvwxyz'\"<>/w .,!\té"
    for trial in range(200):
        if trial % 2:
            prompt = " ".join(rnd.choice(words) for _ in range(rnd.randint(0, 12)))
        else:
            prompt = "".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 40)))
        text = T.preprocess_prompt(prompt)
        for as_written in (True, False):
            want, complete = T.bpe_encode(text, ref, as_written)
            assert prod.encode(text, concat_as_written=as_written) == want, (prompt, as_written)
            if not complete:
                with pytest.raises(TsdError):
                    prod.encode(text, concat_as_written=as_written, strict=True)
    assert list(P.prompt_tokens("a cat", prod)) == T.bpe_encode(T.preprocess_prompt("a cat"), ref)[0]


def test_merge_order_follows_scores_not_position():
    """Greedy rule (utils.mojo:303-326): highest score first, the leftmost pair on ties."""
    keys = ["a", "b", "c", "ab", "bc", "abc"]
    def blob(scores):
        out = [struct.pack("I", 3)]
        for k, s in zip(keys, scores):
            out += [struct.pack("fI", s, len(k)), k.encode()]
        return b"".join(out)
    for scores, want in (([0, 0, 0, 1, 2, 5], [0, 4]),        # "bc" outranks "ab"; "a"+"bc" = "abc" only if intended
                         ([0, 0, 0, 2, 1, 5], [3, 2]),
                         ([0, 0, 0, 1, 1, 5], [3, 2])):       # tie: leftmost
        b = blob(scores)
        prod, ref = P.Tokenizer(b, 6), T.Tokenizer(6, b)
        assert prod.encode("abc", concat_as_written=True) == T.bpe_encode(b"abc", ref, True)[0] == want
        assert prod.encode("abc", concat_as_written=False) == T.bpe_encode(b"abc", ref, False)[0] == [5]


def test_malformed_vocabulary_files(vocab, tmp_path):
    keys, _, blob = vocab
    assert P.Tokenizer(blob, 10).vocab_size == 10            # a smaller vocab_size reads a prefix
    for bad in (blob[:-3], blob[:100], blob[:3], b""):
        with pytest.raises(TsdError):
            P.Tokenizer(bad, len(keys))
    with pytest.raises(TsdError):
        P.Tokenizer(blob, len(keys) + 1)                      # more tokens requested than the file holds
    with pytest.raises(TsdError):
        P.Tokenizer(str(tmp_path / "missing.bin"), 5)
    with pytest.raises(TsdError):
        P.Tokenizer(blob, 0)


def test_bin_writer_matches_tokenizer_creation(vocab):
    """tsd_b200.tokenizer.merge_scores + tokenizer_bin (one pass over substrings) against the restated double
    loop of tokenizer_creation.py:26-48, for "a b" string merges and for [a, b] list merges."""
    keys, merges, blob = vocab
    assert P.tokenizer_bin(keys, P.merge_scores(keys, merges)) == blob
    pairs = [m.split(" ") for m in merges]
    assert P.tokenizer_bin(keys, P.merge_scores(keys, pairs)) == T.tokenizer_bin(keys, pairs)
    assert P.merge_scores(["aa", "a", ""], ["aaa a", "b aa"]) == [2.0, 6.0, 11.0]   # non-overlapping counts


def test_bin_from_tokenizer_json(tmp_path, vocab):
    keys, merges, blob = vocab
    j = tmp_path / "tokenizer.json"
    j.write_text(json.dumps({"model": {"vocab": {k: i for i, k in enumerate(keys)}, "merges": merges}}))
    out = tmp_path / "tokenizer_clip.bin"
    assert P.tokenizer_bin_from_json(str(j), str(out)) == len(keys)
    assert out.read_bytes() == blob


# ---- PNG -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("c,h,w", [(3, 1, 1), (3, 37, 53), (1, 8, 300), (4, 5, 7), (3, 256, 256)])
def test_png_round_trip(c, h, w, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(c * 100 + h)
    img = rng.uniform(-20.0, 280.0, (c, h, w)).astype(np.float32)
    img.reshape(-1)[:3] = [0.5, 254.5, 1.4999]               # round half up: 1, 255, 1
    want = np.clip(np.floor(img.astype(np.float64) + 0.5), 0, 255).astype(np.uint8)
    png = encode_png(img)
    hh, ww, cc, rows = T.png_decode(png)                      # chunk CRCs, zlib stream (Adler-32), filter bytes
    assert (hh, ww, cc) == (h, w, c)
    assert np.array_equal(np.frombuffer(b"".join(rows), np.uint8).reshape(h, w, c), want.transpose(1, 2, 0))
    pil = np.asarray(Image.open(io.BytesIO(png)))
    assert np.array_equal(pil.reshape(h, w, c), want.transpose(1, 2, 0))
    path = tmp_path / "out.png"
    save_png(path, img)
    assert path.read_bytes() == png


def test_png_validation(tmp_path):
    with pytest.raises(TsdError):
        encode_png(np.zeros((2, 4, 4), np.float32))           # 2 channels: not a PNG colour type
    with pytest.raises(TsdError):
        encode_png(np.zeros((4, 4), np.float32))
    with pytest.raises(TsdError):
        save_png(tmp_path / "no_such_dir" / "x.png", np.zeros((3, 2, 2), np.float32))
    nan = np.full((1, 1, 2), np.nan, np.float32)
    assert T.png_decode(encode_png(nan))[3] == [b"\0\0"]


def test_bpe_random_vocabularies_match_restatement():
    """Random vocabularies (binary token bytes >= 0x80, score ties, tokens that are concatenations of others) and random
    byte strings: ids from the C++ encoder == ids from the line-by-line restatement, both str_concat readings."""
    rnd = random.Random(11)
    for trial in range(25):
        alphabet = bytes(rnd.sample(range(1, 256), rnd.randint(3, 12)))           # no NUL: tokens are C strings
        singles = [bytes([b]) for b in alphabet if bytes([b]) not in (b"'", b'"')]
        toks = list(dict.fromkeys(singles))
        while len(toks) < len(singles) + rnd.randint(5, 60):
            a, b = rnd.choice(toks), rnd.choice(toks)
            t = (a + b)[:rnd.randint(2, 7)]
            if t not in toks and t not in (b"\\n", b"\\t"):
                toks.append(t)
        rnd.shuffle(toks)
        scores = [float(rnd.randint(0, 6)) for _ in toks]                          # many ties
        blob = struct.pack("I", max(len(t) for t in toks))
        for t, s in zip(toks, scores):
            blob += struct.pack("fI", s, len(t)) + t
        prod, ref = P.Tokenizer(blob, len(toks)), T.Tokenizer(len(toks), blob)
        for t in toks:
            assert prod.find(t) == ref.find(t)
        for _ in range(20):
            text = bytes(rnd.choice(alphabet + b"\x00'") for _ in range(rnd.randint(0, 30)))
            if b"\x00" in text:
                text = text.replace(b"\x00", b"")                                   # ctypes passes C strings
            for as_written in (True, False):
                want, _ = T.bpe_encode(text, ref, as_written)
                assert prod.encode(text, concat_as_written=as_written) == want, (trial, text, as_written)
        prod.close()
