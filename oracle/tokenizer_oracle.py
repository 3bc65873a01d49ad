"""CPU restatement of the reference's prompt tokenizer and of this repo's PNG container - TEST
INFRASTRUCTURE ONLY (imported by tests/ and tools/make_golden.py; the product path is csrc/host_io.cu).

PARITY UNPINNED: the reference holds no golden vectors for this path and cannot run here (Mojo); the
real asset (tokenizer_clip.bin) needs a network download (tokenizer_creation.py:6-12).  The restatement
follows the reference line by line, including its own quicksort + binary search and str_concat exactly as
written, so that the product (hash-free sorted lookup, std::sort) is checked against an independent
formulation of the same algorithm.

  FileBuf / read_val_*         helpers/utils.mojo:63-141
  string_compare               :143-160
  partition / quicksort        :162-195
  wrap                         :197-206
  str_concat                   :214-224  (as written: every output position receives the first byte)
  Tokenizer                    :228-292
  bpe_encode                   :294-327
  tokenizer_creation.py        :20-48    (the .bin writer, double loop over vocab x merges)
"""
import struct
import sys
import zlib

START_ID, END_ID = "<|startoftext|>", "<|endoftext|>"


def string_compare(a: bytes, b: bytes) -> int:          # utils.mojo:143-160 (C strings: no NUL inside)
    i = 0
    while i < len(a) and i < len(b):
        if a[i] < b[i]:
            return -1
        if a[i] > b[i]:
            return 1
        i += 1
    if i < len(a) and i == len(b):
        return 1
    if i == len(a) and i < len(b):
        return -1
    return 0


def _partition(arr, idx, low, high):                    # utils.mojo:162-186
    pivot = arr[high]
    ii = low - 1
    for jj in range(low, high):
        if string_compare(pivot, arr[jj]) == 1:
            ii += 1
            arr[ii], arr[jj] = arr[jj], arr[ii]
            idx[ii], idx[jj] = idx[jj], idx[ii]
    arr[ii + 1], arr[high] = arr[high], arr[ii + 1]
    idx[ii + 1], idx[high] = idx[high], idx[ii + 1]
    return ii + 1


def quicksort(arr, idx, low, high):                     # utils.mojo:188-195 (explicit stack: same swaps)
    stack = [(low, high)]
    while stack:
        lo, hi = stack.pop()
        if lo < hi:
            pi = _partition(arr, idx, lo, hi)
            stack.append((pi + 1, hi))
            stack.append((lo, pi - 1))


def wrap(token: bytes) -> bytes:                        # utils.mojo:197-206
    if token == b"\\n":
        return b"<0x0A>"
    if token == b"\\t":
        return b"<0x09>"
    if token == b"'":
        return b"<0x27>"
    if token == b'"':
        return b"<0x22>"
    return token


def str_concat(s1: bytes, s2: bytes, as_written: bool = True) -> bytes:   # utils.mojo:214-224
    if not as_written:
        return s1 + s2
    # memcpy[count=1](string.offset(i), s1): the source pointer never advances
    return (s1[:1] * len(s1)) + (s2[:1] * len(s2))


class Tokenizer:                                        # utils.mojo:228-292
    def __init__(self, vocab_size: int, buf: bytes):
        self.vocab_size = vocab_size
        off = 0
        (self.max_token_length,) = struct.unpack_from("<i", buf, off)
        off += 4
        self.vocab, self.vocab_scores = [], []
        for _ in range(vocab_size):
            score, slen = struct.unpack_from("<fi", buf, off)
            off += 8
            tok = bytes(buf[off:off + slen])
            if len(tok) != slen:
                raise ValueError("tokenizer file is truncated")
            off += slen
            nul = tok.find(b"\0")                       # read_val_str yields a C string
            self.vocab.append(tok if nul < 0 else tok[:nul])
            self.vocab_scores.append(score)
        self.sorted_vocab, self.sorted_indices = None, None

    def sort(self):                                     # :264-274
        self.sorted_vocab = list(self.vocab)
        self.sorted_indices = list(range(self.vocab_size))
        quicksort(self.sorted_vocab, self.sorted_indices, 0, self.vocab_size - 1)

    def find(self, token: bytes) -> int:                # :276-292
        token = wrap(token)
        if self.sorted_indices is None:
            self.sort()
        left, right = 0, self.vocab_size - 1
        while left <= right:
            mid = left + (right - left) // 2
            c = string_compare(self.sorted_vocab[mid], token)
            if c == 0:
                return self.sorted_indices[mid]
            if c < 0:
                left = mid + 1
            else:
                right = mid - 1
        return -1


def bpe_encode(text: bytes, tok: Tokenizer, as_written: bool = True):      # utils.mojo:294-327
    """Returns (ids, complete); complete is False when a byte had no token (the reference prints
    "Not a good prompt token" and returns the ids collected so far)."""
    tokens = []
    for pos in range(len(text)):
        tid = tok.find(text[pos:pos + 1])
        if tid == -1:
            return tokens, False
        tokens.append(tid)
    while True:
        best_score, best_id, best_idx = -1e10, -1, -1
        for i in range(len(tokens) - 1):
            s = str_concat(tok.vocab[tokens[i]], tok.vocab[tokens[i + 1]], as_written)
            tid = tok.find(s)
            if tid != -1 and tok.vocab_scores[tid] > best_score:
                best_score, best_id, best_idx = tok.vocab_scores[tid], tid, i
        if best_idx == -1:
            break
        tokens[best_idx] = best_id
        tokens = tokens[:best_idx + 1] + tokens[best_idx + 2:]
    return tokens, True


def preprocess_prompt(prompt: str) -> bytes:            # pipeline.mojo:39-40
    return prompt.replace(" ", "</w>").encode("utf-8")


def tokenizer_bin(vocab_keys, merges) -> bytes:         # tokenizer_creation.py:20-48
    tokens, scores = [], []
    for key in vocab_keys:
        k = key
        if k == START_ID:
            k = "\n<s>\n"
        elif k == END_ID:
            k = "\n</s>\n"
        tokens.append(k.encode("utf-8"))
        score = 0.0
        for merge in merges:
            score += merge.count(key)
        scores.append(score)
    out = [struct.pack("I", max(len(k) for k in tokens))]
    for b, s in zip(tokens, scores):
        out.append(struct.pack("fI", s, len(b)))
        out.append(b)
    return b"".join(out)


# ---- PNG container (this repo's own output format; checked against zlib / the PNG specification) -------
def png_decode(png: bytes):
    """Minimal reader for the files csrc/host_io.cu writes: returns (h, w, c, rows of bytes)."""
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    off, idat, hdr = 8, b"", None
    while off < len(png):
        (n,) = struct.unpack_from(">I", png, off)
        typ, data = png[off + 4:off + 8], png[off + 8:off + 8 + n]
        (crc,) = struct.unpack_from(">I", png, off + 8 + n)
        assert crc == zlib.crc32(typ + data) & 0xFFFFFFFF, "chunk CRC"
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", data)
        elif typ == b"IDAT":
            idat += data
        off += 12 + n
    w, h, depth, ctype, comp, flt, lace = hdr
    assert (depth, comp, flt, lace) == (8, 0, 0, 0)
    c = {0: 1, 2: 3, 6: 4}[ctype]
    raw = zlib.decompress(idat)
    assert len(raw) == h * (w * c + 1)
    rows = []
    for y in range(h):
        r = raw[y * (w * c + 1):(y + 1) * (w * c + 1)]
        assert r[0] == 0
        rows.append(r[1:])
    return h, w, c, rows


def synthetic_vocab(n_merges: int = 400, seed: int = 5):
    """A small BPE vocabulary in the shape tokenizer_creation.py reads (vocab keys in id order + merges):
    printable single characters, then merges learnt greedily from a fixed word list."""
    import random
    rnd = random.Random(seed)
    words = ("a cat flying a spaceship the quick brown fox jumps over lazy dog astronaut riding horse on mars "
             "photo of an old castle at sunset highly detailed oil painting robot candy cyber city night rain "
             "mountain lake forest winter summer flowers portrait woman man child smiling blue red green").split()
    corpus = [list(w) + ["</w>"] for w in words for _ in range(rnd.randint(1, 4))]
    keys = [chr(c) for c in range(33, 127)]
    keys += [k + "</w>" for k in keys]
    merges = []
    for _ in range(n_merges):
        pairs = {}
        for w in corpus:
            for a, b in zip(w, w[1:]):
                pairs[(a, b)] = pairs.get((a, b), 0) + 1
        if not pairs:
            break
        (a, b), _cnt = max(sorted(pairs.items()), key=lambda kv: kv[1])
        merges.append(a + " " + b)
        if a + b not in keys:
            keys.append(a + b)
        for w in corpus:
            i = 0
            while i + 1 < len(w):
                if w[i] == a and w[i + 1] == b:
                    w[i:i + 2] = [a + b]
                else:
                    i += 1
    keys += [START_ID, END_ID]
    return keys, merges


if __name__ == "__main__":
    keys, merges = synthetic_vocab()
    blob = tokenizer_bin(keys, merges)
    t = Tokenizer(len(keys), blob)
    for p in sys.argv[1:] or ["a cat flying a spaceship"]:
        for mode in (True, False):
            ids, ok = bpe_encode(preprocess_prompt(p), t, mode)
            print(mode, ok, ids, [t.vocab[i] for i in ids])
