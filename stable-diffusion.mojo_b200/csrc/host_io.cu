// Host-side data formats either side of the device path: the tokenizer_clip.bin
// vocabulary reader + greedy score-based BPE (helpers/utils.mojo:228-327; file format written by
// tokenizer_creation.py:44-48) and a PNG writer for the (0..255) float image pipeline.generate returns
// (pipeline.mojo:127-128; the reference never stores it).  Plain C++, byte and integer work only - nothing
// here touches the GPU.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/tsd_b200.h"

struct tsd_tokenizer {
  int32_t max_token_length = 0;
  std::vector<std::string> vocab;  // C-string semantics of the reference: cut at the first NUL byte
  std::vector<float> scores;
  std::vector<int32_t> order;  // ids sorted by (token bytes, id): Tokenizer.sort, utils.mojo:264-274
};

namespace {

struct Reader {
  const uint8_t* p;
  int64_t size, off = 0;
  bool take(void* dst, int64_t n) {
    if (n < 0 || off + n > size) return false;
    std::memcpy(dst, p + off, (size_t)n);
    off += n;
    return true;
  }
};

// string_compare, utils.mojo:143-160: unsigned bytewise order, the shorter string first on a tie
inline int cmp_bytes(const std::string& a, const uint8_t* b, size_t nb) {
  const size_t n = std::min(a.size(), nb);
  const int c = n ? std::memcmp(a.data(), b, n) : 0;
  if (c) return c < 0 ? -1 : 1;
  return a.size() == nb ? 0 : (a.size() < nb ? -1 : 1);
}

// wrap, utils.mojo:197-206: the two-character strings backslash-n / backslash-t and the two quote
// characters are looked up under their <0xXX> spelling
inline void wrap(const uint8_t*& s, size_t& n) {
  static const char* nl = "<0x0A>";
  static const char* tab = "<0x09>";
  static const char* sq = "<0x27>";
  static const char* dq = "<0x22>";
  const char* r = nullptr;
  if (n == 2 && s[0] == '\\' && s[1] == 'n') r = nl;
  else if (n == 2 && s[0] == '\\' && s[1] == 't') r = tab;
  else if (n == 1 && s[0] == '\'') r = sq;
  else if (n == 1 && s[0] == '"') r = dq;
  if (r) {
    s = reinterpret_cast<const uint8_t*>(r);
    n = 6;
  }
}

// Tokenizer.find, utils.mojo:276-292.  Duplicate token strings (none in a vocabulary produced by
// tokenizer_creation.py: JSON object keys are unique) resolve to the lowest id.
int32_t find_token(const tsd_tokenizer* t, const uint8_t* s, size_t n) {
  const void* nul = n ? std::memchr(s, 0, n) : nullptr;  // the reference compares C strings
  if (nul) n = (size_t)(static_cast<const uint8_t*>(nul) - s);
  wrap(s, n);
  int64_t lo = 0, hi = (int64_t)t->order.size();
  while (lo < hi) {  // first entry >= s
    const int64_t mid = lo + (hi - lo) / 2;
    if (cmp_bytes(t->vocab[t->order[mid]], s, n) < 0) lo = mid + 1;
    else hi = mid;
  }
  if (lo < (int64_t)t->order.size() && cmp_bytes(t->vocab[t->order[lo]], s, n) == 0) return t->order[lo];
  return -1;
}

int32_t tokenizer_parse(const uint8_t* buf, int64_t size, int32_t vocab_size, tsd_tokenizer** out) {
  if (!buf || !out || vocab_size <= 0 || size < 4) return TSD_ERR_INVALID;
  *out = nullptr;
  tsd_tokenizer* t = new (std::nothrow) tsd_tokenizer();
  if (!t) return TSD_ERR_OOM;
  Reader r{buf, size};
  uint32_t maxlen = 0;
  bool ok = r.take(&maxlen, 4);  // Tokenizer.__init__, utils.mojo:236-249
  t->max_token_length = (int32_t)maxlen;
  t->vocab.resize(vocab_size);
  t->scores.resize(vocab_size);
  for (int32_t i = 0; ok && i < vocab_size; ++i) {
    float score = 0.f;
    uint32_t len = 0;
    ok = r.take(&score, 4) && r.take(&len, 4) && (int64_t)len <= size - r.off;
    if (!ok) break;
    const char* s = reinterpret_cast<const char*>(buf + r.off);
    t->vocab[i].assign(s, strnlen(s, len));
    t->scores[i] = score;
    r.off += len;
  }
  if (!ok) {  // the reference prints "Error reading ..." and carries on with zeros; fail instead
    delete t;
    return TSD_ERR_INVALID;
  }
  t->order.resize(vocab_size);
  for (int32_t i = 0; i < vocab_size; ++i) t->order[i] = i;
  std::sort(t->order.begin(), t->order.end(), [&](int32_t a, int32_t b) {
    const std::string& sa = t->vocab[a];
    const int c = cmp_bytes(sa, reinterpret_cast<const uint8_t*>(t->vocab[b].data()), t->vocab[b].size());
    return c ? c < 0 : a < b;
  });
  *out = t;
  return TSD_OK;
}

// ---- PNG ------------------------------------------------------------------------------------
uint32_t crc_table[256];
bool crc_ready = false;
void crc_init() {
  for (uint32_t n = 0; n < 256; ++n) {
    uint32_t c = n;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    crc_table[n] = c;
  }
  crc_ready = true;
}
uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
  for (size_t i = 0; i < n; ++i) crc = crc_table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
  return crc;
}
void put_be32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24));
  v.push_back((uint8_t)(x >> 16));
  v.push_back((uint8_t)(x >> 8));
  v.push_back((uint8_t)x);
}
void put_chunk(std::vector<uint8_t>& v, const char type[4], const uint8_t* data, size_t n) {
  put_be32(v, (uint32_t)n);
  const size_t start = v.size();
  v.insert(v.end(), type, type + 4);
  if (n) v.insert(v.end(), data, data + n);
  put_be32(v, crc32_update(0xFFFFFFFFu, v.data() + start, n + 4) ^ 0xFFFFFFFFu);
}

// (C,H,W) planar float -> 8-bit PNG, filter 0 on every row, zlib stream of stored (uncompressed) blocks
int32_t png_build(const float* img, int32_t c, int32_t h, int32_t w, std::vector<uint8_t>& png) {
  if (!img || h <= 0 || w <= 0 || (c != 1 && c != 3 && c != 4)) return TSD_ERR_INVALID;
  if (!crc_ready) crc_init();
  const size_t row = (size_t)w * c + 1, raw_n = row * h;
  std::vector<uint8_t> raw(raw_n);
  const size_t plane = (size_t)h * w;
  for (int32_t y = 0; y < h; ++y) {
    uint8_t* r = raw.data() + row * y;
    r[0] = 0;
    for (int32_t x = 0; x < w; ++x)
      for (int32_t k = 0; k < c; ++k) {
        float v = img[k * plane + (size_t)y * w + x];
        v = v != v ? 0.0f : std::floor(v + 0.5f);  // round half up; NaN -> 0
        r[1 + (size_t)x * c + k] = (uint8_t)(v < 0.f ? 0.f : (v > 255.f ? 255.f : v));
      }
  }
  std::vector<uint8_t> z;
  z.reserve(raw_n + raw_n / 65535 * 5 + 16);
  z.push_back(0x78);
  z.push_back(0x01);
  uint32_t a = 1, b = 0;  // Adler-32
  size_t off = 0;
  do {
    const size_t n = std::min<size_t>(65535, raw_n - off);
    z.push_back(off + n == raw_n ? 1 : 0);
    z.push_back((uint8_t)(n & 0xFF));
    z.push_back((uint8_t)(n >> 8));
    z.push_back((uint8_t)(~n & 0xFF));
    z.push_back((uint8_t)((~n >> 8) & 0xFF));
    z.insert(z.end(), raw.data() + off, raw.data() + off + n);
    for (size_t i = 0; i < n; ++i) {
      a += raw[off + i];
      if (a >= 65521) a -= 65521;
      b += a;
      if (b >= 65521) b -= 65521;
    }
    off += n;
  } while (off < raw_n);
  put_be32(z, (b << 16) | a);
  png.clear();
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  png.insert(png.end(), sig, sig + 8);
  std::vector<uint8_t> ihdr;
  put_be32(ihdr, (uint32_t)w);
  put_be32(ihdr, (uint32_t)h);
  const uint8_t tail[5] = {8, (uint8_t)(c == 1 ? 0 : (c == 3 ? 2 : 6)), 0, 0, 0};
  ihdr.insert(ihdr.end(), tail, tail + 5);
  put_chunk(png, "IHDR", ihdr.data(), ihdr.size());
  put_chunk(png, "IDAT", z.data(), z.size());
  put_chunk(png, "IEND", nullptr, 0);
  return TSD_OK;
}

}  // namespace

extern "C" {

int32_t tsd_tokenizer_from_memory(const void* buf, int64_t size, int32_t vocab_size, tsd_tokenizer** out) {
  return tokenizer_parse(static_cast<const uint8_t*>(buf), size, vocab_size, out);
}

// read_file + Tokenizer(vocab_size, buf), pipeline.mojo:32-37
int32_t tsd_tokenizer_load(const char* path, int32_t vocab_size, tsd_tokenizer** out) {
  if (!path || !out) return TSD_ERR_INVALID;
  *out = nullptr;
  FILE* f = std::fopen(path, "rb");
  if (!f) return TSD_ERR_INVALID;
  std::vector<uint8_t> data;
  uint8_t chunk[1 << 16];
  size_t n;
  while ((n = std::fread(chunk, 1, sizeof(chunk), f)) > 0) data.insert(data.end(), chunk, chunk + n);
  std::fclose(f);
  return tokenizer_parse(data.data(), (int64_t)data.size(), vocab_size, out);
}

int32_t tsd_tokenizer_destroy(tsd_tokenizer* t) {
  if (!t) return TSD_ERR_INVALID;
  delete t;
  return TSD_OK;
}
int32_t tsd_tokenizer_vocab_size(const tsd_tokenizer* t) { return t ? (int32_t)t->vocab.size() : 0; }
int32_t tsd_tokenizer_max_token_length(const tsd_tokenizer* t) { return t ? t->max_token_length : 0; }
const uint8_t* tsd_tokenizer_token(const tsd_tokenizer* t, int32_t id, int32_t* len, float* score) {
  if (!t || id < 0 || id >= (int32_t)t->vocab.size()) return nullptr;
  if (len) *len = (int32_t)t->vocab[id].size();
  if (score) *score = t->scores[id];
  return reinterpret_cast<const uint8_t*>(t->vocab[id].data());
}
int32_t tsd_tokenizer_find(const tsd_tokenizer* t, const uint8_t* s, int32_t len) {
  if (!t || len < 0 || (!s && len)) return -1;
  return find_token(t, s, (size_t)len);
}

// bpe_encode, utils.mojo:294-327
int32_t tsd_tokenizer_encode(const tsd_tokenizer* t, const uint8_t* text, int32_t len, int32_t concat_mode,
                             int32_t* ids, int32_t cap, int32_t* n_out) {
  if (!t || len < 0 || (!text && len) || !n_out || cap < 0 || (!ids && cap)) return TSD_ERR_INVALID;
  *n_out = 0;
  std::vector<int32_t> tok;
  tok.reserve(len);
  int32_t status = TSD_OK;
  for (int32_t pos = 0; pos < len; ++pos) {  // one token per byte (:296-302)
    const int32_t id = find_token(t, text + pos, 1);
    if (id < 0) {  // "Not a good prompt token": the reference returns the ids collected so far, unmerged
      status = TSD_ERR_INVALID;
      break;
    }
    tok.push_back(id);
  }
  std::string cat;
  while (status == TSD_OK) {  // greedy merges: the pair whose concatenation has the highest score (:303-326)
    float best_score = -1e10f;
    int32_t best_id = -1, best_idx = -1;
    for (size_t i = 0; i + 1 < tok.size(); ++i) {
      const std::string &a = t->vocab[tok[i]], &b = t->vocab[tok[i + 1]];
      if (concat_mode == 0) {  // str_concat as written (:214-224): every position receives the FIRST byte
        cat.assign(a.size(), a.empty() ? '\0' : a[0]);
        cat.append(b.size(), b.empty() ? '\0' : b[0]);
      } else {
        cat.assign(a);
        cat.append(b);
      }
      const int32_t id = find_token(t, reinterpret_cast<const uint8_t*>(cat.data()), cat.size());
      if (id != -1 && t->scores[id] > best_score) {
        best_score = t->scores[id];
        best_id = id;
        best_idx = (int32_t)i;
      }
    }
    if (best_idx < 0) break;
    tok[best_idx] = best_id;
    tok.erase(tok.begin() + best_idx + 1);
  }
  *n_out = (int32_t)tok.size();
  if ((int32_t)tok.size() > cap) return TSD_ERR_OOM;
  if (!tok.empty()) std::memcpy(ids, tok.data(), tok.size() * sizeof(int32_t));
  return status;
}

int32_t tsd_png_encode(const float* img, int32_t c, int32_t h, int32_t w, uint8_t* out, int64_t cap,
                       int64_t* size) {
  if (!size) return TSD_ERR_INVALID;
  std::vector<uint8_t> png;
  const int32_t rc = png_build(img, c, h, w, png);
  if (rc) return rc;
  *size = (int64_t)png.size();
  if (!out || cap < (int64_t)png.size()) return out ? TSD_ERR_OOM : TSD_OK;  // out == NULL: size query
  std::memcpy(out, png.data(), png.size());
  return TSD_OK;
}

int32_t tsd_png_write(const char* path, const float* img, int32_t c, int32_t h, int32_t w) {
  if (!path) return TSD_ERR_INVALID;
  std::vector<uint8_t> png;
  const int32_t rc = png_build(img, c, h, w, png);
  if (rc) return rc;
  FILE* f = std::fopen(path, "wb");
  if (!f) return TSD_ERR_INVALID;
  const size_t n = std::fwrite(png.data(), 1, png.size(), f);
  const int cl = std::fclose(f);
  return (n == png.size() && cl == 0) ? TSD_OK : TSD_ERR_INVALID;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// safetensors reader (the reference's stated TODO README.md:44,55: "load the weights").
// File layout: uint64 little-endian header length N, N bytes of JSON {"name": {"dtype": "F32", "shape": [..],
// "data_offsets": [begin, end]}, ..., "__metadata__": {...}}, then the tensor bytes (offsets relative to the end
// of the header).  F32 / F16 / BF16 / F64 tensors are delivered as fp32.
// ---------------------------------------------------------------------------------------------------------
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

struct tsd_safetensors {
  struct Entry {
    std::string name, dtype;
    std::vector<int64_t> shape;
    int64_t begin = 0, end = 0, numel = 0;
  };
  std::vector<Entry> entries;
  const uint8_t* base = nullptr;  // start of the tensor bytes
  int64_t data_size = 0;
  void* map = nullptr;            // mmap of the whole file, or nullptr
  size_t map_size = 0;
  std::vector<uint8_t> owned;     // from_memory copy
};

namespace {

struct Json {  // minimal recursive-descent parser for the header's subset of JSON
  const char* p;
  const char* e;
  bool ok = true;
  void ws() {
    while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
  }
  bool eat(char c) {
    ws();
    if (p < e && *p == c) {
      ++p;
      return true;
    }
    return false;
  }
  bool str(std::string& out) {
    ws();
    if (p >= e || *p != '"') return ok = false;
    ++p;
    out.clear();
    while (p < e && *p != '"') {
      if (*p == '\\') {
        if (++p >= e) return ok = false;
        switch (*p) {
          case 'n': out.push_back('\n'); break;
          case 't': out.push_back('\t'); break;
          case 'r': out.push_back('\r'); break;
          case 'b': out.push_back('\b'); break;
          case 'f': out.push_back('\f'); break;
          case 'u': {
            if (e - p < 5) return ok = false;
            unsigned v = 0;
            for (int i = 1; i <= 4; ++i) {
              const char ch = p[i];
              v <<= 4;
              if (ch >= '0' && ch <= '9') v |= ch - '0';
              else if (ch >= 'a' && ch <= 'f') v |= ch - 'a' + 10;
              else if (ch >= 'A' && ch <= 'F') v |= ch - 'A' + 10;
              else return ok = false;
            }
            p += 4;
            if (v < 0x80) out.push_back((char)v);  // UTF-8 encode (BMP; surrogate pairs are kept as two code units)
            else if (v < 0x800) {
              out.push_back((char)(0xC0 | (v >> 6)));
              out.push_back((char)(0x80 | (v & 0x3F)));
            } else {
              out.push_back((char)(0xE0 | (v >> 12)));
              out.push_back((char)(0x80 | ((v >> 6) & 0x3F)));
              out.push_back((char)(0x80 | (v & 0x3F)));
            }
            break;
          }
          default: out.push_back(*p);  // \" \\ \/
        }
        ++p;
      } else {
        out.push_back(*p++);
      }
    }
    if (p >= e) return ok = false;
    ++p;
    return true;
  }
  bool integer(int64_t& v) {
    ws();
    const char* s = p;
    bool neg = false;
    if (p < e && *p == '-') {
      neg = true;
      ++p;
    }
    if (p >= e || *p < '0' || *p > '9') {
      p = s;
      return ok = false;
    }
    v = 0;
    while (p < e && *p >= '0' && *p <= '9') {
      if (v > (INT64_MAX - 9) / 10) return ok = false;
      v = v * 10 + (*p++ - '0');
    }
    if (neg) v = -v;
    return true;
  }
  bool int_array(std::vector<int64_t>& a) {
    a.clear();
    if (!eat('[')) return ok = false;
    if (eat(']')) return true;
    do {
      int64_t v;
      if (!integer(v)) return false;
      a.push_back(v);
    } while (eat(','));
    return eat(']') ? true : (ok = false);
  }
  int depth = 0;  // nesting of the value being skipped: the header is attacker-controlled, recursion is capped
  bool skip_value() {  // __metadata__ and unknown keys
    struct Level {
      int& d;
      explicit Level(int& dd) : d(dd) { ++d; }
      ~Level() { --d; }
    } level(depth);
    if (depth > 64) return ok = false;
    ws();
    if (p >= e) return ok = false;
    if (*p == '"') {
      std::string s;
      return str(s);
    }
    if (*p == '{' || *p == '[') {
      const char open = *p, close = open == '{' ? '}' : ']';
      ++p;
      if (eat(close)) return true;
      do {
        if (open == '{') {
          std::string k;
          if (!str(k) || !eat(':')) return ok = false;
        }
        if (!skip_value()) return false;
      } while (eat(','));
      return eat(close) ? true : (ok = false);
    }
    while (p < e && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n') ++p;  // number / true / false / null
    return true;
  }
};

int dtype_size(const std::string& d) {
  if (d == "F32") return 4;
  if (d == "F16" || d == "BF16") return 2;
  if (d == "F64") return 8;
  return 0;  // I8/I32/... are listed but cannot be read as fp32 parameters
}

float half_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else {  // subnormal: normalise
      int sh = 0;
      while (!(man & 0x400u)) {
        man <<= 1;
        ++sh;
      }
      man &= 0x3FFu;
      bits = sign | ((uint32_t)(127 - 15 - sh + 1) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7F800000u | (man << 13);
  } else {
    bits = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}

int32_t safetensors_parse(tsd_safetensors* st, const uint8_t* file, int64_t size) {
  if (size < 8) return TSD_ERR_INVALID;
  uint64_t hlen = 0;
  std::memcpy(&hlen, file, 8);
  if (hlen > (uint64_t)(size - 8) || hlen > (1ull << 30)) return TSD_ERR_INVALID;
  st->base = file + 8 + hlen;
  st->data_size = size - 8 - (int64_t)hlen;
  Json j{reinterpret_cast<const char*>(file + 8), reinterpret_cast<const char*>(file + 8 + hlen)};
  if (!j.eat('{')) return TSD_ERR_INVALID;
  if (j.eat('}')) return TSD_OK;
  do {
    std::string key;
    if (!j.str(key) || !j.eat(':')) return TSD_ERR_INVALID;
    if (key == "__metadata__") {
      if (!j.skip_value()) return TSD_ERR_INVALID;
      continue;
    }
    tsd_safetensors::Entry en;
    en.name = key;
    bool have_d = false, have_s = false, have_o = false;
    if (!j.eat('{')) return TSD_ERR_INVALID;
    do {
      std::string k;
      if (!j.str(k) || !j.eat(':')) return TSD_ERR_INVALID;
      if (k == "dtype") have_d = j.str(en.dtype);
      else if (k == "shape") have_s = j.int_array(en.shape);
      else if (k == "data_offsets") {
        std::vector<int64_t> o;
        have_o = j.int_array(o) && o.size() == 2;
        if (have_o) {
          en.begin = o[0];
          en.end = o[1];
        }
      } else if (!j.skip_value()) return TSD_ERR_INVALID;
      if (!j.ok) return TSD_ERR_INVALID;
    } while (j.eat(','));
    if (!j.eat('}') || !have_d || !have_s || !have_o) return TSD_ERR_INVALID;
    en.numel = 1;
    for (int64_t d : en.shape) {
      if (d < 0 || (d > 0 && en.numel > INT64_MAX / d)) return TSD_ERR_INVALID;
      en.numel *= d;
    }
    if (en.begin < 0 || en.end < en.begin || en.end > st->data_size) return TSD_ERR_INVALID;
    const int es = dtype_size(en.dtype);
    if (es && en.end - en.begin != en.numel * es) return TSD_ERR_INVALID;
    st->entries.push_back(std::move(en));
  } while (j.eat(','));
  if (!j.eat('}')) return TSD_ERR_INVALID;
  return TSD_OK;
}

}  // namespace

extern "C" {

int32_t tsd_safetensors_open(const char* path, tsd_safetensors** out) {
  if (!path || !out) return TSD_ERR_INVALID;
  *out = nullptr;
  const int fd = ::open(path, O_RDONLY);
  if (fd < 0) return TSD_ERR_INVALID;
  struct stat sb;
  if (fstat(fd, &sb) != 0 || sb.st_size < 8) {
    ::close(fd);
    return TSD_ERR_INVALID;
  }
  void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (m == MAP_FAILED) return TSD_ERR_OOM;
  tsd_safetensors* st = new (std::nothrow) tsd_safetensors();
  if (!st) {
    munmap(m, (size_t)sb.st_size);
    return TSD_ERR_OOM;
  }
  st->map = m;
  st->map_size = (size_t)sb.st_size;
  const int32_t rc = safetensors_parse(st, static_cast<const uint8_t*>(m), (int64_t)sb.st_size);
  if (rc) {
    munmap(m, st->map_size);
    delete st;
    return rc;
  }
  *out = st;
  return TSD_OK;
}

int32_t tsd_safetensors_from_memory(const void* buf, int64_t size, tsd_safetensors** out) {
  if (!buf || !out || size < 8) return TSD_ERR_INVALID;
  *out = nullptr;
  tsd_safetensors* st = new (std::nothrow) tsd_safetensors();
  if (!st) return TSD_ERR_OOM;
  st->owned.assign(static_cast<const uint8_t*>(buf), static_cast<const uint8_t*>(buf) + size);
  const int32_t rc = safetensors_parse(st, st->owned.data(), size);
  if (rc) {
    delete st;
    return rc;
  }
  *out = st;
  return TSD_OK;
}

int32_t tsd_safetensors_close(tsd_safetensors* st) {
  if (!st) return TSD_ERR_INVALID;
  if (st->map) munmap(st->map, st->map_size);
  delete st;
  return TSD_OK;
}
int32_t tsd_safetensors_count(const tsd_safetensors* st) { return st ? (int32_t)st->entries.size() : 0; }
const char* tsd_safetensors_name(const tsd_safetensors* st, int32_t i) {
  return (st && i >= 0 && i < (int32_t)st->entries.size()) ? st->entries[i].name.c_str() : nullptr;
}
int32_t tsd_safetensors_find(const tsd_safetensors* st, const char* name) {
  if (!st || !name) return -1;
  for (size_t i = 0; i < st->entries.size(); ++i)
    if (st->entries[i].name == name) return (int32_t)i;
  return -1;
}
int32_t tsd_safetensors_info(const tsd_safetensors* st, int32_t i, char dtype[8], int32_t* rank, int64_t shape[8],
                             int64_t* numel) {
  if (!st || i < 0 || i >= (int32_t)st->entries.size()) return TSD_ERR_INVALID;
  const tsd_safetensors::Entry& en = st->entries[i];
  if (dtype) {
    std::memset(dtype, 0, 8);
    std::strncpy(dtype, en.dtype.c_str(), 7);
  }
  if (rank) *rank = (int32_t)en.shape.size();
  if (shape)
    for (size_t d = 0; d < en.shape.size() && d < 8; ++d) shape[d] = en.shape[d];
  if (numel) *numel = en.numel;
  return en.shape.size() > 8 ? TSD_ERR_INVALID : TSD_OK;
}
int32_t tsd_safetensors_read_f32(const tsd_safetensors* st, int32_t i, float* out, int64_t cap) {
  if (!st || !out || i < 0 || i >= (int32_t)st->entries.size()) return TSD_ERR_INVALID;
  const tsd_safetensors::Entry& en = st->entries[i];
  if (cap < en.numel) return TSD_ERR_OOM;
  const uint8_t* src = st->base + en.begin;
  if (en.dtype == "F32") {
    std::memcpy(out, src, (size_t)en.numel * 4);
  } else if (en.dtype == "F16") {
    for (int64_t k = 0; k < en.numel; ++k) {
      uint16_t h;
      std::memcpy(&h, src + 2 * k, 2);
      out[k] = half_to_float(h);
    }
  } else if (en.dtype == "BF16") {
    for (int64_t k = 0; k < en.numel; ++k) {
      uint16_t h;
      std::memcpy(&h, src + 2 * k, 2);
      const uint32_t bits = (uint32_t)h << 16;
      std::memcpy(out + k, &bits, 4);
    }
  } else if (en.dtype == "F64") {
    for (int64_t k = 0; k < en.numel; ++k) {
      double d;
      std::memcpy(&d, src + 8 * k, 8);
      out[k] = (float)d;
    }
  } else {
    return TSD_ERR_INVALID;
  }
  return TSD_OK;
}

}  // extern "C"
