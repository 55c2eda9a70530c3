// Batched .c2df header parser (host side of north-star item (a)).
//
// Replaces the per-file Python loop of build_index_from_c2df_dir (reference
// src/build.py:80-88): for every file it walks the little-endian TLV container written by
// pack_c2df (src/filemaker.py:75-100; reader :137-173), locates the `clip_stream` (BYTES)
// and `clip_meta` (JSON) entries, reads clip_meta.dim, and zstd-decodes the stream into one
// u8 row.  The rows then cross PCIe at 1 byte/element and are expanded on the device (K1).
//
// libzstd has no headers in the image, so the four entry points used are resolved with
// dlopen("libzstd.so.1") and hand-declared prototypes (stable public ABI since zstd 1.0).
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sgic.h"
#include "zstd_lit.cuh"

namespace sgic {
void set_error(const std::string& msg);  // defined in sgic_api.cu

namespace {

constexpr size_t kZlMaxFrameBytes = 2560;  // = kZlMaxFrame of the decode kernel (ingest.cuh)

struct ZstdApi {
  void* lib = nullptr;
  void* (*createDCtx)() = nullptr;
  size_t (*freeDCtx)(void*) = nullptr;
  size_t (*decompressDCtx)(void*, void*, size_t, const void*, size_t) = nullptr;
  unsigned long long (*getFrameContentSize)(const void*, size_t) = nullptr;
  unsigned (*isError)(size_t) = nullptr;
  bool ok = false;
};

const ZstdApi& zstd_api() {
  static ZstdApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libzstd.so.1", "libzstd.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (api.lib) break;
    }
    if (!api.lib) return;
    api.createDCtx = reinterpret_cast<void* (*)()>(dlsym(api.lib, "ZSTD_createDCtx"));
    api.freeDCtx = reinterpret_cast<size_t (*)(void*)>(dlsym(api.lib, "ZSTD_freeDCtx"));
    api.decompressDCtx = reinterpret_cast<size_t (*)(void*, void*, size_t, const void*, size_t)>(
        dlsym(api.lib, "ZSTD_decompressDCtx"));
    api.getFrameContentSize =
        reinterpret_cast<unsigned long long (*)(const void*, size_t)>(dlsym(api.lib, "ZSTD_getFrameContentSize"));
    api.isError = reinterpret_cast<unsigned (*)(size_t)>(dlsym(api.lib, "ZSTD_isError"));
    api.ok = api.createDCtx && api.freeDCtx && api.decompressDCtx && api.getFrameContentSize && api.isError;
  });
  return api;
}

// type codes, src/filemaker.py:4-11
enum : uint8_t { T_BYTES = 0, T_STR = 1, T_INT = 2, T_FLOAT = 3, T_JSON = 4, T_NP = 5, T_NONE = 6, T_BOOL = 7 };

// ---- CPython's 'utf-8' codec with errors='strict' (filemaker.py:149,157,118,122,126: every header, key, string
// and JSON payload goes through bytes.decode('utf-8'), and UnicodeDecodeError makes build.py:87-88 skip the file):
// no overlong forms, no surrogates, nothing above U+10FFFF, no stray or missing continuation bytes.
bool utf8_valid(const uint8_t* s, size_t n) {
  {  // almost everything here is ASCII: one OR over the bytes (vectorised by the compiler) settles that
    uint8_t acc = 0;
    for (size_t k = 0; k < n; ++k) acc |= s[k];
    if (!(acc & 0x80)) return true;
  }
  size_t i = 0;
  while (i < n) {
    if (n - i >= 8) {  // ASCII, eight bytes at a time
      uint64_t w;
      std::memcpy(&w, s + i, 8);
      if (!(w & 0x8080808080808080ull)) {
        i += 8;
        continue;
      }
    }
    const uint8_t c = s[i];
    if (c < 0x80) {
      ++i;
      continue;
    }
    auto cont = [&](size_t k) { return i + k < n && (s[i + k] & 0xC0) == 0x80; };
    if (c < 0xC2) return false;  // continuation byte / overlong two-byte lead
    if (c < 0xE0) {
      if (!cont(1)) return false;
      i += 2;
    } else if (c < 0xF0) {
      if (!cont(1) || !cont(2)) return false;
      if (c == 0xE0 && s[i + 1] < 0xA0) return false;   // overlong
      if (c == 0xED && s[i + 1] >= 0xA0) return false;  // UTF-16 surrogates
      i += 3;
    } else if (c < 0xF5) {
      if (!cont(1) || !cont(2) || !cont(3)) return false;
      if (c == 0xF0 && s[i + 1] < 0x90) return false;   // overlong
      if (c == 0xF4 && s[i + 1] >= 0x90) return false;  // above U+10FFFF
      i += 4;
    } else {
      return false;
    }
  }
  return true;
}

// ---- json.loads of CPython (Lib/json/decoder.py + Modules/_json.c, strict=True), as a validator ------------
// The reference parses the header, every JSON entry and clip_meta with it (filemaker.py:149,126), so a file is kept
// only if all of them are texts json.loads accepts: RFC 8259 plus NaN / Infinity / -Infinity, whitespace is exactly
// " \t\n\r", numbers follow the scanner's own grammar (-?(0|[1-9]\d*)(\.\d+)?([eE][-+]?\d+)? — "1." is the number 1
// followed by garbage), strings may not hold raw control characters and only the escapes \" \\ \/ \b \f \n \r \t
// \uXXXX, nothing but whitespace may follow the value, a leading BOM is refused, an integer literal of more than 4300
// digits raises (sys.int_max_str_digits), nesting beyond the interpreter's recursion budget raises.
// While it validates it remembers where the value of the LAST top-level "dim" member of an object lies (a dict keeps
// the last duplicate; the key is compared after unescaping).
struct JsonDim {
  bool top_object = false;  // the text is an object (anything else has no .get — or is falsy and becomes {})
  bool has = false;         // the object has a "dim" member; its value is s[vb, ve)
  size_t vb = 0, ve = 0;
};

struct Json {
  const char* s;
  size_t n;
  size_t i = 0;
  // Nesting json.loads accepts before it raises RecursionError (CPython 3.12, the reference's Dockerfile base:
  // the C scanner's recursion budget; measured 9997 with this interpreter).  build.py:87-88 turns that exception
  // into [SKIP]; deeper payloads are refused here the same way instead of being followed.
  static constexpr size_t kMaxDepth = 9990;
  static constexpr size_t kMaxIntDigits = 4300;

  static bool digit(char c) { return c >= '0' && c <= '9'; }
  void ws() {
    while (i < n && static_cast<unsigned char>(s[i]) <= ' ' && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
  }
  // 1 for the bytes that end a run of plain string characters: '"', '\\' and everything below 0x20
  static const uint8_t* stop_table() {
    static const struct Table {
      uint8_t t[256];
      Table() {
        for (int c = 0; c < 256; ++c) t[c] = (c < 0x20 || c == '"' || c == '\\') ? 1 : 0;
      }
    } table;
    return table.t;
  }
  // at the opening quote; on success i is past the closing quote and [*b, *e) is the raw text between the quotes
  bool string(size_t* b, size_t* e, bool* escaped) {
    if (i >= n || s[i] != '"') return false;
    ++i;
    *b = i;
    *escaped = false;
    const uint8_t* stop = stop_table();
    for (;;) {
      // plain characters (bytes >= 0x80 are parts of characters the UTF-8 pass has already accepted)
      while (i < n && !stop[static_cast<unsigned char>(s[i])]) ++i;
      if (i >= n) return false;  // unterminated
      const unsigned char c = static_cast<unsigned char>(s[i]);
      if (c == '"') break;
      if (c < 0x20) return false;  // "Invalid control character"
      if (c != '\\') {
        ++i;
        continue;
      }
      *escaped = true;
      if (++i >= n) return false;
      const char x = s[i];
      if (x == 'u') {
        if (i + 4 >= n) return false;  // four hex digits and at least the closing quote must follow
        for (int k = 1; k <= 4; ++k) {
          const char h = s[i + k];
          if (!(digit(h) || (h >= 'a' && h <= 'f') || (h >= 'A' && h <= 'F'))) return false;
        }
        i += 5;
      } else if (x == '"' || x == '\\' || x == '/' || x == 'b' || x == 'f' || x == 'n' || x == 'r' || x == 't') {
        ++i;
      } else {
        return false;  // "Invalid \escape"
      }
    }
    *e = i;
    ++i;
    return true;
  }
  bool word(const char* w, size_t len) {
    if (n - i < len || std::memcmp(s + i, w, len) != 0) return false;
    i += len;
    return true;
  }
  // _match_number_unicode of Modules/_json.c
  bool number() {
    size_t j = i;
    if (s[j] == '-') {
      if (++j >= n) return false;
    }
    const size_t int_b = j;
    if (s[j] >= '1' && s[j] <= '9') {
      ++j;
      while (j < n && digit(s[j])) ++j;
    } else if (s[j] == '0') {
      ++j;
    } else {
      return false;
    }
    const size_t int_digits = j - int_b;
    bool is_float = false;
    if (j + 1 < n && s[j] == '.' && digit(s[j + 1])) {
      is_float = true;
      j += 2;
      while (j < n && digit(s[j])) ++j;
    }
    if (j + 1 < n && (s[j] == 'e' || s[j] == 'E')) {
      size_t k = j + 1;
      if (k + 1 < n && (s[k] == '-' || s[k] == '+')) ++k;
      const size_t d0 = k;
      while (k < n && digit(s[k])) ++k;
      if (k > d0) {
        is_float = true;
        j = k;
      }
    }
    if (!is_float && int_digits > kMaxIntDigits) return false;  // ValueError: exceeds the limit for int conversion
    i = j;
    return true;
  }
  // a value that is not a container, at s[i] (i < n)
  bool scalar() {
    size_t b, e;
    bool esc;
    switch (s[i]) {
      case '"': return string(&b, &e, &esc);
      case 'n': return word("null", 4);
      case 't': return word("true", 4);
      case 'f': return word("false", 5);
      case 'N': return word("NaN", 3);
      case 'I': return word("Infinity", 8);
      case '-':
        if (n - i >= 9 && std::memcmp(s + i, "-Infinity", 9) == 0) {
          i += 9;
          return true;
        }
        return number();
      default: return number();
    }
  }
  static void append_utf8(std::string* out, uint32_t cp) {
    if (cp < 0x80) {
      out->push_back(static_cast<char>(cp));
    } else if (cp < 0x800) {
      out->push_back(static_cast<char>(0xC0 | (cp >> 6)));
      out->push_back(static_cast<char>(0x80 | (cp & 0x3F)));
    } else if (cp < 0x10000) {
      out->push_back(static_cast<char>(0xE0 | (cp >> 12)));
      out->push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
      out->push_back(static_cast<char>(0x80 | (cp & 0x3F)));
    } else {
      out->push_back(static_cast<char>(0xF0 | (cp >> 18)));
      out->push_back(static_cast<char>(0x80 | ((cp >> 12) & 0x3F)));
      out->push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
      out->push_back(static_cast<char>(0x80 | (cp & 0x3F)));
    }
  }
  // the str a validated string token [b, e) stands for, as UTF-8 (a lone surrogate comes out as its 3-byte form:
  // it can never be part of a key or a number we compare against)
  static std::string unescape(const char* s, size_t b, size_t e) {
    std::string out;
    out.reserve(e - b);
    auto hex4 = [&](size_t p) {
      uint32_t v = 0;
      for (int k = 0; k < 4; ++k) {
        const char h = s[p + k];
        v = v * 16 + static_cast<uint32_t>(digit(h) ? h - '0' : (h | 0x20) - 'a' + 10);
      }
      return v;
    };
    for (size_t p = b; p < e;) {
      if (s[p] != '\\') {
        out.push_back(s[p++]);
        continue;
      }
      const char x = s[p + 1];
      p += 2;
      switch (x) {
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'n': out.push_back('\n'); break;
        case 'r': out.push_back('\r'); break;
        case 't': out.push_back('\t'); break;
        case 'u': {
          uint32_t cp = hex4(p);
          p += 4;
          if (cp >= 0xD800 && cp < 0xDC00 && p + 6 <= e && s[p] == '\\' && s[p + 1] == 'u') {
            const uint32_t lo = hex4(p + 2);
            if (lo >= 0xDC00 && lo < 0xE000) {
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
              p += 6;
            }
          }
          append_utf8(&out, cp);
          break;
        }
        default: out.push_back(x);  // '"', '\\', '/'
      }
    }
    return out;
  }
  // `"key" :` of an object member; *is_dim: the key is "dim"
  bool member_key(bool* is_dim) {
    ws();
    size_t b, e;
    bool esc;
    if (!string(&b, &e, &esc)) return false;
    *is_dim = esc ? unescape(s, b, e) == "dim" : (e - b == 3 && std::memcmp(s + b, "dim", 3) == 0);
    ws();
    if (i >= n || s[i] != ':') return false;
    ++i;
    ws();
    return true;
  }
  // The whole text.  Iterative (the payload is untrusted and may be 4 GB of '['): the open containers live on an
  // explicit stack of one bit each (object / array) — 64 levels inline, deeper ones on the heap.
  bool parse(JsonDim* cap) {
    if (n >= 3 && std::memcmp(s, "\xEF\xBB\xBF", 3) == 0) return false;  // "Unexpected UTF-8 BOM"
    uint64_t small = 0;
    struct Deep {  // levels beyond 64: allocated only by a file that nests that far
      std::vector<uint8_t>* v = nullptr;
      ~Deep() { delete v; }
    } deep;
    size_t depth = 0;
    auto push = [&](bool obj) {
      if (depth < 64) small = (small & ~(1ull << depth)) | (static_cast<uint64_t>(obj) << depth);
      else {
        if (!deep.v) deep.v = new std::vector<uint8_t>();
        if (deep.v->size() <= depth - 64) deep.v->resize(std::max<size_t>(256, 2 * (depth - 64 + 1)));
        (*deep.v)[depth - 64] = obj;
      }
      ++depth;
    };
    auto top_is_obj = [&]() -> bool {
      return depth <= 64 ? ((small >> (depth - 1)) & 1u) != 0 : (*deep.v)[depth - 1 - 64] != 0;
    };
    bool dim_pending = false;  // the value being read is the one of a top-level "dim"
    size_t dim_vb = 0;
    ws();
    for (;;) {
      if (i >= n) return false;  // "Expecting value"
      const char c = s[i];
      if (c == '{' || c == '[') {
        if (depth >= kMaxDepth) return false;
        const char close = (c == '{') ? '}' : ']';
        if (depth == 0 && c == '{') cap->top_object = true;
        ++i;
        ws();
        if (i < n && s[i] == close) {
          ++i;
        } else {
          push(c == '{');
          if (c == '{') {
            bool is_dim = false;
            if (!member_key(&is_dim)) return false;
            if (depth == 1 && is_dim) {
              dim_pending = true;
              dim_vb = i;
            }
          }
          continue;  // the container's first value
        }
      } else if (!scalar()) {
        return false;
      }
      // a value is complete: close finished containers, or step to the next member
      for (;;) {
        if (dim_pending && depth == 1) {
          cap->has = true;
          cap->vb = dim_vb;
          cap->ve = i;
          dim_pending = false;
        }
        if (depth == 0) {
          ws();
          return i == n;  // "Extra data" otherwise
        }
        ws();
        if (i >= n) return false;
        if (s[i] == ',') {
          ++i;
          if (top_is_obj()) {
            bool is_dim = false;
            if (!member_key(&is_dim)) return false;
            if (depth == 1 && is_dim) {
              dim_pending = true;
              dim_vb = i;
            }
          } else {
            ws();
          }
          break;
        }
        if (s[i] == (top_is_obj() ? '}' : ']')) {
          ++i;
          --depth;
          continue;
        }
        return false;
      }
    }
  }
};

// The files of one corpus carry the same clip_meta text (model, dim, quantiser) and, for images of one size, the same
// header text: a worker remembers the last text it judged (it lies in the caller's blob, which outlives the call) and
// settles a repeat with one memcmp.
struct JsonMemo {
  const uint8_t* p = nullptr;
  size_t n = 0;
  bool ok = false;
  JsonDim dim;
};

// bytes.decode('utf-8') + json.loads
bool json_loads_ok(const uint8_t* p, size_t n, JsonDim* cap, JsonMemo* memo = nullptr) {
  if (memo && memo->p && memo->n == n && std::memcmp(memo->p, p, n) == 0) {
    *cap = memo->dim;
    return memo->ok;
  }
  *cap = JsonDim();
  bool ok = utf8_valid(p, n);
  if (ok) {
    Json j{reinterpret_cast<const char*>(p), n};
    ok = j.parse(cap);
  }
  if (memo) {
    memo->p = p;
    memo->n = n;
    memo->ok = ok;
    memo->dim = *cap;
  }
  return ok;
}

struct WalkMemo {
  JsonMemo header, meta;
};

// Byte length of the whitespace character (str.isspace) at p, 0 if there is none: what int(str) strips.
size_t py_space_len(const uint8_t* p, size_t n) {
  if (n == 0) return 0;
  const uint8_t c = p[0];
  if (c == ' ' || (c >= 0x09 && c <= 0x0D)) return 1;  // (0x1C-0x1F are str.isspace() but int() keeps ASCII as it is)
  if (c == 0xC2 && n >= 2 && (p[1] == 0x85 || p[1] == 0xA0)) return 2;
  if (n >= 3) {
    if (c == 0xE1 && p[1] == 0x9A && p[2] == 0x80) return 3;                             // U+1680
    if (c == 0xE2 && p[1] == 0x80 && ((p[2] >= 0x80 && p[2] <= 0x8A) || p[2] == 0xA8 || p[2] == 0xA9 || p[2] == 0xAF))
      return 3;                                                                          // U+2000-200A, 2028, 2029, 202F
    if (c == 0xE2 && p[1] == 0x81 && p[2] == 0x9F) return 3;                             // U+205F
    if (c == 0xE3 && p[1] == 0x80 && p[2] == 0x80) return 3;                             // U+3000
  }
  return 0;
}

// Decimal digit value of a non-ASCII code point (Py_UNICODE_TODECIMAL), -1 if it has none: the Nd characters of
// Unicode 15.0 (CPython 3.12), every script's digits being ten consecutive code points starting at its zero.
int unicode_decimal(uint32_t cp) {
  static const uint32_t zeros[] = {
      0x660,   0x6f0,   0x7c0,   0x966,   0x9e6,   0xa66,   0xae6,   0xb66,   0xbe6,   0xc66,   0xce6,   0xd66,   0xde6,   0xe50,
      0xed0,   0xf20,   0x1040,  0x1090,  0x17e0,  0x1810,  0x1946,  0x19d0,  0x1a80,  0x1a90,  0x1b50,  0x1bb0,  0x1c40,  0x1c50,
      0xa620,  0xa8d0,  0xa900,  0xa9d0,  0xa9f0,  0xaa50,  0xabf0,  0xff10,  0x104a0, 0x10d30, 0x11066, 0x110f0, 0x11136, 0x111d0,
      0x112f0, 0x11450, 0x114d0, 0x11650, 0x116c0, 0x11730, 0x118e0, 0x11950, 0x11c50, 0x11d50, 0x11da0, 0x11f50, 0x16a60, 0x16ac0,
      0x16b50, 0x1d7ce, 0x1d7d8, 0x1d7e2, 0x1d7ec, 0x1d7f6, 0x1e140, 0x1e2f0, 0x1e4f0, 0x1e950, 0x1fbf0};
  for (uint32_t z : zeros)
    if (cp >= z && cp < z + 10) return static_cast<int>(cp - z);
  return -1;
}

// int(str) for base 10: surrounding whitespace, an optional sign, decimal digits (of any script) with single
// underscores between them.
bool py_int_from_str(const std::string& v, long long* out) {
  const uint8_t* p = reinterpret_cast<const uint8_t*>(v.data());
  size_t a = 0, b = v.size();
  for (size_t k; a < b && (k = py_space_len(p + a, b - a)) != 0;) a += k;
  for (;;) {  // trailing whitespace: step back over one character at a time
    size_t k = 0;
    for (size_t len = 1; len <= 3 && len <= b - a; ++len)
      if (py_space_len(p + b - len, len) == len) k = len;
    if (!k) break;
    b -= k;
  }
  bool neg = false;
  if (a < b && (p[a] == '+' || p[a] == '-')) neg = p[a++] == '-';
  if (a >= b) return false;
  long long sv = 0;
  int digits = 0;
  bool prev_us = true;  // an underscore may not lead
  for (size_t t = a; t < b;) {
    int dv;
    if (p[t] == '_') {
      if (prev_us) return false;
      prev_us = true;
      ++t;
      continue;
    }
    if (p[t] < 0x80) {
      if (p[t] < '0' || p[t] > '9') return false;
      dv = p[t] - '0';
      ++t;
    } else {  // (valid UTF-8: it came out of a validated JSON string)
      const int len = p[t] >= 0xF0 ? 4 : p[t] >= 0xE0 ? 3 : 2;
      if (t + static_cast<size_t>(len) > b) return false;
      uint32_t cp = p[t] & (0xFFu >> (len + 1));
      for (int k = 1; k < len; ++k) cp = (cp << 6) | (p[t + k] & 0x3Fu);
      dv = unicode_decimal(cp);
      if (dv < 0) return false;
      t += static_cast<size_t>(len);
    }
    prev_us = false;
    if (sv != 0 || dv != 0) ++digits;
    if (digits > 15) return false;  // far beyond any dimension (Python would carry on; the file is skipped either way)
    sv = sv * 10 + dv;
  }
  if (prev_us) return false;  // ... or trail
  *out = neg ? -sv : sv;
  return true;
}

// int(meta.get('dim', 0)) for the clip_meta text (search.py:30).  Returns false when the reference would have raised
// (the text is not a dict, the value is not convertible); *dim = 0 when the key is absent.
bool clip_meta_dim(const char* s, const JsonDim& cap, long long* dim) {
  *dim = 0;
  if (!cap.top_object) return false;  // falsy values become {} (dim 0 -> ValueError), the others have no .get
  if (!cap.has) return true;
  const size_t vb = cap.vb, ve = cap.ve;
  const char c = s[vb];
  if (c == '"') {
    const std::string v = Json::unescape(s, vb + 1, ve - 1);
    return py_int_from_str(v, dim);
  }
  if (c == 't') {
    *dim = 1;
    return true;
  }
  if (c == 'f') return true;                                       // int(False) == 0
  if (c == 'n' || c == 'N' || c == 'I' || c == '{' || c == '[') return false;  // None / nan / inf / containers
  if (c == '-' && ve - vb > 1 && s[vb + 1] == 'I') return false;
  bool is_float = false;
  for (size_t t = vb; t < ve; ++t) is_float = is_float || s[t] == '.' || s[t] == 'e' || s[t] == 'E';
  if (!is_float) {
    // what the reference writes is a short run of digits ("dim": 512)
    size_t t = vb;
    const bool neg = s[t] == '-';
    if (neg) ++t;
    if (ve - t > 15) return false;  // cannot be a dimension
    long long iv = 0;
    for (; t < ve; ++t) iv = iv * 10 + (s[t] - '0');
    *dim = neg ? -iv : iv;
    return true;
  }
  const std::string v(s + vb, ve - vb);
  char* end = nullptr;
  const double dv = std::strtod(v.c_str(), &end);
  if (end == v.c_str() || *end != '\0' || dv != dv || dv > 9e15 || dv < -9e15) return false;  // int(inf) raises
  *dim = static_cast<long long>(dv);  // int() truncates toward zero
  return true;
}

// numpy's itemsize for the dtype strings np.dtype() understands among the ones an array's dtype.str / dtype.name can
// be (filemaker.py:25,128-129); 0 = "data type not understood" (TypeError in the reference).  Structured, sub-array
// ("4i4"), datetime and object dtypes are not recognised: a file carrying such an array is skipped.
size_t np_itemsize(const uint8_t* p, size_t n) {
  // array-protocol type strings first ("<i4", "|u1", "<f4": what dtype.str gives and the reference writes)
  size_t i = 0;
  if (i < n && (p[i] == '<' || p[i] == '>' || p[i] == '=' || p[i] == '|')) ++i;
  if (i >= n) return 0;
  const char kind = static_cast<char>(p[i++]);
  if (i == n) {  // one-character type codes
    switch (kind) {
      case '?': case 'b': case 'B': return 1;
      case 'h': case 'H': case 'e': return 2;
      case 'i': case 'I': case 'f': return 4;
      case 'l': case 'L': case 'q': case 'Q': case 'd': case 'p': case 'P': case 'F': return 8;
      case 'g': case 'D': return 16;
      case 'G': return 32;
      default: return 0;
    }
  }
  size_t size = 0;
  bool digits = true;
  for (size_t k = i; k < n && digits; ++k) {
    digits = p[k] >= '0' && p[k] <= '9' && size <= (1u << 24);
    size = size * 10 + static_cast<size_t>(p[k] - '0');
  }
  if (digits) {
    switch (kind) {
      case 'b': return size == 1 ? 1 : 0;
      case 'i': case 'u': return (size == 1 || size == 2 || size == 4 || size == 8) ? size : 0;
      case 'f': return (size == 2 || size == 4 || size == 8 || size == 16) ? size : 0;
      case 'c': return (size == 8 || size == 16 || size == 32) ? size : 0;
      case 'S': case 'V': return size;         // (size 0: np.frombuffer refuses an itemsize of zero)
      case 'U': return size * 4;
      default: return 0;
    }
  }
  static const struct { const char* name; size_t size; } names[] = {
      {"bool", 1}, {"int8", 1}, {"uint8", 1}, {"int16", 2}, {"uint16", 2}, {"int32", 4}, {"uint32", 4}, {"int64", 8},
      {"uint64", 8}, {"float16", 2}, {"float32", 4}, {"float64", 8}, {"complex64", 8}, {"complex128", 16},
      {"int", 8}, {"float", 8}, {"double", 8}, {"single", 4}, {"half", 2}, {"byte", 1}, {"ubyte", 1}, {"short", 2},
      {"ushort", 2}, {"intc", 4}, {"uintc", 4}, {"longlong", 8}, {"ulonglong", 8}, {"complex", 16}, {"bool_", 1}};
  for (const auto& e : names)
    if (std::strlen(e.name) == n && std::memcmp(e.name, p, n) == 0) return e.size;
  return 0;
}

// _load_entry for an array (filemaker.py:127-135): dtype, ndim, shape, data length, data; np.frombuffer needs a
// whole number of items and .reshape(shape) exactly as many as the shape holds.  Slices clamp, unpack_from raises.
bool np_entry_loads(const uint8_t* p, size_t n) {
  if (n < 1) return false;
  const size_t dt_len = p[0];
  const size_t dt_have = std::min(dt_len, n - 1);
  if (!utf8_valid(p + 1, dt_have)) return false;
  const size_t item = np_itemsize(p + 1, dt_have);
  if (item == 0) return false;
  size_t off = 1 + dt_len;
  if (off + 1 > n) return false;
  const size_t ndim = p[off++];
  if (ndim > 64) return false;  // numpy's maximum number of dimensions (numpy >= 2)
  if (off + 4 * ndim + 4 > n) return false;
  auto u32 = [&](size_t o) {
    return static_cast<uint64_t>(p[o]) | (static_cast<uint64_t>(p[o + 1]) << 8) | (static_cast<uint64_t>(p[o + 2]) << 16) |
           (static_cast<uint64_t>(p[o + 3]) << 24);
  };
  unsigned __int128 count = 1;
  for (size_t k = 0; k < ndim; ++k, off += 4) {
    count *= u32(off);
    if (count > (static_cast<unsigned __int128>(1) << 70)) count = static_cast<unsigned __int128>(1) << 70;  // far beyond any buffer, and never 0 again... unless a later dimension is 0
  }
  const uint64_t data_len = u32(off);
  off += 4;
  const size_t have = static_cast<size_t>(std::min<uint64_t>(data_len, n - off));
  if (have % item != 0) return false;  // "buffer size must be a multiple of element size"
  return count == static_cast<unsigned __int128>(have / item);
}

struct Found {
  const uint8_t* stream = nullptr;
  size_t stream_len = 0;
  bool has_stream = false, stream_is_bytes = false;
  bool has_meta = false;
  uint8_t meta_type = T_NONE;
  const char* meta = nullptr;  // clip_meta's JSON text and where its "dim" lies
  JsonDim meta_dim;
};

// unpack_c2df (filemaker.py:137-173) as a validator.  The reference LOADS every entry as it walks
// (_load_entry, :102-135) and parses the header, so each of them must load for the file to be kept; Python slices
// clamp at the end of the data where struct.unpack_from raises, which is why an entry may be cut off by the end of the
// file and still load (a BYTES payload, say) while a missing length field is an error.  A later entry of the same key
// replaces an earlier one (dict).
int walk_one(const uint8_t* data, size_t n, Found* f, WalkMemo* memo = nullptr) {
  if (n < 4 || std::memcmp(data, "C2DF", 4) != 0) return SGIC_C2DF_BAD_MAGIC;
  uint64_t off = 4;  // may run past n: slices clamp, the next fixed-size read fails
  auto rd = [&](unsigned bytes, uint64_t* v) -> bool {
    if (off + bytes > n) return false;
    uint64_t x = 0;
    for (unsigned k = 0; k < bytes; ++k) x |= static_cast<uint64_t>(data[off + k]) << (8 * k);
    *v = x;
    off += bytes;
    return true;
  };
  auto slice = [&](uint64_t len, const uint8_t** p, size_t* have) {  // data[off:off+len]; off += len
    *p = data + std::min<uint64_t>(off, n);
    *have = off >= n ? 0 : static_cast<size_t>(std::min<uint64_t>(len, n - off));
    off += len;
  };
  uint64_t ver, hlen, n_items;
  if (!rd(2, &ver) || !rd(4, &hlen)) return SGIC_C2DF_TRUNCATED;
  if (hlen > 0) {
    const uint8_t* hp;
    size_t hh;
    slice(hlen, &hp, &hh);
    JsonDim ignore;
    if (!json_loads_ok(hp, hh, &ignore, memo ? &memo->header : nullptr)) return SGIC_C2DF_BAD_ENTRY;
  }
  if (!rd(4, &n_items)) return SGIC_C2DF_TRUNCATED;
  for (uint64_t it = 0; it < n_items; ++it) {
    uint64_t klen, t;
    if (!rd(2, &klen)) return SGIC_C2DF_TRUNCATED;
    const uint8_t* key;
    size_t key_have;
    slice(klen, &key, &key_have);
    if (!utf8_valid(key, key_have)) return SGIC_C2DF_BAD_ENTRY;
    if (!rd(1, &t)) return SGIC_C2DF_TRUNCATED;
    const uint8_t* payload = nullptr;
    size_t plen = 0;
    if (t == T_INT || t == T_FLOAT) {
      slice(8, &payload, &plen);
      if (plen < 8) return SGIC_C2DF_TRUNCATED;  // struct.error in _load_entry
    } else if (t == T_BOOL) {
      slice(1, &payload, &plen);
      if (plen < 1) return SGIC_C2DF_TRUNCATED;
    } else if (t == T_NONE) {
      plen = 0;
    } else {
      uint64_t L;
      if (!rd(4, &L)) return SGIC_C2DF_TRUNCATED;
      slice(L, &payload, &plen);
      if (t > T_BOOL) return SGIC_C2DF_BAD_TYPE;
    }
    const bool is_stream = (key_have == 11 && std::memcmp(key, "clip_stream", 11) == 0);
    const bool is_meta = (key_have == 9 && std::memcmp(key, "clip_meta", 9) == 0);
    // BYTES / STR / JSON payloads carry a second u32 length in front of the data
    const uint8_t* inner = payload;
    size_t inner_len = plen;
    JsonDim jd;
    if (t == T_BYTES || t == T_STR || t == T_JSON) {
      if (plen < 4) return SGIC_C2DF_TRUNCATED;
      const uint32_t L = static_cast<uint32_t>(payload[0]) | (static_cast<uint32_t>(payload[1]) << 8) |
                         (static_cast<uint32_t>(payload[2]) << 16) | (static_cast<uint32_t>(payload[3]) << 24);
      inner = payload + 4;
      inner_len = (L <= plen - 4) ? L : plen - 4;  // Python slicing clamps
      if (t == T_STR && !utf8_valid(inner, inner_len)) return SGIC_C2DF_BAD_ENTRY;
      if (t == T_JSON && !json_loads_ok(inner, inner_len, &jd, (memo && is_meta) ? &memo->meta : nullptr))
        return SGIC_C2DF_BAD_ENTRY;
    } else if (t == T_NP) {
      if (!np_entry_loads(payload, plen)) return SGIC_C2DF_BAD_ENTRY;
    }
    if (is_stream) {
      f->has_stream = true;
      f->stream_is_bytes = (t == T_BYTES);
      f->stream = inner;
      f->stream_len = inner_len;
    } else if (is_meta) {
      f->has_meta = true;
      f->meta_type = static_cast<uint8_t>(t);
      f->meta = reinterpret_cast<const char*>(inner);
      f->meta_dim = jd;
    }
  }
  return SGIC_C2DF_OK;
}

// dev_frame != nullptr: frames inside the device decoder's profile (csrc/zstd_lit.cuh) are not decoded here;
// their position is reported instead (*dev_frame = stream start, *dev_len = its length) and the row is filled
// in by zstd_lit_decode_kernel.  libzstd stays the judge of everything else.
int parse_one(const ZstdApi& z, void* dctx, const uint8_t* data, size_t n, int want_dim, uint8_t* out_row,
              int32_t* dim_out, std::vector<uint8_t>& scratch, WalkMemo* memo, const uint8_t** dev_frame = nullptr,
              uint32_t* dev_len = nullptr) {
  *dim_out = 0;
  Found f;
  int st = walk_one(data, n, &f, memo);
  if (st != SGIC_C2DF_OK) return st;
  if (!f.has_stream || !f.has_meta) return SGIC_C2DF_NO_CLIP;
  long long dim = 0;
  if (f.meta_type == T_JSON) {
    if (!clip_meta_dim(f.meta, f.meta_dim, &dim)) return SGIC_C2DF_BAD_DIM;
  } else if (f.meta_type == T_NONE) {
    dim = 0;  // None or {} -> dim 0
  } else {
    return SGIC_C2DF_BAD_DIM;  // no .get on bytes / str / int / ndarray
  }
  if (dim <= 0 || dim > (1ll << 24)) return SGIC_C2DF_BAD_DIM;
  *dim_out = static_cast<int32_t>(dim);
  if (!f.stream_is_bytes) return SGIC_C2DF_ZSTD;  // decompress(str) raises
  // python-zstandard's one-shot decompress needs the content size in the frame header
  const unsigned long long fcs = z.getFrameContentSize(f.stream, f.stream_len);
  if (fcs == ~0ull || fcs == ~0ull - 1) return SGIC_C2DF_ZSTD;
  if (fcs > (1ull << 26)) return SGIC_C2DF_ZSTD;
  if (static_cast<long long>(fcs) == dim && dim == want_dim) {
    if (dev_frame != nullptr && f.stream_len <= kZlMaxFrameBytes) {
      zl::FrameInfo fi;
      if (zl::parse_frame(f.stream, static_cast<uint32_t>(f.stream_len), fi) == zl::ZL_OK &&
          fi.content_size == static_cast<uint32_t>(dim)) {
        *dev_frame = f.stream;
        *dev_len = static_cast<uint32_t>(f.stream_len);
        return SGIC_C2DF_OK;
      }
    }
    const size_t r = z.decompressDCtx(dctx, out_row, static_cast<size_t>(dim), f.stream, f.stream_len);
    if (z.isError(r)) return SGIC_C2DF_ZSTD;
    if (static_cast<long long>(r) != dim) return SGIC_C2DF_DIM_MISMATCH;
    return SGIC_C2DF_OK;
  }
  // mismatching sizes: still decode (to report a corrupt frame as such), never into out_row
  scratch.resize(static_cast<size_t>(fcs) + 1);
  const size_t r = z.decompressDCtx(dctx, scratch.data(), scratch.size(), f.stream, f.stream_len);
  if (z.isError(r)) return SGIC_C2DF_ZSTD;
  if (static_cast<long long>(r) != dim) return SGIC_C2DF_DIM_MISMATCH;
  return SGIC_C2DF_WRONG_D;
}

}  // namespace

// frame_off / frame_len (both or neither): per file, the offset into `blob` and length of a clip_stream frame
// left for the device decoder, or -1 / 0 when the row was decoded here.
int c2df_parse_batch(const uint8_t* blob, const int64_t* offsets, int64_t n, int dim, uint8_t* out_u8,
                     int32_t* status_out, int32_t* dim_out, int n_threads, int64_t* frame_off, uint32_t* frame_len) {
  const ZstdApi& z = zstd_api();
  if (!z.ok) {
    set_error("libzstd.so.1 could not be loaded (dlopen) — required for clip_stream decoding");
    return 1;
  }
  if (n < 0 || dim <= 0) {
    set_error("c2df_parse: bad arguments");
    return 1;
  }
  if (n_threads <= 0) n_threads = static_cast<int>(std::thread::hardware_concurrency());
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  if (static_cast<int64_t>(n_threads) > n) n_threads = static_cast<int>(n > 0 ? n : 1);
  std::atomic<int64_t> next{0};
  constexpr int64_t kGrain = 256;
  auto work = [&]() {
    void* dctx = z.createDCtx();
    std::vector<uint8_t> scratch;
    WalkMemo memo;
    for (;;) {
      const int64_t b = next.fetch_add(kGrain);
      if (b >= n) break;
      const int64_t e = (b + kGrain < n) ? b + kGrain : n;
      for (int64_t i = b; i < e; ++i) {
        const int64_t o0 = offsets[i], o1 = offsets[i + 1];
        // The blob is far larger than the caches and the walk is bound by memory latency, not by instructions: it
        // visits the header, every entry's key / type / length and the small payloads (JSON, strings, shapes), which
        // lie between the two image bitstreams and behind them.  Pull in the file two ahead — all of it up to 4 KB,
        // else its first and last 2 KB (pack_c2df puts clip_stream and clip_meta last, insertion order).
        if (i + 2 < n) {
          const int64_t p0 = offsets[i + 2], p1 = offsets[i + 3];
          if (p0 >= 0 && p1 > p0) {
            const uint8_t* pb = blob + p0;
            const int64_t len = p1 - p0;
            const int64_t head = len <= 4096 ? len : 2048;
            for (int64_t o = 0; o < head; o += 64) __builtin_prefetch(pb + o, 0, 1);
            if (len > 4096)
              for (int64_t o = (len - 2048) & ~int64_t(63); o < len; o += 64) __builtin_prefetch(pb + o, 0, 1);
          }
        }
        int32_t dd = 0;
        int st;
        const uint8_t* dfr = nullptr;
        uint32_t dlen = 0;
        if (o1 < o0) {
          st = SGIC_C2DF_TRUNCATED;
        } else {
          st = parse_one(z, dctx, blob + o0, static_cast<size_t>(o1 - o0), dim,
                         out_u8 + static_cast<size_t>(i) * dim, &dd, scratch, &memo, frame_off ? &dfr : nullptr, &dlen);
        }
        if (frame_off) {
          frame_off[i] = dfr ? static_cast<int64_t>(dfr - blob) : -1;
          frame_len[i] = dlen;
        }
        status_out[i] = st;
        if (dim_out) dim_out[i] = dd;
      }
    }
    z.freeDCtx(dctx);
  };
  if (n_threads == 1) {
    work();
  } else {
    std::vector<std::thread> th;
    th.reserve(n_threads);
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  return 0;
}

// Second half of the device-decode route: after c2df_parse_batch has classified a slab, lay the good files out
// for the GPU.  Row w of the u8 matrix belongs to the w-th good file; a file decoded on the host has its row
// copied to rows_out[w], a file left to the device has its frame copied (16-byte aligned) into frames_out and
// gets a descriptor.  The prefix sums are serial (a few ns per file); the copies run on `n_threads` threads.
int c2df_pack_batch(const uint8_t* blob, int64_t cnt, int dim, const int32_t* status, const int64_t* frame_off,
                    const uint32_t* frame_len, const uint8_t* rows_in, uint8_t* rows_out, uint8_t* frames_out,
                    size_t frames_cap, ZlDesc* desc_out, int64_t* n_rows_out, int64_t* n_frames_out,
                    int64_t* n_host_rows_out, size_t* frames_bytes_out, int n_threads) {
  std::vector<int64_t> wi(static_cast<size_t>(cnt));      // output row of file i (-1: skipped)
  std::vector<uint32_t> fo(static_cast<size_t>(cnt));     // frame offset (device files)
  std::vector<uint32_t> fk(static_cast<size_t>(cnt));     // descriptor index (device files)
  int64_t w = 0, nf = 0, nh = 0;
  size_t fbytes = 0;
  for (int64_t i = 0; i < cnt; ++i) {
    if (status[i] != SGIC_C2DF_OK) {
      wi[static_cast<size_t>(i)] = -1;
      continue;
    }
    wi[static_cast<size_t>(i)] = w++;
    if (frame_off[i] >= 0) {
      fo[static_cast<size_t>(i)] = static_cast<uint32_t>(fbytes);
      fk[static_cast<size_t>(i)] = static_cast<uint32_t>(nf++);
      fbytes += (static_cast<size_t>(frame_len[i]) + 15u) & ~static_cast<size_t>(15);
    } else {
      ++nh;
    }
  }
  if (fbytes + 16 > frames_cap) {
    set_error("c2df_pack: frame buffer too small");
    return 1;
  }
  *n_rows_out = w;
  *n_frames_out = nf;
  *n_host_rows_out = nh;
  *frames_bytes_out = fbytes;
  if (n_threads <= 0) n_threads = static_cast<int>(std::thread::hardware_concurrency());
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  std::atomic<int64_t> next{0};
  constexpr int64_t kGrain = 1024;
  const size_t d = static_cast<size_t>(dim);
  auto work = [&]() {
    for (;;) {
      const int64_t b = next.fetch_add(kGrain);
      if (b >= cnt) break;
      const int64_t e = (b + kGrain < cnt) ? b + kGrain : cnt;
      for (int64_t i = b; i < e; ++i) {
        if (i + 3 < e && frame_off[i + 3] >= 0) {  // the frames were last touched a whole slab ago: fetch ahead
          const uint8_t* pf = blob + frame_off[i + 3];
          for (uint32_t o = 0; o < frame_len[i + 3]; o += 64) __builtin_prefetch(pf + o, 0, 0);
        }
        const int64_t r = wi[static_cast<size_t>(i)];
        if (r < 0) continue;
        if (frame_off[i] >= 0) {
          std::memcpy(frames_out + fo[static_cast<size_t>(i)], blob + frame_off[i], frame_len[i]);
          desc_out[fk[static_cast<size_t>(i)]] =
              ZlDesc{static_cast<uint32_t>(r), fo[static_cast<size_t>(i)], frame_len[i], 0u};
        } else {
          std::memcpy(rows_out + static_cast<size_t>(r) * d, rows_in + static_cast<size_t>(i) * d, d);
        }
      }
    }
  };
  if (n_threads == 1 || cnt < 4 * kGrain) {
    work();
  } else {
    std::vector<std::thread> th;
    th.reserve(n_threads);
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  return 0;
}

}  // namespace sgic
