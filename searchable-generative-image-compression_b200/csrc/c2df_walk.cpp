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

struct Cursor {
  const uint8_t* p;
  size_t n;
  size_t off = 0;
  bool take(size_t len, const uint8_t** out) {
    if (len > n - off) return false;
    *out = p + off;
    off += len;
    return true;
  }
  bool u8(uint8_t* v) {
    const uint8_t* q;
    if (!take(1, &q)) return false;
    *v = q[0];
    return true;
  }
  bool u16(uint16_t* v) {
    const uint8_t* q;
    if (!take(2, &q)) return false;
    *v = static_cast<uint16_t>(q[0] | (q[1] << 8));
    return true;
  }
  bool u32(uint32_t* v) {
    const uint8_t* q;
    if (!take(4, &q)) return false;
    *v = static_cast<uint32_t>(q[0]) | (static_cast<uint32_t>(q[1]) << 8) | (static_cast<uint32_t>(q[2]) << 16) |
         (static_cast<uint32_t>(q[3]) << 24);
    return true;
  }
};

// ---- a JSON skipper just big enough to read the top-level "dim" of clip_meta ----------
struct Json {
  const char* s;
  size_t n;
  size_t i = 0;
  void ws() {
    while (i < n && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
  }
  bool string(size_t* b, size_t* e) {  // raw span between the quotes
    if (i >= n || s[i] != '"') return false;
    ++i;
    *b = i;
    while (i < n && s[i] != '"') {
      if (s[i] == '\\') ++i;
      ++i;
    }
    if (i >= n) return false;
    *e = i;
    ++i;
    return true;
  }
  // Nesting json.loads accepts before it raises RecursionError (CPython 3.12, the reference's Dockerfile base:
  // the C scanner's recursion budget; measured 9997 with this interpreter).  build.py:87-88 turns that exception
  // into [SKIP]; deeper payloads are refused here the same way instead of being followed.
  static constexpr size_t kMaxDepth = 9990;
  bool key_colon() {  // "key" :
    ws();
    size_t b, e;
    if (!string(&b, &e)) return false;
    ws();
    if (i >= n || s[i] != ':') return false;
    ++i;
    return true;
  }
  // Iterative (the payload is untrusted and may be 4 GB of '['): the open containers live on an explicit stack of
  // one bit each (object / array) — 64 levels inline, deeper ones on the heap.
  bool skip_value() {
    uint64_t small = 0;
    std::vector<uint8_t> big;
    size_t depth = 0;
    auto push = [&](bool obj) {
      if (depth < 64) small = (small & ~(1ull << depth)) | (static_cast<uint64_t>(obj) << depth);
      else {
        if (big.size() <= depth - 64) big.resize(std::max<size_t>(256, 2 * (depth - 64 + 1)));
        big[depth - 64] = obj;
      }
      ++depth;
    };
    auto top_is_obj = [&]() -> bool { return depth <= 64 ? ((small >> (depth - 1)) & 1u) != 0 : big[depth - 1 - 64] != 0; };
    for (;;) {
      ws();
      if (i >= n) return false;
      const char c = s[i];
      if (c == '"') {
        size_t b, e;
        if (!string(&b, &e)) return false;
      } else if (c == '{' || c == '[') {
        if (depth >= kMaxDepth) return false;
        const char close = (c == '{') ? '}' : ']';
        ++i;
        ws();
        if (i < n && s[i] == close) {
          ++i;
        } else {
          push(c == '{');
          if (c == '{' && !key_colon()) return false;
          continue;  // the container's first value
        }
      } else {
        // number / true / false / null / NaN / Infinity
        const size_t b = i;
        while (i < n && s[i] != ',' && s[i] != '}' && s[i] != ']' && s[i] != ' ' && s[i] != '\t' && s[i] != '\n' &&
               s[i] != '\r')
          ++i;
        if (i == b) return false;
      }
      // a value is complete: close finished containers, or step to the next member
      for (;;) {
        if (depth == 0) return true;
        ws();
        if (i >= n) return false;
        if (s[i] == ',') {
          ++i;
          if (top_is_obj() && !key_colon()) return false;
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

// int(meta.get('dim', 0)) for a JSON object payload.  Returns false when Python would have
// raised (not an object, value not convertible); *dim = 0 when the key is absent / meta falsy.
bool clip_meta_dim(const char* s, size_t n, long long* dim) {
  *dim = 0;
  Json j{s, n};
  j.ws();
  if (j.i >= n) return false;
  if (s[j.i] != '{') {
    // `meta = enc_result['clip_meta'] or {}`: null / false / 0 / "" / [] become {}
    const std::string v(s + j.i, n - j.i);
    if (v.rfind("null", 0) == 0 || v.rfind("false", 0) == 0 || v.rfind("[]", 0) == 0 || v.rfind("\"\"", 0) == 0 ||
        v == "0")
      return true;
    return false;  // a list / str / number has no .get -> AttributeError in the reference
  }
  ++j.i;
  j.ws();
  if (j.i < n && s[j.i] == '}') return true;
  for (;;) {
    j.ws();
    size_t kb, ke;
    if (!j.string(&kb, &ke)) return false;
    j.ws();
    if (j.i >= n || s[j.i] != ':') return false;
    ++j.i;
    j.ws();
    const size_t vb = j.i;
    if (!j.skip_value()) return false;
    const size_t ve = j.i;
    if (ke - kb == 3 && std::memcmp(s + kb, "dim", 3) == 0) {
      // what the reference writes is a short run of digits ("dim": 512): no allocation, no strtod for that
      bool digits = ve > vb && ve - vb <= 9 && !(s[vb] == '0' && ve - vb > 1);
      long long iv = 0;
      for (size_t t = vb; digits && t < ve; ++t) {
        digits = s[t] >= '0' && s[t] <= '9';
        iv = iv * 10 + (s[t] - '0');
      }
      if (digits) {
        *dim = iv;
        goto next_member;
      }
      std::string v(s + vb, ve - vb);
      if (v.size() >= 2 && v.front() == '"' && v.back() == '"') {
        // int("512"): Python strips surrounding whitespace and takes an optional sign and decimal digits — no
        // point, no exponent (int("512.0") raises)
        size_t a = 1, b = v.size() - 1;
        auto is_ws = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; };
        while (a < b && is_ws(v[a])) ++a;
        while (b > a && is_ws(v[b - 1])) --b;
        bool neg = false;
        if (a < b && (v[a] == '+' || v[a] == '-')) neg = v[a++] == '-';
        if (a >= b || b - a > 15) return false;
        long long sv = 0;
        for (size_t t = a; t < b; ++t) {
          if (v[t] < '0' || v[t] > '9') return false;
          sv = sv * 10 + (v[t] - '0');
        }
        *dim = neg ? -sv : sv;
        goto next_member;
      }
      if (v == "true") {
        *dim = 1;
      } else if (v == "false") {
        *dim = 0;
      } else {
        char* end = nullptr;
        const double dv = std::strtod(v.c_str(), &end);
        if (end == v.c_str() || *end != '\0' || dv != dv || dv > 9e15 || dv < -9e15) return false;
        *dim = static_cast<long long>(dv);  // int() truncates toward zero
      }
    }
  next_member:
    j.ws();
    if (j.i >= n) return false;
    if (s[j.i] == ',') {
      ++j.i;
      continue;
    }
    if (s[j.i] == '}') return true;
    return false;
  }
}

struct Found {
  const uint8_t* stream = nullptr;
  size_t stream_len = 0;
  bool has_stream = false, stream_is_bytes = false;
  const uint8_t* meta = nullptr;
  size_t meta_len = 0;
  bool has_meta = false;
  uint8_t meta_type = T_NONE;
};

int walk_one(const uint8_t* data, size_t n, Found* f) {
  if (n < 4 || std::memcmp(data, "C2DF", 4) != 0) return SGIC_C2DF_BAD_MAGIC;
  Cursor c{data, n, 4};
  uint16_t ver;
  uint32_t hlen, n_items;
  const uint8_t* skip;
  if (!c.u16(&ver) || !c.u32(&hlen) || !c.take(hlen, &skip) || !c.u32(&n_items)) return SGIC_C2DF_TRUNCATED;
  for (uint32_t it = 0; it < n_items; ++it) {
    uint16_t klen;
    const uint8_t* key;
    uint8_t t;
    if (!c.u16(&klen) || !c.take(klen, &key) || !c.u8(&t)) return SGIC_C2DF_TRUNCATED;
    const uint8_t* payload = nullptr;
    size_t plen = 0;
    if (t == T_INT || t == T_FLOAT) {
      plen = 8;
      if (!c.take(8, &payload)) return SGIC_C2DF_TRUNCATED;
    } else if (t == T_BOOL) {
      plen = 1;
      if (!c.take(1, &payload)) return SGIC_C2DF_TRUNCATED;
    } else if (t == T_NONE) {
      plen = 0;
    } else {
      uint32_t L;
      if (!c.u32(&L) || !c.take(L, &payload)) return SGIC_C2DF_TRUNCATED;
      plen = L;
      if (t > T_BOOL) return SGIC_C2DF_BAD_TYPE;
    }
    const bool is_stream = (klen == 11 && std::memcmp(key, "clip_stream", 11) == 0);
    const bool is_meta = (klen == 9 && std::memcmp(key, "clip_meta", 9) == 0);
    if (!is_stream && !is_meta) continue;
    // BYTES / STR / JSON payloads carry a second u32 length in front of the data
    const uint8_t* inner = payload;
    size_t inner_len = plen;
    if (t == T_BYTES || t == T_STR || t == T_JSON) {
      if (plen < 4) return SGIC_C2DF_TRUNCATED;
      const uint32_t L = static_cast<uint32_t>(payload[0]) | (static_cast<uint32_t>(payload[1]) << 8) |
                         (static_cast<uint32_t>(payload[2]) << 16) | (static_cast<uint32_t>(payload[3]) << 24);
      inner = payload + 4;
      inner_len = (L <= plen - 4) ? L : plen - 4;  // Python slicing clamps
    }
    if (is_stream) {
      f->has_stream = true;
      f->stream_is_bytes = (t == T_BYTES);
      f->stream = inner;
      f->stream_len = inner_len;
    } else {
      f->has_meta = true;
      f->meta_type = t;
      f->meta = inner;
      f->meta_len = inner_len;
    }
  }
  return SGIC_C2DF_OK;
}

// dev_frame != nullptr: frames inside the device decoder's profile (csrc/zstd_lit.cuh) are not decoded here;
// their position is reported instead (*dev_frame = stream start, *dev_len = its length) and the row is filled
// in by zstd_lit_decode_kernel.  libzstd stays the judge of everything else.
int parse_one(const ZstdApi& z, void* dctx, const uint8_t* data, size_t n, int want_dim, uint8_t* out_row,
              int32_t* dim_out, std::vector<uint8_t>& scratch, const uint8_t** dev_frame = nullptr,
              uint32_t* dev_len = nullptr) {
  *dim_out = 0;
  Found f;
  int st = walk_one(data, n, &f);
  if (st != SGIC_C2DF_OK) return st;
  if (!f.has_stream || !f.has_meta) return SGIC_C2DF_NO_CLIP;
  long long dim = 0;
  if (f.meta_type == T_JSON) {
    if (!clip_meta_dim(reinterpret_cast<const char*>(f.meta), f.meta_len, &dim)) return SGIC_C2DF_BAD_DIM;
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
    for (;;) {
      const int64_t b = next.fetch_add(kGrain);
      if (b >= n) break;
      const int64_t e = (b + kGrain < n) ? b + kGrain : n;
      for (int64_t i = b; i < e; ++i) {
        const int64_t o0 = offsets[i], o1 = offsets[i + 1];
        // The walk touches a handful of cache lines per 2 KB file and the blob is far larger than the caches:
        // it is bound by memory latency, not by instructions.  Pull in the lines that are certain to be read of
        // the file two ahead: the container header and the tail, where pack_c2df (insertion order) puts
        // clip_stream and clip_meta.
        if (i + 2 < n) {
          const int64_t p0 = offsets[i + 2], p1 = offsets[i + 3];
          if (p0 >= 0 && p1 > p0) {
            const uint8_t* pb = blob + p0;
            const int64_t len = p1 - p0;
            for (int64_t o = 0; o < 256 && o < len; o += 64) __builtin_prefetch(pb + o, 0, 1);
            for (int64_t o = (len > 640 ? len - 640 : 0) & ~int64_t(63); o < len; o += 64) __builtin_prefetch(pb + o, 0, 1);
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
                         out_u8 + static_cast<size_t>(i) * dim, &dd, scratch, frame_off ? &dfr : nullptr, &dlen);
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
