// K4 — dense Q·Dᵀ on the 5th-gen tensor cores with the top-k fused into the epilogue
// (replaces FAISS exhaustive_inner_product_blas = sgemm blocks + heap/reservoir
// behind src/search.py:115 for batches of queries).
//
// Layout: queries on M (one TMEM lane = one query), database rows on N, d on K.
//   operands  : TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes of 64 elements = 128 B) -> shared memory,
//               completion on mbarriers
//   MMA       : one elected thread issues tcgen05.mma.kind::f16 (K=16 per instruction), fp32 accumulators in
//               TMEM, two accumulator stages (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of
//               tile i+1; tcgen05.commit releases smem stages / publishes accumulators
//   epilogue  : 4 warps; thread t owns TMEM lane t = query t.  tcgen05.ld a chunk of columns, a tree of
//               3-input max + one compare against the query's running k-th score rejects the whole chunk.
//               k <= 32: every thread keeps ITS query's sorted list in shared memory ([rank][query], conflict
//               free) and inserts survivors itself, so 32 queries insert in parallel and the threshold is
//               always fresh.  k > 32: reservoir selection — survivors are appended to the query's reservoir
//               in an L2-resident workspace and the warp compacts a full reservoir to its k best by bisection
//               over the key bits.  No score matrix ever leaves the SM.
// Work split: items = (query tile m, database slice s), m fastest, dealt round-robin to the CTAs (pairs), so the
// CTAs that run concurrently share a few slices and re-use each other's database tiles from L2 while they stay
// close (10M rows: 12.7 GB read from DRAM for a 10.2 GB database; 100M rows: the query tiles drift apart over
// 10.5k tiles and DRAM reads grow to 988 GB).  Shorter slices and a soft lockstep of the producers were measured
// and dropped; what removes the re-reads by construction is scan_dense2b_kernel at the end of this file (the
// database tile stays, the queries stream) — and the measurement that both run at the same speed shows the batch
// regime is limited by the power cap on the tensor pipe, not by DRAM.
// Lists: every (query tile, slice) item is visited once and writes list `slice` of its queries in `partial`
// ([nq][n_slices][k]); merge_keys_*_kernel merges the slices' lists.  Items of one query that run later inherit the
// score bound earlier ones published (gthr).  (The optional short-slice schedule revisits query tiles instead: a
// CTA then parks a tile's list in its own slot and resumes it, threshold included.)
//
// Kernels (plus scan_dense_t_kernel for small batches and scan_dense2b_kernel for very large shards, below):
//   scan_dense_kernel   cta_group::1 — 128 queries x 256 rows per CTA; A and B both streamed per K chunk.
//   scan_dense2_kernel  cta_group::2 — a CTA PAIR (cluster of 2 on one TPC) computes 256 queries x 256 rows:
//                       each CTA holds 128 queries (its TMEM lanes) and loads HALF of the database tile; the
//                       pair's tensor cores read both halves, so shared-memory fill and L2->SM traffic per
//                       flop halve (measured: the 1-CTA kernel is bound by exactly that, profiles/).  With
//                       d <= 512 the query tile stays resident in shared memory for a whole item.
#pragma once
#include <cuda.h>

#include "topk_common.cuh"

namespace sgic {

constexpr int kDenseThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int kDenseBM = 128;       // queries per CTA (TMEM lanes)
constexpr int kDenseBK = 64;        // K chunk: 64 x 16-bit = one 128-byte swizzle row

struct DenseParams {
  uint64_t* partial;   // [nq][n_lists][k] keys (list = slice, or CTA (pair) when lists are parked and resumed)
  uint64_t* lists_ws;  // k > 32: [grid][128][res_cap] reservoirs (L2-resident workspace)
  uint32_t* gthr;      // [nq] by_slice only, zeroed before the launch: per query, the largest "k rows score above
                       // this" bound any finished item has published (order-preserving uint of the score just
                       // below its k-th).  A later item of the same query starts from it instead of -inf — pure
                       // pruning: every published bound is valid, so whichever a racing reader sees, the merged
                       // answer is the same.
  uint32_t n_rows, nq, k, res_cap;
  uint32_t m_tiles;        // query tiles: of 128 (1-CTA kernel) or 256 (2-CTA kernel)
  uint32_t n_slices, tiles_per_slice, n_tiles;
  uint32_t n_lists;        // lists per query in `partial`: the slices (by_slice) or the CTAs (CTA pairs) launched
  uint32_t by_slice;       // 1: every (query tile, slice) item is visited once and writes list `slice` (nothing to
                           //    resume, no zeroing needed); 0: a CTA parks / resumes its lists in its own slot
  uint32_t kc;             // K chunks of 64 elements (ceil(d/64); TMA zero-fills the tail)
  uint32_t tp;             // k <= 32: thread-private lists [k][128] in shared memory; else reservoirs
  uint32_t a_rows;         // rows of the query box (1-CTA kernel; < 128 for small batches)
  uint32_t a_region;       // 1-CTA kernel: bytes reserved per stage for the query box (1 KB granularity)
  uint32_t n_stages;       // 1-CTA kernel: ring depth
  uint32_t idesc;          // UMMA instruction descriptor
  uint32_t db_evict_first; // single query tile: the database is streamed once -> evict_first
  uint32_t debug;          // timing experiments only (wrong results): 1 skip A loads, 2 skip B loads, 4 skip epilogue
  uint32_t epi_wait_ns;    // epilogue warps sleep this long after a failed try on acc_full (0: poll)
};

// shared memory of the epilogue: thread-private lists (k <= 32); the reservoirs of larger k live in L2
__host__ __device__ __forceinline__ uint32_t dense_epi_bytes(uint32_t k, uint32_t tp) {
  return tp ? k * kDenseBM * 8u : 0u;
}

template <int BN>
struct DenseCfg {
  static constexpr uint32_t kABytes = kDenseBM * kDenseBK * 2;
  static constexpr uint32_t kBBytes = BN * kDenseBK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BN;
};

// v[j] for a run-time j in [0,32): binary select tree over the registers (16+8+4+2+1 SEL)
__device__ __forceinline__ float select32(const float (&v)[32], uint32_t j) {
  float a[16], b[8], c[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1u) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2u) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4u) ? b[2 * i + 1] : b[2 * i];
  const float d0 = (j & 8u) ? c[1] : c[0], d1 = (j & 8u) ? c[3] : c[2];
  return (j & 16u) ? d1 : d0;
}

// one of eight registers for a run-time j in [0,8): 4+2+1 SEL
__device__ __forceinline__ float select8(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7,
                                         uint32_t j) {
  const float b0 = (j & 1u) ? a1 : a0, b1 = (j & 1u) ? a3 : a2, b2 = (j & 1u) ? a5 : a4, b3 = (j & 1u) ? a7 : a6;
  const float c0 = (j & 2u) ? b1 : b0, c1 = (j & 2u) ? b3 : b2;
  return (j & 4u) ? c1 : c0;
}

// k <= 32: thread t keeps the sorted list of ITS query at lst[rank * 128] (shared memory; the rank stride of
// 1 KB keeps every lane on its own bank pair whatever rank it touches).  A survivor is inserted by its own
// thread with a short shift loop — 32 queries insert in parallel and `thr` is always the current k-th score.
struct DenseEpiTP {
  uint64_t* lst;
  uint32_t k;
  float thr;
  float thr0;  // bound inherited from finished items of the same query (gthr); -inf if none

  __device__ __forceinline__ void reset(float inherited = -INFINITY) {
    for (uint32_t i = 0; i < k; ++i) lst[i * kDenseBM] = 0ull;
    thr0 = inherited;
    thr = inherited;
  }
  // requires key > current k-th key (the caller compared the score with thr).  Shift loop from the tail,
  // four predecessors loaded per round trip to shared memory (the loop is a latency chain: one warp per
  // scheduler, nothing else to hide it); the new k-th key is tracked in registers.
  __device__ __forceinline__ void insert(uint64_t key) {
    uint32_t i = k - 1;
    uint64_t kth = key;
#pragma unroll 1
    while (i > 0) {
      const uint64_t a0 = lst[(i - 1) * kDenseBM];
      const uint64_t a1 = (i >= 2) ? lst[(i - 2) * kDenseBM] : ~0ull;
      const uint64_t a2 = (i >= 3) ? lst[(i - 3) * kDenseBM] : ~0ull;
      const uint64_t a3 = (i >= 4) ? lst[(i - 4) * kDenseBM] : ~0ull;
      if (a0 > key) break;
      if (i == k - 1) kth = a0;  // the old (k-1)-th key becomes the k-th
      lst[i * kDenseBM] = a0;
      --i;
      if (a1 > key) break;
      lst[i * kDenseBM] = a1;
      --i;
      if (a2 > key) break;
      lst[i * kDenseBM] = a2;
      --i;
      if (a3 > key) break;
      lst[i * kDenseBM] = a3;
      --i;
    }
    lst[i * kDenseBM] = key;
    thr = (kth == 0ull) ? thr0 : key_score(kth);  // every key in the list is above thr0
  }

  template <int BN>
  __device__ __forceinline__ void scan_tile(uint32_t taddr0, uint32_t row0, uint32_t n_valid, bool q_valid) {
#pragma unroll 1
    for (uint32_t c = 0; c < BN / 32; ++c) {
      if (c * 32 >= n_valid) break;
      float v[32];
      ptx::tmem_ld_32x32b_x32(taddr0 + c * 32, v);
      float m4[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float a = fmaxf(fmaxf(v[8 * g], v[8 * g + 1]), v[8 * g + 2]);
        const float b = fmaxf(fmaxf(v[8 * g + 3], v[8 * g + 4]), v[8 * g + 5]);
        m4[g] = fmaxf(fmaxf(a, b), fmaxf(v[8 * g + 6], v[8 * g + 7]));
      }
      const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const bool hit = q_valid && m > thr;
      if (__any_sync(0xffffffffu, hit)) {
        // Rare once the lists have warmed up.  Straight-line survivor mask (2 instructions per column, no
        // branches); then every lane walks ITS OWN survivors — the score comes out of the lane's registers
        // through a 5-level select tree (31 SEL; registers cannot be indexed dynamically) — so the 32 queries
        // of the warp insert in parallel even when their survivors sit in different columns.
        const uint32_t col0 = c * 32;
        uint32_t mask = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) mask |= (v[j] > thr) ? (1u << j) : 0u;
        if (!q_valid) mask = 0;
        if (n_valid - col0 < 32u) mask &= (1u << (n_valid - col0)) - 1u;
        while (mask) {
          const uint32_t j = __ffs(mask) - 1;
          mask &= mask - 1;
          const float sc = select32(v, j);
          if (sc > thr) insert(make_key(sc, row0 + col0 + j));
        }
      }
      __syncwarp();  // tcgen05.ld is warp-collective: reconverge before the next chunk
    }
  }

  // Items are short (a few dozen tiles, so that the query tiles sharing a database slice stay together in L2)
  // and a CTA meets the same query tile again many items later: the list is parked in the CTA's own slot of
  // `partial` between visits and picked up — with its threshold — at the next one.
  __device__ __forceinline__ void load(const uint64_t* partial, uint32_t q, uint32_t nq, uint32_t n_lists, uint32_t slot) {
    if (q >= nq) {
      reset();
      return;
    }
    const uint64_t* src = partial + (static_cast<size_t>(q) * n_lists + slot) * k;
    uint64_t kth = 0ull;
    for (uint32_t i = 0; i < k; ++i) {
      kth = __ldcg(src + i);
      lst[i * kDenseBM] = kth;
    }
    thr0 = -INFINITY;
    thr = (kth == 0ull) ? -INFINITY : key_score(kth);
  }
  __device__ __forceinline__ void store(uint64_t* partial, uint32_t q, uint32_t nq, uint32_t n_lists, uint32_t slot) {
    if (q >= nq) return;
    uint64_t* dst = partial + (static_cast<size_t>(q) * n_lists + slot) * k;
    for (uint32_t i = 0; i < k; ++i) __stcg(dst + i, lst[i * kDenseBM]);
  }
};

// k > 32: reservoir selection (what FAISS does for k >= 100, ReservoirTopN behind exhaustive_inner_product_blas).
// Thread t appends every score above its threshold to ITS query's reservoir of C = max(256, 2*kp) keys in an
// L2-resident workspace — one fire-and-forget store, nothing to wait for.  When a reservoir is full the warp
// compacts it together: the k-th largest key is found by bisection over the key bits (count(key >= cand) with a
// REDUX per step; C = 256 keeps the keys in registers, 8 per lane), the k survivors move to the front, and the
// query's threshold becomes the k-th score.  A stale threshold only lets extra candidates in, never drops one,
// so the k keys left at the end of an item are exactly its top-k (unsorted; the merge kernel sorts).
struct DenseEpiRes {
  uint64_t* res_warp;  // reservoir of this warp's lane 0; lane L's starts at res_warp + L * C
  uint32_t k, C;
  int lane;
  uint32_t cnt;
  float thr;
  uint32_t* gthr_q = nullptr;  // this lane's query in DenseParams::gthr (nullptr: no sharing)

  // after a compaction: publish the score just below the new k-th and pick up a better bound if another list of
  // the same query (a concurrent slice, or an item of the previous round) has one
  __device__ __forceinline__ void exchange_bound() {
    if (gthr_q == nullptr || !(thr > -INFINITY)) return;
    const uint32_t mine = score_to_ord(thr) - 1u;
    const uint32_t old = atomicMax(gthr_q, mine);
    if (old > mine) thr = ord_to_score(old);
  }

  __device__ __forceinline__ void reset() {
    cnt = 0;
    thr = -INFINITY;
  }
  __device__ __forceinline__ void append(uint64_t key) {
    __stcg(res_warp + static_cast<size_t>(lane) * C + cnt, key);
    ++cnt;
  }

  // Warp-cooperative: keep the k largest of R[0..n) (n > k) in R[0..k); returns the k-th largest key.
  // KPLR > 0: the reservoir (n <= 32 * KPLR keys) is held in registers, KPLR keys per lane; 0: re-read from L2.
  template <int KPLR>
  __device__ __forceinline__ uint64_t select_topk(uint64_t* R, uint32_t n) {
    constexpr bool REG = KPLR > 0;
    constexpr int KPL = REG ? KPLR : 1;
    uint64_t key[KPL];
    if (REG) {
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        const uint32_t idx = i * 32 + lane;
        key[i] = idx < n ? __ldcg(R + idx) : 0ull;  // 0 sorts below every real key
      }
    }
    // bits above the highest bit in which two keys differ are common to all keys: start the bisection there
    const uint64_t k0 = __ldcg(R);
    uint32_t xh = 0, xl = 0;
    if (REG) {
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        const uint64_t x = key[i] ? key[i] ^ k0 : 0ull;
        xh |= static_cast<uint32_t>(x >> 32);
        xl |= static_cast<uint32_t>(x);
      }
    } else {
      for (uint32_t idx = lane; idx < n; idx += 32) {
        const uint64_t x = __ldcg(R + idx) ^ k0;
        xh |= static_cast<uint32_t>(x >> 32);
        xl |= static_cast<uint32_t>(x);
      }
    }
    xh = __reduce_or_sync(0xffffffffu, xh);
    xl = __reduce_or_sync(0xffffffffu, xl);
    int b = xh ? 63 - __clz(xh) : 31 - __clz(xl);  // keys are unique and n >= 2: some bit differs
    uint64_t prefix = (b == 63) ? 0ull : (k0 >> (b + 1)) << (b + 1);
    for (; b >= 0; --b) {
      const uint64_t cand = prefix | (1ull << b);
      uint32_t c = 0;
      if (REG) {
#pragma unroll
        for (int i = 0; i < KPL; ++i) c += key[i] >= cand ? 1u : 0u;
      } else {
        for (uint32_t idx = lane; idx < n; idx += 32) c += __ldcg(R + idx) >= cand ? 1u : 0u;
      }
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= k) prefix = cand;
      if (c == k) break;  // exactly k keys at or above cand: they are the top-k
    }
    // k-th key = the smallest key >= prefix
    uint64_t kth = ~0ull;
    if (REG) {
#pragma unroll
      for (int i = 0; i < KPL; ++i)
        if (key[i] >= prefix && key[i] < kth) kth = key[i];
    } else {
      for (uint32_t idx = lane; idx < n; idx += 32) {
        const uint64_t x = __ldcg(R + idx);
        if (x >= prefix && x < kth) kth = x;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const uint64_t o = __shfl_xor_sync(0xffffffffu, kth, off);
      kth = o < kth ? o : kth;
    }
    // survivors to the front (stable, in place: the write position never passes the read position)
    uint32_t out = 0;
    const uint32_t lt = (1u << lane) - 1u;
    if (REG) {
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        const bool keep = key[i] >= kth;
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (keep) __stcg(R + out + __popc(m & lt), key[i]);
        out += __popc(m);
      }
    } else {
      for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t idx = base + lane;
        const uint64_t x = idx < n ? __ldcg(R + idx) : 0ull;
        const bool keep = x >= kth;
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) __stcg(R + out + __popc(m & lt), x);
        out += __popc(m);
        __syncwarp();
      }
    }
    __syncwarp();
    return kth;
  }

  // Large reservoirs (C = 1024 / 2048, 256 < k <= 1024): 64-bit keys do not fit in registers, but their SCORE words
  // do (KH per lane).  The k-th largest score word T is found by bisection in registers; only the entries that tie
  // with T — normally one — are looked at again in memory to settle the id word.  Three passes over the reservoir
  // (load, tie, compaction) instead of one per bisection step.
  template <int KH>
  __device__ __forceinline__ uint64_t select_topk_hi(uint64_t* R, uint32_t n) {
    uint32_t hi[KH];
#pragma unroll
    for (int i = 0; i < KH; ++i) {
      const uint32_t idx = i * 32 + lane;
      hi[i] = idx < n ? static_cast<uint32_t>(__ldcg(R + idx) >> 32) : 0u;  // 0: empty (no real score maps to 0)
    }
    const uint32_t h0 = __shfl_sync(0xffffffffu, hi[0], 0);
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < KH; ++i) x |= hi[i] ? hi[i] ^ h0 : 0u;
    x = __reduce_or_sync(0xffffffffu, x);
    uint32_t T = h0;  // all score words equal: T is that word
    if (x) {
      int b = 31 - __clz(x);
      T = (b == 31) ? 0u : (h0 >> (b + 1)) << (b + 1);
      for (; b >= 0; --b) {  // T = largest value with count(hi >= T) >= k = the k-th largest score word
        const uint32_t cand = T | (1u << b);
        uint32_t c = 0;
#pragma unroll
        for (int i = 0; i < KH; ++i) c += hi[i] >= cand ? 1u : 0u;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= k) T = cand;
      }
    }
    uint32_t c_gt = 0, c_eq = 0;
#pragma unroll
    for (int i = 0; i < KH; ++i) {
      c_gt += hi[i] > T ? 1u : 0u;
      c_eq += hi[i] == T ? 1u : 0u;
    }
    c_gt = __reduce_add_sync(0xffffffffu, c_gt);
    c_eq = __reduce_add_sync(0xffffffffu, c_eq);
    const uint32_t r = k - c_gt;  // how many of the entries that tie with T belong to the top-k (1 <= r <= c_eq)
    // id word of the k-th key = the r-th largest id word among the ties
    uint32_t lo_kth = 0;
    if (r == c_eq) {  // all ties are in: the smallest id word among them
      uint32_t mn = 0xffffffffu;
#pragma unroll
      for (int i = 0; i < KH; ++i)
        if (hi[i] == T) mn = min(mn, static_cast<uint32_t>(__ldcg(R + i * 32 + lane)));
      lo_kth = __reduce_min_sync(0xffffffffu, mn);
    } else {  // exact score ties straddle the k-th place (duplicate rows): bisection over the ties' id words
      for (int b = 31; b >= 0; --b) {
        const uint32_t cand = lo_kth | (1u << b);
        uint32_t c = 0;
#pragma unroll
        for (int i = 0; i < KH; ++i)
          if (hi[i] == T) c += static_cast<uint32_t>(__ldcg(R + i * 32 + lane)) >= cand ? 1u : 0u;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= r) lo_kth = cand;
      }
    }
    const uint64_t kth = (static_cast<uint64_t>(T) << 32) | lo_kth;
    // survivors to the front (stable, in place: the write position never passes the read position)
    uint32_t out = 0;
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t base = 0; base < n; base += 32) {
      const uint32_t idx = base + lane;
      const uint64_t v = idx < n ? __ldcg(R + idx) : 0ull;
      const bool keep = v >= kth;
      const uint32_t m = __ballot_sync(0xffffffffu, keep);
      __syncwarp();
      if (keep) __stcg(R + out + __popc(m & lt), v);
      out += __popc(m);
      __syncwarp();
    }
    __syncwarp();
    return kth;
  }

  // Small reservoirs (C = 32 * KPL keys, KPL keys per lane): rank sort in registers.  Every lane counts the keys
  // larger than its own (32 * KPL shuffles of 64 bits); a key of rank r < k goes to slot r, so the k survivors come
  // out SORTED and the k-th key is the one of rank k-1.  The next reservoir's keys are fetched from L2 while this
  // one is ranked (the startup of a scan compacts all 32 lanes' reservoirs back to back).
  template <int KPL>
  __device__ __forceinline__ void fetch_small(const uint64_t* R, uint32_t n, uint64_t (&key)[KPL]) {
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const uint32_t idx = j * 32 + lane;
      key[j] = idx < n ? __ldcg(R + idx) : 0ull;  // 0 = empty: below every real key, never stored
    }
  }
  template <int KPL>
  __device__ __forceinline__ void compact_small(uint32_t need) {
    __syncwarp();
    int L = __ffs(need) - 1;
    uint32_t n = __shfl_sync(0xffffffffu, cnt, L);
    uint64_t key[KPL];
    fetch_small<KPL>(res_warp + static_cast<size_t>(L) * C, n, key);
    while (true) {
      need &= need - 1;
      int Ln = 0;
      uint32_t nn = 0;
      uint64_t nk[KPL];
      if (need) {
        Ln = __ffs(need) - 1;
        nn = __shfl_sync(0xffffffffu, cnt, Ln);
        fetch_small<KPL>(res_warp + static_cast<size_t>(Ln) * C, nn, nk);
      }
      uint32_t rank[KPL];
#pragma unroll
      for (int j = 0; j < KPL; ++j) rank[j] = 0;
#pragma unroll 4
      for (int src = 0; src < 32; ++src) {
#pragma unroll
        for (int jj = 0; jj < KPL; ++jj) {
          const uint64_t o = __shfl_sync(0xffffffffu, key[jj], src);
#pragma unroll
          for (int j = 0; j < KPL; ++j) rank[j] += o > key[j] ? 1u : 0u;
        }
      }
      uint64_t* R = res_warp + static_cast<size_t>(L) * C;
      uint64_t kth = 0ull;
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        const bool real = key[j] != 0ull;
        if (real && rank[j] < k) __stcg(R + rank[j], key[j]);
        const uint32_t m = __ballot_sync(0xffffffffu, real && rank[j] == k - 1);
        if (m) kth = __shfl_sync(0xffffffffu, key[j], __ffs(m) - 1);
      }
      if (lane == L) {
        cnt = n < k ? n : k;
        if (kth != 0ull) {
          thr = key_score(kth);
          exchange_bound();
        }
      }
      if (!need) break;
      L = Ln;
      n = nn;
#pragma unroll
      for (int j = 0; j < KPL; ++j) key[j] = nk[j];
    }
    __syncwarp();
  }

  // compact the reservoirs of the lanes named in `need` (warp-uniform mask)
  __device__ __forceinline__ void compact(uint32_t need) {
    if (C == 32u) return compact_small<1>(need);
    if (C == 64u) return compact_small<2>(need);
    __syncwarp();
    while (need) {
      const int L = __ffs(need) - 1;
      need &= need - 1;
      const uint32_t n = __shfl_sync(0xffffffffu, cnt, L);
      uint64_t* R = res_warp + static_cast<size_t>(L) * C;
      const uint64_t kth = (C == 256u)    ? select_topk<8>(R, n)
                           : (C == 512u)  ? select_topk<16>(R, n)
                           : (C == 1024u) ? select_topk_hi<32>(R, n)
                           : (C == 2048u) ? select_topk_hi<64>(R, n)
                                          : select_topk<0>(R, n);
      if (lane == L) {
        cnt = k;
        thr = key_score(kth);
        exchange_bound();
      }
    }
  }

  template <int BN>
  __device__ __forceinline__ void scan_tile(uint32_t taddr0, uint32_t row0, uint32_t n_valid, bool q_valid) {
#pragma unroll 1
    for (uint32_t c = 0; c < BN / 32; ++c) {
      if (c * 32 >= n_valid) break;
      float v[32];
      ptx::tmem_ld_32x32b_x32(taddr0 + c * 32, v);
      float m4[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float a = fmaxf(fmaxf(v[8 * g], v[8 * g + 1]), v[8 * g + 2]);
        const float b = fmaxf(fmaxf(v[8 * g + 3], v[8 * g + 4]), v[8 * g + 5]);
        m4[g] = fmaxf(fmaxf(a, b), fmaxf(v[8 * g + 6], v[8 * g + 7]));
      }
      const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const bool hit = q_valid && m > thr;
      if (__any_sync(0xffffffffu, hit)) {
        // Large k keeps this path busy (k = 100: ~1000 appends per list), so it is cut to the bone: the four
        // group maxima are already there — one vote per group of 8 columns, a survivor mask only for a group
        // that has one, and the score comes out of 8 registers (7 SEL) because the group is a compile-time index.
        const uint32_t col0 = c * 32;
        const uint32_t lim = n_valid - col0;  // valid columns in this chunk (>= 1)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (!__any_sync(0xffffffffu, q_valid && m4[g] > thr)) continue;
          uint32_t mask = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) mask |= (v[8 * g + j] > thr) ? (1u << j) : 0u;
          if (!q_valid) mask = 0;
          if (lim < 8u * g + 8u) mask &= (lim > 8u * g) ? ((1u << (lim - 8u * g)) - 1u) : 0u;
          // warp-uniform loop: every lane appends its next survivor, full reservoirs are compacted at once
          while (__any_sync(0xffffffffu, mask != 0u)) {
            if (mask) {
              const uint32_t j = __ffs(mask) - 1;
              mask &= mask - 1;
              const float sc = select8(v[8 * g], v[8 * g + 1], v[8 * g + 2], v[8 * g + 3], v[8 * g + 4], v[8 * g + 5],
                                       v[8 * g + 6], v[8 * g + 7], j);
              if (sc > thr) append(make_key(sc, row0 + col0 + 8u * g + j));
            }
            const uint32_t full = __ballot_sync(0xffffffffu, cnt == C);
            if (full) compact(full);
          }
        }
      }
      __syncwarp();
    }
  }

  // item start: pick up the (unsorted, front-compacted) top-k parked in this CTA's slot of `partial` by its
  // previous visit of the query tile; the threshold is the smallest of k keys, or -inf while fewer are known
  __device__ __forceinline__ void load(const uint64_t* partial, uint32_t q0w, uint32_t nq, uint32_t n_lists,
                                       uint32_t slot) {
    reset();
    for (int L = 0; L < 32; ++L) {
      const uint32_t q = q0w + L;
      if (q >= nq) break;
      const uint64_t* src = partial + (static_cast<size_t>(q) * n_lists + slot) * k;
      uint64_t* R = res_warp + static_cast<size_t>(L) * C;
      uint32_t n = 0;
      uint64_t mn = ~0ull;
      for (uint32_t i = lane; i < k; i += 32) {
        const uint64_t x = __ldcg(src + i);
        __stcg(R + i, x);
        if (x != 0ull) {
          ++n;
          mn = x < mn ? x : mn;
        }
      }
      n = __reduce_add_sync(0xffffffffu, n);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const uint64_t o = __shfl_xor_sync(0xffffffffu, mn, off);
        mn = o < mn ? o : mn;
      }
      if (lane == L) {
        cnt = n;
        thr = (n == k) ? key_score(mn) : -INFINITY;
      }
    }
    __syncwarp();
  }

  // item finished: reservoirs down to k keys, then the warp copies each query's keys out (coalesced)
  __device__ __forceinline__ void store(uint64_t* partial, uint32_t q0w, uint32_t nq, uint32_t n_slices,
                                        uint32_t slice) {
    const uint32_t over = __ballot_sync(0xffffffffu, cnt > k);
    if (over) compact(over);
    __syncwarp();
    for (int L = 0; L < 32; ++L) {
      const uint32_t q = q0w + L;
      if (q >= nq) break;
      const uint32_t n = __shfl_sync(0xffffffffu, cnt, L);
      const uint64_t* R = res_warp + static_cast<size_t>(L) * C;
      uint64_t* dst = partial + (static_cast<size_t>(q) * n_slices + slice) * k;
      for (uint32_t i = lane; i < k; i += 32) __stcg(dst + i, i < n ? __ldcg(R + i) : 0ull);
    }
    __syncwarp();
  }
};

// gthr protocol (see DenseParams): read the bound a query inherited / publish the score just below a list's k-th
__device__ __forceinline__ float gthr_read(const uint32_t* gthr, uint32_t q) {
  const uint32_t g = __ldcg(gthr + q);
  return g ? ord_to_score(g) : -INFINITY;
}
__device__ __forceinline__ void gthr_publish(uint32_t* gthr, uint32_t q, float thr) {
  // ord - 1 is the next float below thr: rows that TIE with the k-th score stay admissible everywhere, so ties at
  // the boundary are still resolved by row number in the merge
  if (thr > -INFINITY) atomicMax(gthr + q, score_to_ord(thr) - 1u);
}

// The epilogue warps' loop over this CTA's items; `arrive(as)` hands accumulator stage `as` back to the
// MMA issuer.  q_off: first query of this CTA inside the item's query tile (pairs: rank * 128).
template <int BN, typename Arrive>
__device__ __forceinline__ void dense_epilogue(const DenseParams& p, uint32_t unit, uint32_t n_units, uint32_t q_tile,
                                               uint32_t q_off, uint32_t tmem_base, uint8_t* epi_smem,
                                               uint64_t* acc_full, int warp, int lane, Arrive arrive) {
  const uint32_t n_items = p.m_tiles * p.n_slices;
  const int row0w = (warp & 3) * 32;  // TMEM lane quadrant this warp may read = warp_id % 4
  const int t = row0w + lane;
  const uint32_t tq = tmem_base + (static_cast<uint32_t>(row0w) << 16);
  uint32_t tc = 0;
  if (p.tp) {
    DenseEpiTP epi;
    epi.lst = reinterpret_cast<uint64_t*>(epi_smem) + t;
    epi.k = p.k;
    for (uint32_t item = unit; item < n_items; item += n_units) {
      const uint32_t m_tile = item % p.m_tiles, slice = item / p.m_tiles;
      const uint32_t t0 = slice * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
      const uint32_t q0 = m_tile * q_tile + q_off;
      const bool q_valid = q0 + t < p.nq;
      const bool warp_valid = q0 + row0w < p.nq;
      const uint32_t slot = p.by_slice ? slice : unit;
      if (p.by_slice) epi.reset((p.gthr && q_valid) ? gthr_read(p.gthr, q0 + t) : -INFINITY);
      else epi.load(p.partial, q0 + t, p.nq, p.n_lists, slot);
      for (uint32_t tile = t0; tile < t1; ++tile, ++tc) {
        const uint32_t as = tc & 1;
        ptx::mbar_wait_relaxed(&acc_full[as], (tc >> 1) & 1, p.epi_wait_ns);
        ptx::tc_fence_after();
        const uint32_t row0 = tile * BN;
        const uint32_t n_valid = min(static_cast<uint32_t>(BN), p.n_rows - row0);
        if (warp_valid && !(p.debug & 4u)) epi.template scan_tile<BN>(tq + as * BN, row0, n_valid, q_valid);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(as);
      }
      epi.store(p.partial, q0 + t, p.nq, p.n_lists, slot);
      if (p.by_slice && p.gthr && q_valid) gthr_publish(p.gthr, q0 + t, epi.thr);
    }
  } else {
    DenseEpiRes epi;
    epi.k = p.k;
    epi.C = p.res_cap;
    epi.lane = lane;
    epi.res_warp = p.lists_ws + (static_cast<size_t>(blockIdx.x) * kDenseBM + row0w) * p.res_cap;
    for (uint32_t item = unit; item < n_items; item += n_units) {
      const uint32_t m_tile = item % p.m_tiles, slice = item / p.m_tiles;
      const uint32_t t0 = slice * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
      const uint32_t q0 = m_tile * q_tile + q_off;
      const bool q_valid = q0 + t < p.nq;
      const bool warp_valid = q0 + row0w < p.nq;
      const uint32_t slot = p.by_slice ? slice : unit;
      if (warp_valid && !p.by_slice) epi.load(p.partial, q0 + row0w, p.nq, p.n_lists, slot);
      else epi.reset();
      epi.gthr_q = (p.by_slice && p.gthr && q_valid) ? p.gthr + q0 + t : nullptr;
      if (epi.gthr_q) epi.thr = gthr_read(p.gthr, q0 + t);
      for (uint32_t tile = t0; tile < t1; ++tile, ++tc) {
        const uint32_t as = tc & 1;
        // Compact BEFORE taking the next accumulator, while this warp would only wait for it: a reservoir with
        // fewer than 64 free slots is brought down to its k best now, so that (almost) no compaction happens while
        // the warp holds a TMEM stage the MMA issuer is waiting to get back (ncu: the issuer spent 12 % of its time
        // on acc_empty at k = 100).  A lane that still overflows inside a tile compacts there, as before.
        if (warp_valid && !(p.debug & 8u)) {
          const uint32_t soon = __ballot_sync(0xffffffffu, epi.cnt + 64u > epi.C && epi.cnt > epi.k);
          if (soon) epi.compact(soon);
        }
        ptx::mbar_wait_relaxed(&acc_full[as], (tc >> 1) & 1, p.epi_wait_ns);
        ptx::tc_fence_after();
        const uint32_t row0 = tile * BN;
        const uint32_t n_valid = min(static_cast<uint32_t>(BN), p.n_rows - row0);
        if (warp_valid && !(p.debug & 4u)) epi.template scan_tile<BN>(tq + as * BN, row0, n_valid, q_valid);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(as);
      }
      if (warp_valid) epi.store(p.partial, q0 + row0w, p.nq, p.n_lists, slot);
      if (p.by_slice && p.gthr && q_valid) gthr_publish(p.gthr, q0 + t, epi.thr);
    }
  }
}

// ================================================================================================ 1-CTA kernel
constexpr int kDenseMaxStages = 8;

// Ring of p.n_stages stages; a stage = the query box of this K chunk (p.a_region bytes, 1 KB granularity — the box
// is cut to the batch, so small batches get a deeper ring of database tiles) + a 256-row database box.
template <int BN>
__global__ void __launch_bounds__(kDenseThreads, 1)
scan_dense_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_db,
                  const DenseParams p) {
  using Cfg = DenseCfg<BN>;
  extern __shared__ uint8_t dense_smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* smem = dense_smem_raw + ((1024u - (ptx::smem_u32(dense_smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;
  const uint32_t ns = p.n_stages, stage_bytes = p.a_region + Cfg::kBBytes;
  uint8_t* epi_smem = smem + static_cast<size_t>(ns) * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + dense_epi_bytes(p.k, p.tp));
  uint64_t* full = bars;
  uint64_t* empty = bars + kDenseMaxStages;
  uint64_t* acc_full = bars + 2 * kDenseMaxStages;
  uint64_t* acc_empty = bars + 2 * kDenseMaxStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kDenseMaxStages + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_db);
    for (uint32_t s = 0; s < ns; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&acc_full[a], 1);
      ptx::mbar_init(&acc_empty[a], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t n_items = p.m_tiles * p.n_slices;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      const uint64_t pol_q = ptx::policy_evict_last();
      const uint64_t pol_db = p.db_evict_first ? ptx::policy_evict_first() : ptx::policy_evict_last();
      uint32_t s = 0, round = 0;  // ring position; `round` counts completed trips around the ring
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t m_tile = item % p.m_tiles, slice = item / p.m_tiles;
        const uint32_t t0 = slice * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
        for (uint32_t tile = t0; tile < t1; ++tile) {
          for (uint32_t kc = 0; kc < p.kc; ++kc) {
            if (round > 0) ptx::mbar_wait(&empty[s], (round - 1) & 1);
            uint8_t* a_dst = stages + static_cast<size_t>(s) * stage_bytes;
            const bool ld_a = !(p.debug & 1u) || round == 0, ld_b = !(p.debug & 2u) || round == 0;
            ptx::mbar_expect_tx(&full[s], (ld_a ? p.a_rows * (kDenseBK * 2u) : 0u) + (ld_b ? Cfg::kBBytes : 0u));
            if (ld_a)
              ptx::tma_load_2d(a_dst, &tm_q, static_cast<int32_t>(kc * kDenseBK),
                               static_cast<int32_t>(m_tile * kDenseBM), &full[s], pol_q);
            if (ld_b)
              ptx::tma_load_2d(a_dst + p.a_region, &tm_db, static_cast<int32_t>(kc * kDenseBK),
                               static_cast<int32_t>(tile * BN), &full[s], pol_db);
            if (++s == ns) {
              s = 0;
              ++round;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    // whole warp, warp-uniform operands, one elected lane issues (see ptx::elect_one)
    {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t stages_u = ptx::smem_u32(stages);
      const uint32_t idesc = p.idesc, kcs = p.kc, a_region = p.a_region;
      uint32_t s = 0, round = 0, tc = 0;
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t slice = item / p.m_tiles;
        const uint32_t t0 = slice * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
        for (uint32_t tile = t0; tile < t1; ++tile, ++tc) {
          const uint32_t as = tc & 1, ua = tc >> 1;
          if (ua > 0) ptx::mbar_wait(&acc_empty[as], (ua - 1) & 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_u + as * BN;
          for (uint32_t kc = 0; kc < kcs; ++kc) {
            ptx::mbar_wait(&full[s], round & 1);
            ptx::tc_fence_after();
            const uint32_t a_addr = stages_u + s * stage_bytes;
            const uint64_t a_desc = ptx::umma_desc_k_sw128(a_addr);
            const uint64_t b_desc = ptx::umma_desc_k_sw128(a_addr + a_region);
            if (ptx::elect_one()) {
#pragma unroll
              for (uint32_t kk = 0; kk < kDenseBK / 16; ++kk)
                ptx::tc_mma_f16(d_tmem, a_desc + kk * 2, b_desc + kk * 2, idesc, (kc | kk) != 0 ? 1u : 0u);
              ptx::tc_commit(&empty[s]);  // smem stage reusable once these MMAs have read it
            }
            __syncwarp();
            if (++s == ns) {
              s = 0;
              ++round;
            }
          }
          if (ptx::elect_one()) ptx::tc_commit(&acc_full[as]);  // accumulator of this tile complete
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================================ epilogue: fused top-k
    dense_epilogue<BN>(p, blockIdx.x, gridDim.x, kDenseBM, 0u, tmem_base, epi_smem, acc_full, warp, lane,
                       [&](uint32_t as) { ptx::mbar_arrive(&acc_empty[as]); });
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ================================================================================================ transposed kernel
// Small batches (HBM regime).  With the queries on M the tensor cores always multiply a 128-row query tile, so a
// batch of 8 burns 16x the useful tensor energy; measured on a 1000 W B200 that drives the SM clock to 637 MHz
// under the power cap and the scan to 5.5 TB/s, while K3 streams the same rows at 7.3 TB/s.  Here the DATABASE
// rows sit on M (one TMEM lane = one database row, M = 128) and the queries on N = ceil16(nq), so the tensor work
// is proportional to the batch.  The query matrix is loaded once and stays in shared memory; every stage is one
// 256-row x 64-element database box (two M = 128 MMAs per K step, two accumulators of N columns).
//
// Epilogue: warp w owns lanes 32w.. (32 database rows per sub-tile), thread = row, column = query.  Per chunk of
// 16 queries: one tcgen05.ld.x16, the 16 per-query thresholds from shared memory (4 broadcast LDS.128), 16
// compares OR-ed into one predicate and ONE vote rejects 32 rows x 16 queries.  Survivors (rare) are found with
// one ballot per query and inserted warp-cooperatively into the (warp, query) list; the four warps' lists of a
// query are merged when the CTA's slice is done.
struct DenseTParams {
  uint64_t* partial;  // [nq][n_lists][k] keys
  uint32_t n_rows, nq, k;
  uint32_t n_pad;     // ceil16(nq): query columns the epilogue walks (lists, thresholds, accumulator stride)
  uint32_t n_mma;     // MMA N = rows of the query box: n_pad, or 8 for batches of up to 8 queries (the columns
                      // 8..15 of an accumulator are then never written; their thresholds are +inf)
  uint32_t n_slices, tiles_per_slice, n_tiles, n_lists;
  uint32_t kc, n_stages, idesc, db_evict_first, debug;
  uint32_t epi_wait_ns;  // epilogue warps sleep this long after a failed try on acc_full (0: poll)
};

constexpr uint32_t kDtStageBytes = 256u * kDenseBK * 2u;  // 32 KB: 256 database rows x 128 B

__host__ __device__ __forceinline__ uint32_t dense_t_fixed_bytes(uint32_t n_pad, uint32_t kc, uint32_t k) {
  // resident query chunks + (query, warp) lists + thresholds
  return kc * n_pad * (kDenseBK * 2u) + n_pad * 4u * k * 8u + 4u * n_pad * 4u;
}

__global__ void __launch_bounds__(kDenseThreads, 1)
scan_dense_t_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_db,
                    const DenseTParams p) {
  extern __shared__ uint8_t dense_smem_raw[];
  uint8_t* smem = dense_smem_raw + ((1024u - (ptx::smem_u32(dense_smem_raw) & 1023u)) & 1023u);
  const uint32_t ns = p.n_stages, q_chunk = p.n_mma * (kDenseBK * 2u);
  uint8_t* stages = smem;
  uint8_t* q_res = smem + static_cast<size_t>(ns) * kDtStageBytes;  // kc chunks of [n_pad][128 B], SWIZZLE_128B
  uint64_t* lists = reinterpret_cast<uint64_t*>(q_res + static_cast<size_t>(p.kc) * q_chunk);  // [n_pad][4][k]
  float* thr = reinterpret_cast<float*>(lists + static_cast<size_t>(p.n_pad) * 4 * p.k);      // [4][n_pad]
  uint64_t* bars = reinterpret_cast<uint64_t*>(thr + 4 * p.n_pad);
  uint64_t* full = bars;
  uint64_t* empty = bars + kDenseMaxStages;
  uint64_t* acc_full = bars + 2 * kDenseMaxStages;
  uint64_t* acc_empty = bars + 2 * kDenseMaxStages + 2;
  uint64_t* q_full = bars + 2 * kDenseMaxStages + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kDenseMaxStages + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_db);
    for (uint32_t s = 0; s < ns; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&acc_full[a], 1);
      ptx::mbar_init(&acc_empty[a], 4);
    }
    ptx::mbar_init(q_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t n_items = p.n_slices;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      const uint64_t pol_q = ptx::policy_evict_last();
      const uint64_t pol_db = p.db_evict_first ? ptx::policy_evict_first() : ptx::policy_evict_last();
      ptx::mbar_expect_tx(q_full, p.kc * q_chunk);
      for (uint32_t kc = 0; kc < p.kc; ++kc)
        ptx::tma_load_2d(q_res + static_cast<size_t>(kc) * q_chunk, &tm_q, static_cast<int32_t>(kc * kDenseBK), 0, q_full,
                         pol_q);
      uint32_t s = 0, round = 0;
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t t0 = item * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
        for (uint32_t tile = t0; tile < t1; ++tile) {
          for (uint32_t kc = 0; kc < p.kc; ++kc) {
            if (round > 0) ptx::mbar_wait(&empty[s], (round - 1) & 1);
            ptx::mbar_expect_tx(&full[s], kDtStageBytes);
            ptx::tma_load_2d(stages + static_cast<size_t>(s) * kDtStageBytes, &tm_db, static_cast<int32_t>(kc * kDenseBK),
                             static_cast<int32_t>(tile * 256u), &full[s], pol_db);
            if (++s == ns) {
              s = 0;
              ++round;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    // whole warp, warp-uniform operands, one elected lane issues (see ptx::elect_one)
    {
      ptx::mbar_wait(q_full, 0);
      ptx::tc_fence_after();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t stages_u = ptx::smem_u32(stages), q_res_u = ptx::smem_u32(q_res);
      const uint32_t idesc = p.idesc, kcs = p.kc, n_pad = p.n_pad;
      uint32_t s = 0, round = 0, tc = 0;
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t t0 = item * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
        for (uint32_t tile = t0; tile < t1; ++tile, ++tc) {
          const uint32_t as = tc & 1, ua = tc >> 1;
          if (ua > 0) ptx::mbar_wait(&acc_empty[as], (ua - 1) & 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_u + as * 256u;
          for (uint32_t kc = 0; kc < kcs; ++kc) {
            ptx::mbar_wait(&full[s], round & 1);
            ptx::tc_fence_after();
            const uint32_t db_addr = stages_u + s * kDtStageBytes;
            const uint64_t q_desc = ptx::umma_desc_k_sw128(q_res_u + kc * q_chunk);
            if (ptx::elect_one()) {
#pragma unroll
              for (uint32_t sub = 0; sub < 2; ++sub) {
                const uint64_t db_desc = ptx::umma_desc_k_sw128(db_addr + sub * (128u * kDenseBK * 2u));
#pragma unroll
                for (uint32_t kk = 0; kk < kDenseBK / 16; ++kk)
                  ptx::tc_mma_f16(d_tmem + sub * n_pad, db_desc + kk * 2, q_desc + kk * 2, idesc,
                                  (kc | kk) != 0 ? 1u : 0u);
              }
              ptx::tc_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == ns) {
              s = 0;
              ++round;
            }
          }
          if (ptx::elect_one()) ptx::tc_commit(&acc_full[as]);
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================================ epilogue
    const uint32_t w = warp & 3;  // TMEM lane quadrant of this warp
    float* thr_w = thr + w * p.n_pad;
    const uint32_t k = p.k;
    for (uint32_t q = lane; q < p.n_pad; q += 32) thr_w[q] = q < p.nq ? -INFINITY : INFINITY;
    for (uint32_t q = 0; q < p.n_pad; ++q)
      for (uint32_t i = lane; i < k; i += 32) lists[(static_cast<size_t>(q) * 4 + w) * k + i] = 0ull;
    __syncwarp();
    const uint32_t tq = tmem_base + ((w * 32u) << 16);
    uint32_t tc = 0;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const uint32_t t0 = item * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
      for (uint32_t tile = t0; tile < t1; ++tile, ++tc) {
        const uint32_t as = tc & 1;
        ptx::mbar_wait_relaxed(&acc_full[as], (tc >> 1) & 1, p.epi_wait_ns);
        ptx::tc_fence_after();
        if (!(p.debug & 4u)) {
#pragma unroll 1
          for (uint32_t sub = 0; sub < 2; ++sub) {
            const uint32_t row0w = tile * 256u + sub * 128u + w * 32u;  // database row of lane 0
            if (row0w >= p.n_rows) break;
            const bool rvalid = row0w + lane < p.n_rows;
#pragma unroll 1
            for (uint32_t c = 0; c < p.n_pad; c += 16) {
              float v[16];
              ptx::tmem_ld_32x32b_x16(tq + as * 256u + sub * p.n_pad + c, v);
              float t[16];
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const float4 x = *reinterpret_cast<const float4*>(thr_w + c + 4 * g);
                t[4 * g] = x.x;
                t[4 * g + 1] = x.y;
                t[4 * g + 2] = x.z;
                t[4 * g + 3] = x.w;
              }
              bool hit = false;
#pragma unroll
              for (int j = 0; j < 16; ++j) hit |= v[j] > t[j];
              if (__any_sync(0xffffffffu, hit && rvalid)) {
                // which of the chunk's 16 queries have a survivor in some lane
                uint32_t qm = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) qm |= (rvalid && v[j] > t[j]) ? (1u << j) : 0u;
                qm = __reduce_or_sync(0xffffffffu, qm);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (!((qm >> j) & 1u)) continue;  // warp-uniform
                  const uint32_t q = c + j;
                  uint64_t* Lq = lists + (static_cast<size_t>(q) * 4 + w) * k;
                  uint32_t b = __ballot_sync(0xffffffffu, rvalid && v[j] > t[j]);
                  if (k <= 32u) {
                    // the (warp, query) list lives in the lanes while this query's survivors are inserted: lane i
                    // holds rank i; an insert is one ballot (position) and one shuffle-up (shift), no shared memory
                    uint64_t Li = static_cast<uint32_t>(lane) < k ? Lq[lane] : 0ull;
                    while (b) {
                      const int L = __ffs(b) - 1;
                      b &= b - 1;
                      const float sc = __shfl_sync(0xffffffffu, v[j], L);
                      const uint64_t key = make_key(sc, row0w + L);
                      const uint32_t pos = __popc(__ballot_sync(0xffffffffu, Li > key));
                      if (pos < k) {
                        const uint64_t up = __shfl_up_sync(0xffffffffu, Li, 1);
                        Li = static_cast<uint32_t>(lane) < pos ? Li : (static_cast<uint32_t>(lane) == pos ? key : up);
                        if (static_cast<uint32_t>(lane) >= k) Li = 0ull;
                      }
                    }
                    if (static_cast<uint32_t>(lane) < k) Lq[lane] = Li;
                    const uint64_t kth = __shfl_sync(0xffffffffu, Li, static_cast<int>(k) - 1);
                    if (lane == 0) thr_w[q] = (kth == 0ull) ? -INFINITY : key_score(kth);
                    __syncwarp();
                  } else {
                    while (b) {
                      const int L = __ffs(b) - 1;
                      b &= b - 1;
                      const float sc = __shfl_sync(0xffffffffu, v[j], L);
                      if (sc > thr_w[q]) {  // the threshold may have moved since the chunk's copy was read
                        const uint64_t kth = warp_list_insert(Lq, static_cast<int>(k), make_key(sc, row0w + L), lane);
                        if (lane == 0) thr_w[q] = (kth == 0ull) ? -INFINITY : key_score(kth);
                        __syncwarp();
                      }
                    }
                  }
                }
              }
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc_empty[as]);
      }
      // slice done: the four warps' lists of a query -> k keys in this CTA's slot
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (uint32_t q = w; q < p.nq; q += 4) {
        uint64_t* dst = p.partial + (static_cast<size_t>(q) * p.n_lists + blockIdx.x) * k;
        warp_multiway_merge<1>(lists + static_cast<size_t>(q) * 4 * k, 4u, k, k, lane,
                               [=](uint32_t r, uint64_t key) { dst[r] = key; });
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (item + gridDim.x < n_items) {  // another slice for this CTA: start its lists afresh
        for (uint32_t q = lane; q < p.n_pad; q += 32) thr_w[q] = q < p.nq ? -INFINITY : INFINITY;
        for (uint32_t q = 0; q < p.n_pad; ++q)
          for (uint32_t i = lane; i < k; i += 32) lists[(static_cast<size_t>(q) * 4 + w) * k + i] = 0ull;
        __syncwarp();
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

// ================================================================================================ 2-CTA kernel
namespace ptx2 {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release at CTA scope), as cutlass::arch::ClusterBarrier::arrive(cta_id) does: the
  // accumulator hand-over is ordered by tcgen05.fence::before/after_thread_sync, so no cluster-scope memory
  // fence (MEMBAR.GPU + ERRBAR in SASS, 10 % of the epilogue's samples) is needed
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive on the mbarrier at the same smem offset in the CTAs of `mask` once all prior MMAs completed
__device__ __forceinline__ void tc_commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(ptx::smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tc_mma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// TMA load executed by either CTA of the pair; the transaction bytes are credited to the mbarrier of the
// LEADER CTA (rank 0): clearing bit 24 of a shared::cta address names the same offset in CTA 0
// (cute::Sm100MmaPeerBitMask, cute/arch/copy_sm100_tma.hpp).
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, int32_t c0, int32_t c1,
                                                uint64_t* leader_bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(tmap), "r"(ptx::smem_u32(leader_bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1),
      "l"(policy)
      : "memory");
}
}  // namespace ptx2

constexpr uint32_t kD2HalfBytes = kDenseBM * kDenseBK * 2;  // 16 KB: 128 rows x 64 elements (A chunk or B half)
constexpr int kD2MaxKc = 8;                                 // resident A: d <= 512

// ARES: the CTA's 128 x d query tile stays in shared memory for a whole item; only the database streams.
// NS  : ring depth (ARES: stages of one B half; otherwise stages of A chunk + B half).
template <bool ARES, int NS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDenseThreads, 1)
scan_dense2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_db,
                   const DenseParams p) {
  constexpr int BN = 256;
  constexpr uint32_t kStage = ARES ? kD2HalfBytes : 2 * kD2HalfBytes;
  extern __shared__ uint8_t dense_smem_raw[];
  uint8_t* smem = dense_smem_raw + ((1024u - (ptx::smem_u32(dense_smem_raw) & 1023u)) & 1023u);
  uint8_t* a_res = smem;  // ARES only: kc chunks of 16 KB
  uint8_t* stages = smem + (ARES ? static_cast<size_t>(kD2MaxKc) * kD2HalfBytes : 0);
  uint8_t* epi_smem = stages + static_cast<size_t>(NS) * kStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + dense_epi_bytes(p.k, p.tp));
  uint64_t* full = bars;                     // leader's copy is the live one
  uint64_t* empty = bars + NS;               // each CTA its own (multicast commit)
  uint64_t* acc_full = bars + 2 * NS;        // each CTA its own (multicast commit)
  uint64_t* acc_empty = bars + 2 * NS + 2;   // leader's copy: 8 arrivals (4 epilogue warps x 2 CTAs)
  uint64_t* a_full = bars + 2 * NS + 4;      // leader's copy (ARES)
  uint64_t* a_free = bars + 2 * NS + 5;      // each CTA its own (ARES; multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = ptx2::cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_db);
    for (int s = 0; s < NS; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&acc_full[a], 1);
      ptx::mbar_init(&acc_empty[a], 8);
    }
    ptx::mbar_init(a_full, 1);
    ptx::mbar_init(a_free, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx2::tmem_alloc2(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx2::cluster_sync();  // both CTAs' barriers are initialised before anyone signals across the pair
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t n_items = p.m_tiles * p.n_slices;  // m_tiles counts tiles of 256 queries here

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs)
    if (lane == 0) {
      const uint64_t pol_q = ptx::policy_evict_last();
      const uint64_t pol_db = p.db_evict_first ? ptx::policy_evict_first() : ptx::policy_evict_last();
      uint32_t it = 0, ic = 0;
      for (uint32_t item = pair; item < n_items; item += n_pairs, ++ic) {
        const uint32_t m_tile = item % p.m_tiles, slice = item / p.m_tiles;
        const uint32_t t0 = slice * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
        const int32_t q_row = static_cast<int32_t>(m_tile * 256 + rank * kDenseBM);
        if (ARES) {
          if (ic > 0) ptx::mbar_wait(a_free, (ic - 1) & 1);  // previous item's MMAs no longer read A
          if (leader) ptx::mbar_expect_tx(a_full, 2 * p.kc * kD2HalfBytes);
          for (uint32_t kc = 0; kc < p.kc; ++kc)
            ptx2::tma_load_2d_2sm(a_res + static_cast<size_t>(kc) * kD2HalfBytes, &tm_q,
                                  static_cast<int32_t>(kc * kDenseBK), q_row, a_full, pol_q);
        }
        for (uint32_t tile = t0; tile < t1; ++tile) {
          const int32_t db_row = static_cast<int32_t>(tile * BN + rank * kDenseBM);
          for (uint32_t kc = 0; kc < p.kc; ++kc, ++it) {
            const uint32_t s = it % NS, u = it / NS;
            if (u > 0) ptx::mbar_wait(&empty[s], (u - 1) & 1);
            uint8_t* dst = stages + static_cast<size_t>(s) * kStage;
            if (leader) ptx::mbar_expect_tx(&full[s], 2 * kStage);  // both CTAs' bytes land on this barrier
            if (!ARES) {
              ptx2::tma_load_2d_2sm(dst, &tm_q, static_cast<int32_t>(kc * kDenseBK), q_row, &full[s], pol_q);
              dst += kD2HalfBytes;
            }
            ptx2::tma_load_2d_2sm(dst, &tm_db, static_cast<int32_t>(kc * kDenseBK), db_row, &full[s], pol_db);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA only)
    // The whole warp runs the loops on warp-uniform values and one elected lane issues (ptx::elect_one): the
    // descriptors stay in uniform registers.  ncu on C3 (10M x 768, nq = 4096) showed the issuing thread itself was
    // the critical path — ~105 instructions per K chunk, 663 cycles against 512 cycles of tensor work.
    if (leader) {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t stages_u = ptx::smem_u32(stages), a_res_u = ptx::smem_u32(a_res);
      const uint32_t idesc = p.idesc, kcs = p.kc;
      uint32_t s = 0, ph = 0, tc = 0, ic = 0;  // ring slot and its phase bit (no division in the loop)
      for (uint32_t item = pair; item < n_items; item += n_pairs, ++ic) {
        const uint32_t slice = item / p.m_tiles;
        const uint32_t t0 = slice * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
        if (ARES) {
          ptx::mbar_wait(a_full, ic & 1);
          ptx::tc_fence_after();
        }
        for (uint32_t tile = t0; tile < t1; ++tile, ++tc) {
          const uint32_t as = tc & 1, ua = tc >> 1;
          if (ua > 0) ptx::mbar_wait(&acc_empty[as], (ua - 1) & 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_u + as * BN;
          for (uint32_t kc = 0; kc < kcs; ++kc) {
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after();
            const uint32_t st_addr = stages_u + s * kStage;
            const uint32_t a_addr = ARES ? a_res_u + kc * kD2HalfBytes : st_addr;
            const uint64_t a_desc = ptx::umma_desc_k_sw128(a_addr);
            const uint64_t b_desc = ptx::umma_desc_k_sw128(ARES ? st_addr : st_addr + kD2HalfBytes);
            if (ptx::elect_one()) {
#pragma unroll
              for (uint32_t kk = 0; kk < kDenseBK / 16; ++kk)
                ptx2::tc_mma2_f16(d_tmem, a_desc + kk * 2, b_desc + kk * 2, idesc, (kc | kk) != 0 ? 1u : 0u);
              ptx2::tc_commit2(&empty[s], 3);
            }
            __syncwarp();
            if (++s == NS) {
              s = 0;
              ph ^= 1u;
            }
          }
          if (ptx::elect_one()) ptx2::tc_commit2(&acc_full[as], 3);
          __syncwarp();
        }
        if (ARES) {
          if (ptx::elect_one()) ptx2::tc_commit2(a_free, 3);
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================================ epilogue (both CTAs): fused top-k
    const uint32_t acc_empty_leader0 = ptx2::mapa(ptx::smem_u32(&acc_empty[0]), 0);
    dense_epilogue<BN>(p, pair, n_pairs, 256u, rank * kDenseBM, tmem_base, epi_smem, acc_full, warp, lane,
                       [&](uint32_t as) { ptx2::mbar_arrive_cluster(acc_empty_leader0 + as * 8); });
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx2::cluster_sync();  // the peer's smem / TMEM / barriers stay alive until both CTAs are done
  if (warp == 1) ptx2::tmem_dealloc2(tmem_base, 512);
}

// ================================================================================================ 2-CTA kernel, database tile resident
// Large shards, several query tiles (nq > 256).  scan_dense2_kernel keeps a QUERY tile resident and streams a
// database slice past it; the query tiles that share a slice are meant to meet in L2, but over the 10.5k tiles of
// a 100M-row slice they drift apart and the database is read from DRAM ~10 times (988 GB for 102.4 GB, ncu).
// Here the roles are swapped: the pair keeps one DATABASE tile (256 rows x d, half per CTA, d <= 512) resident
// and streams every query tile of the block past it (the whole query block is a few MB and lives in L2).  The
// database is read from DRAM exactly once per query block BY CONSTRUCTION — no lockstep, nothing to drift — and
// L2 -> SM traffic per flop is what it was.  The TMEM layout does not change (lane = query, column = database row).
//
// Pipeline: the resident tile is a ring of 8 chunk slots (16 KB: this CTA's 128 rows x 64 elements); chunk c of
// the pair's next tile replaces chunk c of the current one as soon as the LAST query tile's MMAs have read it
// (b_free[slot], committed by the MMA thread), i.e. while the last query tile is still being multiplied: the
// tensor pipe never drains between database tiles.  The query chunks stream through the 4-stage ring as before.
//
// Top-k: a thread now serves m_tiles queries in turn (one per query tile), so the per-query state cannot stay in
// registers / shared-memory lists.  Every (CTA, query) pair owns a small RESERVOIR in global memory (C = 32 keys
// for k <= 16, 64 for k <= 32, max(256, 2 next_pow2(k)) beyond): a survivor is appended with one fire-and-forget
// st.global.cg; a full reservoir is compacted to its k best by the warp (rank sort in registers for C <= 64,
// bisection over the key bits beyond) and the k-th score becomes the query's threshold.  Only the threshold and
// the fill count live on chip: thr_s / cnt_s [m_tiles][128] in shared memory.  A stale threshold only admits
// extra candidates; expected appends per (pair, query) ~ (C-k) log(n/C) / log(C/k) (about 200 for 1.35M rows at
// k = 10) — a dozen compactions per reservoir for the whole scan.
struct DenseBParams {
  uint64_t* partial;   // [nq][n_lists][k] keys out (list u = CTA pair u)
  uint64_t* lists_ws;  // [grid][m_tiles][128][res_cap] reservoirs
  uint32_t n_rows, nq, k, res_cap;
  uint32_t m_tiles;    // query tiles of 256
  uint32_t n_tiles;    // database tiles of 256 rows; pair u owns tiles u, u + n_pairs, ...
  uint32_t n_lists;
  uint32_t kc, idesc, debug;
  uint32_t epi_wait_ns;  // epilogue warps sleep this long after a failed try on acc_full (0: poll)
};

constexpr int kD2bSlots = 8;   // resident database chunk slots (d <= 512)
constexpr int kD2bStages = 4;  // query chunk ring

__host__ __device__ __forceinline__ size_t dense_b_smem_bytes(uint32_t m_tiles) {
  return 1024 + static_cast<size_t>(kD2bSlots + kD2bStages) * (kDenseBM * kDenseBK * 2) +
         static_cast<size_t>(m_tiles) * kDenseBM * 8 + 256;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDenseThreads, 1)
scan_dense2b_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_db,
                    const DenseBParams p) {
  constexpr int BN = 256, NS = kD2bStages, NB = kD2bSlots;
  extern __shared__ uint8_t dense_smem_raw[];
  uint8_t* smem = dense_smem_raw + ((1024u - (ptx::smem_u32(dense_smem_raw) & 1023u)) & 1023u);
  uint8_t* b_res = smem;                                            // NB slots of 16 KB
  uint8_t* stages = smem + static_cast<size_t>(NB) * kD2HalfBytes;  // NS stages of 16 KB
  float* thr_s = reinterpret_cast<float*>(stages + static_cast<size_t>(NS) * kD2HalfBytes);  // [m_tiles][128]
  uint32_t* cnt_s = reinterpret_cast<uint32_t*>(thr_s + static_cast<size_t>(p.m_tiles) * kDenseBM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(cnt_s + static_cast<size_t>(p.m_tiles) * kDenseBM);
  uint64_t* full = bars;                         // leader's copy is the live one
  uint64_t* empty = bars + NS;                   // each CTA its own (multicast commit)
  uint64_t* b_full = bars + 2 * NS;              // leader's copy
  uint64_t* b_free = bars + 2 * NS + NB;         // each CTA its own (multicast commit)
  uint64_t* acc_full = bars + 2 * NS + 2 * NB;   // each CTA its own (multicast commit)
  uint64_t* acc_empty = acc_full + 2;            // leader's copy: 8 arrivals (4 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = ptx2::cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_db);
    for (int s = 0; s < NS; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < NB; ++s) {
      ptx::mbar_init(&b_full[s], 1);
      ptx::mbar_init(&b_free[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&acc_full[a], 1);
      ptx::mbar_init(&acc_empty[a], 8);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx2::tmem_alloc2(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx2::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t KC = p.kc, MT = p.m_tiles;
  const uint32_t n_my = pair < p.n_tiles ? (p.n_tiles - pair + n_pairs - 1) / n_pairs : 0u;  // this pair's tiles

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs)
    if (lane == 0) {
      const uint64_t pol_q = ptx::policy_evict_last();    // the query block is re-read by every pair, every tile
      const uint64_t pol_db = ptx::policy_evict_first();  // the database passes through once
      const uint32_t total_b = n_my * KC;
      uint32_t nb = 0;  // next database chunk (sequence number: tile index * KC + chunk) to request
      auto issue_b = [&](uint32_t b) {
        const uint32_t slot = b % NB, gen = b / NB;
        if (gen > 0) ptx::mbar_wait(&b_free[slot], (gen - 1) & 1);
        if (leader) ptx::mbar_expect_tx(&b_full[slot], 2 * kD2HalfBytes);
        const uint32_t tile = pair + (b / KC) * n_pairs, c = b % KC;
        ptx2::tma_load_2d_2sm(b_res + static_cast<size_t>(slot) * kD2HalfBytes, &tm_db,
                              static_cast<int32_t>(c * kDenseBK), static_cast<int32_t>(tile * BN + rank * kDenseBM),
                              &b_full[slot], pol_db);
      };
      while (nb < total_b && nb < static_cast<uint32_t>(NB)) issue_b(nb++);
      uint32_t g = 0;  // step = (tile index, query tile, chunk)
      for (uint32_t ti = 0; ti < n_my; ++ti) {
        for (uint32_t mi = 0; mi < MT; ++mi) {
          const uint32_t m = (mi + pair) % MT;  // pairs start at different query tiles: no L2 hot spot
          const int32_t q_row = static_cast<int32_t>(m * 256 + rank * kDenseBM);
          for (uint32_t kc = 0; kc < KC; ++kc, ++g) {
            const uint32_t s = g % NS, u = g / NS;
            if (u > 0) ptx::mbar_wait(&empty[s], (u - 1) & 1);  // the MMAs of step g - NS are done
            if (leader) ptx::mbar_expect_tx(&full[s], 2 * kD2HalfBytes);
            ptx2::tma_load_2d_2sm(stages + static_cast<size_t>(s) * kD2HalfBytes, &tm_q,
                                  static_cast<int32_t>(kc * kDenseBK), q_row, &full[s], pol_q);
            // database chunks whose slot saw its last use at a step <= g - NS can be replaced now
            while (nb < total_b && g >= static_cast<uint32_t>(NS)) {
              const uint32_t prev = nb - NB;
              const uint32_t last_use = ((prev / KC) * MT + (MT - 1)) * KC + prev % KC;
              if (last_use + NS > g) break;
              issue_b(nb++);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA only)
    // whole warp, warp-uniform operands, one elected lane issues (see ptx::elect_one)
    if (leader) {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t stages_u = ptx::smem_u32(stages), b_res_u = ptx::smem_u32(b_res);
      const uint32_t idesc = p.idesc;
      uint32_t g = 0, tc = 0;
      for (uint32_t ti = 0; ti < n_my; ++ti) {
        for (uint32_t mi = 0; mi < MT; ++mi, ++tc) {
          const uint32_t as = tc & 1, ua = tc >> 1;
          if (ua > 0) ptx::mbar_wait(&acc_empty[as], (ua - 1) & 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_u + as * BN;
          for (uint32_t kc = 0; kc < KC; ++kc, ++g) {
            const uint32_t s = g % NS;
            const uint32_t b = ti * KC + kc, slot = b % NB;
            ptx::mbar_wait(&b_full[slot], (b / NB) & 1);  // completes once per tile; later query tiles pass at once
            ptx::mbar_wait(&full[s], (g / NS) & 1);
            ptx::tc_fence_after();
            const uint64_t a_desc = ptx::umma_desc_k_sw128(stages_u + s * kD2HalfBytes);
            const uint64_t b_desc = ptx::umma_desc_k_sw128(b_res_u + slot * kD2HalfBytes);
            if (ptx::elect_one()) {
#pragma unroll
              for (uint32_t kk = 0; kk < kDenseBK / 16; ++kk)
                ptx2::tc_mma2_f16(d_tmem, a_desc + kk * 2, b_desc + kk * 2, idesc, (kc | kk) != 0 ? 1u : 0u);
              ptx2::tc_commit2(&empty[s], 3);
              if (mi + 1 == MT) ptx2::tc_commit2(&b_free[slot], 3);  // last query tile: the chunk slot may be refilled
            }
            __syncwarp();
          }
          if (ptx::elect_one()) ptx2::tc_commit2(&acc_full[as], 3);
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================================ epilogue (both CTAs): fused top-k
    const uint32_t acc_empty_leader0 = ptx2::mapa(ptx::smem_u32(&acc_empty[0]), 0);
    const int row0w = (warp & 3) * 32;  // TMEM lane quadrant this warp may read = warp_id % 4
    const int t = row0w + lane;
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(row0w) << 16);
    for (uint32_t m = 0; m < MT; ++m) {  // a thread only ever touches its own entries: no barrier needed
      thr_s[m * kDenseBM + t] = -INFINITY;
      cnt_s[m * kDenseBM + t] = 0u;
    }
    DenseEpiRes epi;
    epi.k = p.k;
    epi.C = p.res_cap;
    epi.lane = lane;
    uint64_t* res_cta = p.lists_ws + static_cast<size_t>(blockIdx.x) * MT * kDenseBM * p.res_cap;
    uint32_t tc = 0;
    for (uint32_t ti = 0; ti < n_my; ++ti) {
      const uint32_t row0 = (pair + ti * n_pairs) * BN;
      const uint32_t n_valid = min(static_cast<uint32_t>(BN), p.n_rows - row0);
      for (uint32_t mi = 0; mi < MT; ++mi, ++tc) {
        const uint32_t m = (mi + pair) % MT;
        const uint32_t q0 = m * 256 + rank * kDenseBM;
        const bool q_valid = q0 + t < p.nq;
        const bool warp_valid = q0 + row0w < p.nq;
        const uint32_t as = tc & 1;
        ptx::mbar_wait_relaxed(&acc_full[as], (tc >> 1) & 1, p.epi_wait_ns);
        ptx::tc_fence_after();
        if (warp_valid && !(p.debug & 4u)) {
          epi.res_warp = res_cta + (static_cast<size_t>(m) * kDenseBM + row0w) * p.res_cap;
          epi.thr = thr_s[m * kDenseBM + t];
          epi.cnt = cnt_s[m * kDenseBM + t];
          epi.template scan_tile<BN>(tq + as * BN, row0, n_valid, q_valid);
          thr_s[m * kDenseBM + t] = epi.thr;
          cnt_s[m * kDenseBM + t] = epi.cnt;
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx2::mbar_arrive_cluster(acc_empty_leader0 + as * 8);
      }
    }
    // every reservoir down to its k best -> this pair's list of the query.  Small reservoirs come out of the rank
    // sort in order (so even a reservoir that never filled is passed through it); large ones stay unsorted.
    const bool small = p.res_cap <= 64u;
    for (uint32_t m = 0; m < MT; ++m) {
      const uint32_t q0w = m * 256 + rank * kDenseBM + row0w;
      if (q0w >= p.nq) continue;
      epi.res_warp = res_cta + (static_cast<size_t>(m) * kDenseBM + row0w) * p.res_cap;
      epi.thr = thr_s[m * kDenseBM + t];
      epi.cnt = cnt_s[m * kDenseBM + t];
      const uint32_t need = __ballot_sync(0xffffffffu, small ? epi.cnt > 0u : epi.cnt > p.k);
      if (need) epi.compact(need);
      __syncwarp();
      for (int L = 0; L < 32; ++L) {
        const uint32_t q = q0w + L;
        if (q >= p.nq) break;
        const uint32_t n = __shfl_sync(0xffffffffu, epi.cnt, L);
        const uint64_t* R = epi.res_warp + static_cast<size_t>(L) * p.res_cap;
        uint64_t* dst = p.partial + (static_cast<size_t>(q) * p.n_lists + pair) * p.k;
        for (uint32_t i = lane; i < p.k; i += 32) __stcg(dst + i, i < n ? __ldcg(R + i) : 0ull);
      }
      __syncwarp();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx2::cluster_sync();
  if (warp == 1) ptx2::tmem_dealloc2(tmem_base, 512);
}


}  // namespace sgic
