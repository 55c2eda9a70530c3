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
//               always fresh.  k > 32: survivors go to a per-query buffer and the warp merges them into the
//               query's list together (warp_list_insert).  No score matrix ever leaves the SM.
// Work split: items = (query tile m, database slice s), m fastest, so CTAs that run concurrently share a
// slice and each database tile is fetched from HBM once and re-used from L2 by the other query tiles.
// Every item writes k keys per query; merge_keys_small_kernel merges the slices.
//
// Two kernels:
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
constexpr int kDenseBuf = 16;       // candidate buffer depth per query (entries)

struct DenseParams {
  uint64_t* partial;   // [nq][n_slices][k] keys
  uint64_t* lists_ws;  // [grid][128][kp] keys, used when the lists do not fit in shared memory
  uint32_t n_rows, nq, k, kp;
  uint32_t m_tiles;        // query tiles: of 128 (1-CTA kernel) or 256 (2-CTA kernel)
  uint32_t n_slices, tiles_per_slice, n_tiles;
  uint32_t kc;             // K chunks of 64 elements (ceil(d/64); TMA zero-fills the tail)
  uint32_t lists_in_smem;  // k > 32 only: 128*kp*8 bytes fit next to the stages
  uint32_t tp;             // k <= 32: thread-private lists [k][128] in shared memory, no candidate buffer
  uint32_t a_rows;         // rows of the query box (1-CTA kernel; < 128 for small batches)
  uint32_t idesc;          // UMMA instruction descriptor
  uint32_t db_evict_first; // single query tile: the database is streamed once -> evict_first
  uint32_t debug;          // timing experiments only (wrong results): 1 skip A loads, 2 skip B loads, 4 skip epilogue
};

// shared memory of the epilogue: thread-private lists (k <= 32) or candidate buffer (+ lists when they fit)
__host__ __device__ __forceinline__ uint32_t dense_epi_bytes(uint32_t k, uint32_t kp, uint32_t tp,
                                                             uint32_t lists_in_smem) {
  if (tp) return k * kDenseBM * 8u;
  return kDenseBuf * kDenseBM * 8u + (lists_in_smem ? kDenseBM * kp * 8u : 0u);
}

template <int BN>
struct DenseCfg {
  static constexpr uint32_t kABytes = kDenseBM * kDenseBK * 2;
  static constexpr uint32_t kBBytes = BN * kDenseBK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BN;
};

// Merge every lane's buffered candidates into that lane's sorted list; all 32 lanes cooperate on one
// list at a time (ballot-count position + parallel shift).  Updates each lane's threshold.
__device__ __noinline__ void dense_warp_flush(const uint64_t* cand, uint64_t* lists, uint32_t kp, uint32_t k,
                                              int row0, int lane, uint32_t& cnt, float& thr) {
  __syncwarp();
  for (int L = 0; L < 32; ++L) {
    const uint32_t n = __shfl_sync(0xffffffffu, cnt, L);
    if (n == 0) continue;
    uint64_t* list = lists + static_cast<size_t>(row0 + L) * kp;
    uint64_t kth = 0ull;
    for (uint32_t e = 0; e < n; ++e) kth = warp_list_insert(list, static_cast<int>(k), cand[e * kDenseBM + row0 + L], lane);
    if (lane == L) thr = (kth == 0ull) ? -INFINITY : key_score(kth);
  }
  cnt = 0;
  __syncwarp();
}

// Per-warp epilogue state and the scan of one accumulator tile (BN columns of this warp's 32 TMEM lanes).
struct DenseEpi {
  uint64_t* cand;   // [kDenseBuf][128] candidate buffer (shared memory)
  uint64_t* lists;  // [128][kp] sorted lists of this CTA (shared or global)
  uint32_t kp, k;
  int row0w, lane, t;
  uint32_t cnt;
  float thr;

  __device__ __forceinline__ void reset() {
    for (int L = 0; L < 32; ++L)
      for (uint32_t i = lane; i < kp; i += 32) lists[static_cast<size_t>(row0w + L) * kp + i] = 0ull;
    cnt = 0;
    thr = -INFINITY;
    __syncwarp();
  }
  __device__ __forceinline__ void flush() { dense_warp_flush(cand, lists, kp, k, row0w, lane, cnt, thr); }

  // taddr0: TMEM address of column 0 of the tile for this warp's lane quadrant
  template <int BN>
  __device__ __forceinline__ void scan_tile(uint32_t taddr0, uint32_t row0, uint32_t n_valid, bool q_valid) {
#pragma unroll 1
    for (uint32_t c = 0; c < BN / 64; ++c) {
      float v[64];
      ptx::tmem_ld_32x32b_x64(taddr0 + c * 64, v);
      // per-group-of-8 maxima as balanced trees (3-input max): short dependency chains
      float m8[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float a = fmaxf(fmaxf(v[8 * g], v[8 * g + 1]), v[8 * g + 2]);
        const float b = fmaxf(fmaxf(v[8 * g + 3], v[8 * g + 4]), v[8 * g + 5]);
        m8[g] = fmaxf(fmaxf(a, b), fmaxf(v[8 * g + 6], v[8 * g + 7]));
      }
      const float m = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])),
                            fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
      const bool hit = q_valid && (m > thr) && (c * 64 < n_valid);
      if (__any_sync(0xffffffffu, hit)) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (!__any_sync(0xffffffffu, q_valid && m8[g] > thr)) continue;  // no lane has a survivor here
          if (__any_sync(0xffffffffu, cnt > kDenseBuf - 8)) flush();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t col = c * 64 + g * 8 + j;
            if (q_valid && v[g * 8 + j] > thr && col < n_valid) {
              cand[cnt * kDenseBM + t] = make_key(v[g * 8 + j], row0 + col);
              ++cnt;
            }
          }
        }
      }
    }
  }

  // item finished: every query's list -> partial[(q * n_slices + slice) * k]
  __device__ __forceinline__ void store_lists(uint64_t* partial, uint32_t q0, uint32_t nq, uint32_t n_slices,
                                              uint32_t slice) {
    flush();
    for (int L = 0; L < 32; ++L) {
      const uint32_t q = q0 + row0w + L;
      if (q >= nq) break;
      uint64_t* dst = partial + (static_cast<size_t>(q) * n_slices + slice) * k;
      const uint64_t* src = lists + static_cast<size_t>(row0w + L) * kp;
      for (uint32_t i = lane; i < k; i += 32) dst[i] = src[i];
    }
    __syncwarp();
  }
};

// k <= 32: thread t keeps the sorted list of ITS query at lst[rank * 128] (shared memory; the rank stride of
// 1 KB keeps every lane on its own bank pair whatever rank it touches).  A survivor is inserted by its own
// thread with a short shift loop — 32 queries insert in parallel and `thr` is always the current k-th score.
struct DenseEpiTP {
  uint64_t* lst;
  uint32_t k;
  float thr;

  __device__ __forceinline__ void reset() {
    for (uint32_t i = 0; i < k; ++i) lst[i * kDenseBM] = 0ull;
    thr = -INFINITY;
  }
  // requires key > current k-th key (the caller compared the score with thr)
  __device__ __forceinline__ void insert(uint64_t key) {
    uint32_t i = k - 1;
#pragma unroll 1
    while (i > 0) {
      const uint64_t a = lst[(i - 1) * kDenseBM];
      if (a > key) break;
      lst[i * kDenseBM] = a;
      --i;
    }
    lst[i * kDenseBM] = key;
    const uint64_t kth = lst[(k - 1) * kDenseBM];
    thr = (kth == 0ull) ? -INFINITY : key_score(kth);
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
        // branches), then the warp walks the UNION of the lanes' survivor columns and re-reads each one from
        // TMEM (the accumulator is the only dynamically indexable copy of the scores): one single-column
        // tcgen05.ld per distinct column, the owning lanes insert.
        const uint32_t col0 = c * 32;
        uint32_t mask = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) mask |= (v[j] > thr) ? (1u << j) : 0u;
        if (!q_valid) mask = 0;
        if (n_valid - col0 < 32u) mask &= (1u << (n_valid - col0)) - 1u;
        uint32_t um = __reduce_or_sync(0xffffffffu, mask);
        while (um) {
          const uint32_t j = __ffs(um) - 1;
          um &= um - 1;
          const float sc = ptx::tmem_ld_32x32b_x1(taddr0 + col0 + j);
          if (((mask >> j) & 1u) && sc > thr) insert(make_key(sc, row0 + col0 + j));
          __syncwarp();
        }
      }
      __syncwarp();  // tcgen05.ld is warp-collective: reconverge before the next chunk
    }
  }

  __device__ __forceinline__ void store(uint64_t* partial, uint32_t q, uint32_t nq, uint32_t n_slices, uint32_t slice) {
    if (q >= nq) return;
    uint64_t* dst = partial + (static_cast<size_t>(q) * n_slices + slice) * k;
    for (uint32_t i = 0; i < k; ++i) dst[i] = lst[i * kDenseBM];
  }
};

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
      epi.reset();
      for (uint32_t tile = t0; tile < t1; ++tile, ++tc) {
        const uint32_t as = tc & 1;
        ptx::mbar_wait(&acc_full[as], (tc >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t row0 = tile * BN;
        const uint32_t n_valid = min(static_cast<uint32_t>(BN), p.n_rows - row0);
        if (warp_valid && !(p.debug & 4u)) epi.template scan_tile<BN>(tq + as * BN, row0, n_valid, q_valid);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(as);
      }
      epi.store(p.partial, q0 + t, p.nq, p.n_slices, slice);
    }
  } else {
    DenseEpi epi;
    epi.cand = reinterpret_cast<uint64_t*>(epi_smem);
    epi.lists = p.lists_in_smem ? epi.cand + kDenseBuf * kDenseBM
                                : p.lists_ws + static_cast<size_t>(blockIdx.x) * kDenseBM * p.kp;
    epi.kp = p.kp;
    epi.k = p.k;
    epi.row0w = row0w;
    epi.lane = lane;
    epi.t = t;
    for (uint32_t item = unit; item < n_items; item += n_units) {
      const uint32_t m_tile = item % p.m_tiles, slice = item / p.m_tiles;
      const uint32_t t0 = slice * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
      const uint32_t q0 = m_tile * q_tile + q_off;
      const bool q_valid = q0 + t < p.nq;
      const bool warp_valid = q0 + row0w < p.nq;
      epi.reset();
      for (uint32_t tile = t0; tile < t1; ++tile, ++tc) {
        const uint32_t as = tc & 1;
        ptx::mbar_wait(&acc_full[as], (tc >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t row0 = tile * BN;
        const uint32_t n_valid = min(static_cast<uint32_t>(BN), p.n_rows - row0);
        if (warp_valid && !(p.debug & 4u)) epi.template scan_tile<BN>(tq + as * BN, row0, n_valid, q_valid);
        // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(as);
        // keep thresholds fresh: merge buffered survivors as soon as a few have piled up
        if (__any_sync(0xffffffffu, epi.cnt >= 4)) epi.flush();
      }
      epi.store_lists(p.partial, q0, p.nq, p.n_slices, slice);
    }
  }
}

// ================================================================================================ 1-CTA kernel
template <int BN, int NS>
__global__ void __launch_bounds__(kDenseThreads, 1)
scan_dense_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_db,
                  const DenseParams p) {
  using Cfg = DenseCfg<BN>;
  extern __shared__ uint8_t dense_smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* smem = dense_smem_raw + ((1024u - (ptx::smem_u32(dense_smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;
  uint8_t* epi_smem = smem + static_cast<size_t>(NS) * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + dense_epi_bytes(p.k, p.kp, p.tp, p.lists_in_smem));
  uint64_t* full = bars;
  uint64_t* empty = bars + NS;
  uint64_t* acc_full = bars + 2 * NS;
  uint64_t* acc_empty = bars + 2 * NS + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_db);
    for (int s = 0; s < NS; ++s) {
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
      uint32_t it = 0;
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t m_tile = item % p.m_tiles, slice = item / p.m_tiles;
        const uint32_t t0 = slice * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
        for (uint32_t tile = t0; tile < t1; ++tile) {
          for (uint32_t kc = 0; kc < p.kc; ++kc, ++it) {
            const uint32_t s = it % NS, u = it / NS;
            if (u > 0) ptx::mbar_wait(&empty[s], (u - 1) & 1);
            uint8_t* a_dst = stages + static_cast<size_t>(s) * Cfg::kStageBytes;
            const bool ld_a = !(p.debug & 1u) || it < NS, ld_b = !(p.debug & 2u) || it < NS;
            ptx::mbar_expect_tx(&full[s], (ld_a ? p.a_rows * (kDenseBK * 2u) : 0u) + (ld_b ? Cfg::kBBytes : 0u));
            if (ld_a)
              ptx::tma_load_2d(a_dst, &tm_q, static_cast<int32_t>(kc * kDenseBK),
                               static_cast<int32_t>(m_tile * kDenseBM), &full[s], pol_q);
            if (ld_b)
              ptx::tma_load_2d(a_dst + Cfg::kABytes, &tm_db, static_cast<int32_t>(kc * kDenseBK),
                               static_cast<int32_t>(tile * BN), &full[s], pol_db);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      uint32_t it = 0, tc = 0;
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t slice = item / p.m_tiles;
        const uint32_t t0 = slice * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
        for (uint32_t tile = t0; tile < t1; ++tile, ++tc) {
          const uint32_t as = tc & 1, ua = tc >> 1;
          if (ua > 0) ptx::mbar_wait(&acc_empty[as], (ua - 1) & 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BN;
          for (uint32_t kc = 0; kc < p.kc; ++kc, ++it) {
            const uint32_t s = it % NS;
            ptx::mbar_wait(&full[s], (it / NS) & 1);
            ptx::tc_fence_after();
            const uint32_t a_addr = ptx::smem_u32(stages + static_cast<size_t>(s) * Cfg::kStageBytes);
            const uint64_t a_desc = ptx::umma_desc_k_sw128(a_addr);
            const uint64_t b_desc = ptx::umma_desc_k_sw128(a_addr + Cfg::kABytes);
#pragma unroll
            for (uint32_t kk = 0; kk < kDenseBK / 16; ++kk)
              ptx::tc_mma_f16(d_tmem, a_desc + kk * 2, b_desc + kk * 2, p.idesc, (kc | kk) != 0 ? 1u : 0u);
            ptx::tc_commit(&empty[s]);  // smem stage reusable once these MMAs have read it
          }
          ptx::tc_commit(&acc_full[as]);  // accumulator of this tile complete
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
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + dense_epi_bytes(p.k, p.kp, p.tp, p.lists_in_smem));
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
    if (leader && lane == 0) {
      uint32_t it = 0, tc = 0, ic = 0;
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
          const uint32_t d_tmem = tmem_base + as * BN;
          for (uint32_t kc = 0; kc < p.kc; ++kc, ++it) {
            const uint32_t s = it % NS;
            ptx::mbar_wait(&full[s], (it / NS) & 1);
            ptx::tc_fence_after();
            const uint32_t st_addr = ptx::smem_u32(stages + static_cast<size_t>(s) * kStage);
            const uint32_t a_addr = ARES ? ptx::smem_u32(a_res + static_cast<size_t>(kc) * kD2HalfBytes) : st_addr;
            const uint64_t a_desc = ptx::umma_desc_k_sw128(a_addr);
            const uint64_t b_desc = ptx::umma_desc_k_sw128(ARES ? st_addr : st_addr + kD2HalfBytes);
#pragma unroll
            for (uint32_t kk = 0; kk < kDenseBK / 16; ++kk)
              ptx2::tc_mma2_f16(d_tmem, a_desc + kk * 2, b_desc + kk * 2, p.idesc, (kc | kk) != 0 ? 1u : 0u);
            ptx2::tc_commit2(&empty[s], 3);
          }
          ptx2::tc_commit2(&acc_full[as], 3);
        }
        if (ARES) ptx2::tc_commit2(a_free, 3);
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

}  // namespace sgic
