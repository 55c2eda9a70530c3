// K4 — dense Q·Dᵀ on the 5th-gen tensor cores with the top-k fused into the epilogue
// (replaces FAISS exhaustive_inner_product_blas = sgemm blocks + heap/reservoir
// behind src/search.py:115 for batches of queries).
//
// Layout: queries on M (one TMEM lane = one query), database rows on N, d on K.
//   tile      : 128 queries x BN (256) database rows, K streamed in chunks of 64 (one 128-byte swizzle row)
//   operands  : TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) -> shared-memory ring of NS stages,
//               each stage = A chunk (128x64) + B chunk (BNx64); completion on mbarriers
//   MMA       : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16), fp32
//               accumulators in TMEM, two accumulator stages (2 x BN columns) so the epilogue of tile i
//               overlaps the MMAs of tile i+1; tcgen05.commit releases smem stages / publishes accumulators
//   epilogue  : 4 warps; thread t owns TMEM lane t = query t.  tcgen05.ld 32 columns at a time, one
//               3-input-max tree + one compare against the query's running k-th score rejects a whole
//               32-score chunk; survivors are appended to a per-query buffer in shared memory and merged
//               into the query's sorted list by the whole warp (warp_list_insert) when a buffer fills.
//               No score matrix ever leaves the SM.
// Work split: items = (query tile m, database slice s), m fastest, so CTAs that run concurrently share a
// slice and each database tile is fetched from HBM once and re-used from L2 by the other query tiles.
// Every item writes k keys per query; merge_keys_small_kernel merges the slices.
#pragma once
#include <cuda.h>

#include "topk_common.cuh"

namespace sgic {

constexpr int kDenseThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int kDenseBM = 128;
constexpr int kDenseBK = 64;
constexpr int kDenseBuf = 16;  // candidate buffer depth per query (entries)

struct DenseParams {
  uint64_t* partial;   // [nq][n_slices][k] keys
  uint64_t* lists_ws;  // [grid][128][kp] keys, used when the lists do not fit in shared memory
  uint32_t n_rows, nq, k, kp;
  uint32_t m_tiles, n_slices, tiles_per_slice, n_tiles;
  uint32_t kc;             // K chunks of 64 elements (ceil(d/64); TMA zero-fills the tail)
  uint32_t lists_in_smem;  // 128*kp*8 bytes fit next to the stages
  uint32_t idesc;          // UMMA instruction descriptor
  uint32_t db_evict_first; // single query tile: the database is streamed once -> evict_first
};

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

template <int BN, int NS>
__global__ void __launch_bounds__(kDenseThreads, 1)
scan_dense_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_db,
                  const DenseParams p) {
  using Cfg = DenseCfg<BN>;
  extern __shared__ uint8_t dense_smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* smem = dense_smem_raw + ((1024u - (ptx::smem_u32(dense_smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;
  uint64_t* cand = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(NS) * Cfg::kStageBytes);
  uint64_t* lists_sm = cand + kDenseBuf * kDenseBM;
  uint64_t* bars = lists_sm + (p.lists_in_smem ? static_cast<size_t>(kDenseBM) * p.kp : 0);
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
            ptx::mbar_expect_tx(&full[s], Cfg::kStageBytes);
            ptx::tma_load_2d(a_dst, &tm_q, static_cast<int32_t>(kc * kDenseBK), static_cast<int32_t>(m_tile * kDenseBM),
                             &full[s], pol_q);
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
    const int wq = warp & 3;              // TMEM lane quadrant this warp may read
    const int row0w = wq * 32;            // first query row (TMEM lane) of this warp
    const int t = row0w + lane;           // my query row inside the tile
    uint64_t* lists = p.lists_in_smem ? lists_sm
                                      : p.lists_ws + static_cast<size_t>(blockIdx.x) * kDenseBM * p.kp;
    uint32_t tc = 0;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const uint32_t m_tile = item % p.m_tiles, slice = item / p.m_tiles;
      const uint32_t t0 = slice * p.tiles_per_slice, t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
      const bool q_valid = m_tile * kDenseBM + t < p.nq;
      for (int L = 0; L < 32; ++L)
        for (uint32_t i = lane; i < p.kp; i += 32) lists[static_cast<size_t>(row0w + L) * p.kp + i] = 0ull;
      uint32_t cnt = 0;
      float thr = -INFINITY;
      __syncwarp();
      for (uint32_t tile = t0; tile < t1; ++tile, ++tc) {
        const uint32_t as = tc & 1;
        ptx::mbar_wait(&acc_full[as], (tc >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t row0 = tile * BN;
        const uint32_t n_valid = min(static_cast<uint32_t>(BN), p.n_rows - row0);
        const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(row0w) << 16) + as * BN;
#pragma unroll 1
        for (uint32_t c = 0; c < BN / 32; ++c) {
          float v[32];
          ptx::tmem_ld_32x32b_x32(taddr0 + c * 32, v);
          float m = v[0];
#pragma unroll
          for (int i = 1; i < 32; ++i) m = fmaxf(m, v[i]);
          const bool hit = q_valid && (m > thr) && (c * 32 < n_valid);
          if (__any_sync(0xffffffffu, hit)) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (__any_sync(0xffffffffu, cnt > kDenseBuf - 8))
                dense_warp_flush(cand, lists, p.kp, p.k, row0w, lane, cnt, thr);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t col = c * 32 + g * 8 + j;
                if (q_valid && v[g * 8 + j] > thr && col < n_valid) {
                  cand[cnt * kDenseBM + t] = make_key(v[g * 8 + j], row0 + col);
                  ++cnt;
                }
              }
            }
          }
        }
        // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc_empty[as]);
      }
      dense_warp_flush(cand, lists, p.kp, p.k, row0w, lane, cnt, thr);
      for (int L = 0; L < 32; ++L) {
        const uint32_t q = m_tile * kDenseBM + row0w + L;
        if (q >= p.nq) break;
        uint64_t* dst = p.partial + (static_cast<size_t>(q) * p.n_slices + slice) * p.k;
        const uint64_t* src = lists + static_cast<size_t>(row0w + L) * p.kp;
        for (uint32_t i = lane; i < p.k; i += 32) dst[i] = src[i];
      }
      __syncwarp();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace sgic
