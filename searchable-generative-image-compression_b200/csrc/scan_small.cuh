// K3 — streaming score + top-k for 1..4 queries (replaces FAISS
// exhaustive_inner_product_seq = fvec_inner_product + heap behind src/search.py:115).
//
// HBM-bound: the database is read exactly once per call (algorithmic bytes = N*d*2).
// One persistent CTA per SM.  A producer thread streams contiguous row tiles into a
// ring of shared-memory stages with 1-D bulk async copies (TMA, SASS UBLKCP) that
// complete on mbarriers; 8 consumer warps read the staged rows with 128-bit LDS, form
// the dot products with mixed-precision FMAs (fp16/bf16 x fp16/bf16 -> fp32, SASS
// FHFMA; products exact, fp32 accumulation), reduce RB rows x NQ queries at once with a
// transposing butterfly, and keep a warp-private sorted top-k in shared memory guarded
// by a register threshold.  No score matrix is materialised.  At the end the CTA merges
// its warps' lists (bitonic sort) and writes k keys per query; merge_keys_kernel
// finishes the job.
#pragma once
#include "topk_common.cuh"

namespace sgic {

constexpr int kScanConsumerWarps = 8;
constexpr int kScanThreads = (kScanConsumerWarps + 1) * 32;  // + producer warp
constexpr int kScanMaxStages = 8;
constexpr int kScanInlineQ = 1024;  // fp32 query elements that can ride in the kernel parameters (one query up to d = 1024)

struct ScanSmallParams {
  const void* db;        // [n_rows][d] 16-bit, row-major
  const float* q;        // [nq][d] fp32 (rounded to the db dtype in-kernel)
  uint64_t* partial;     // [NQ][gridDim.x][k] keys
  uint32_t n_rows;
  uint32_t d;            // multiple of 8
  uint32_t nq;           // valid queries (<= NQ)
  uint32_t k;
  uint32_t kp;           // list stride (power of two >= k)
  uint32_t n_tiles;
  uint32_t n_stages;
  uint32_t stage_bytes;  // R * d * 2 rounded up to 128
  uint32_t evict_first;  // L2 policy for the database stream
  uint32_t wait_ns;      // consumers sleep this long after a failed mbarrier try (0: poll)
  // fused final merge: the last CTA to finish merges every CTA's list and writes the answer
  uint32_t fused;        // 0: partial lists only (merge_keys_kernel follows)
  uint32_t* counter;     // zero on entry, reset to zero by the last CTA
  // Tiles [0, n_static) are dealt round-robin (tile = cta + i * grid: equal shares); the rest is claimed two at a
  // time from *steal (zero on entry; reset by the last CTA when fused, by the host otherwise).  CTAs do not stream
  // at the same pace (per-CTA %globaltimer stamps, 1M rows: first done at 140.6 us, last at 147.6 us; 12.5M rows:
  // 1687 / 1732 us) — with equal shares the slowest sets the time, with a shared tail they finish together.
  uint32_t n_static;
  uint32_t* steal;
  float* D;              // [nq][k]
  long long* I;          // [nq][k]
  long long id_base;
  // k > 1024 runs as several passes of <= 1024: a pass only admits candidates whose key lies strictly below
  // *bound (the last key the previous pass returned; device memory, nullptr / 0 = no bound).  Keys order
  // candidates by (score desc, id asc), so the passes enumerate the answer in order without gaps or repeats.
  const uint64_t* bound;
  // "trace" option: per CTA, %globaltimer at kernel entry / end of the scan / lists written / kernel exit
  // ([gridDim.x][4] u64, nullptr = off); scripts/probe_k3_trace.py turns it into the launch / ramp / tail picture
  unsigned long long* trace;
  // q == nullptr: the (nq * d <= kScanInlineQ) query elements are HERE, in the kernel parameters — a host-buffer
  // search of one or two queries then needs no H2D copy at all (a copy engine costs ~8 us of start-up latency, and
  // 148 CTAs fetching 2 KB each from pinned host memory over PCIe were measured slower still: 32-byte reads)
  float q_inline[kScanInlineQ];
  // Multi-GPU exchange folded into the scan (SURVEY.md §8e "optional fusion: peer-store epilogue + flag"): when x_n > 0
  // the last CTA's merge stores the answer — global row numbers — into x_n destinations (gather slots in the home GPU's
  // or every peer's HBM, mapped over NVLink) instead of (D, I), and then raises one system-scope release flag per
  // destination to x_epoch.  The merge kernel on the other side spins on the flags: no push kernel, no event.
  uint32_t x_n, x_epoch;
  float* xD[16];
  long long* xI[16];
  uint32_t* xF[16];
};

// next pass's bound = key of the last answer slot of this pass (1 = "nothing left" if that slot is padding)
__global__ void next_bound_kernel(const float* D, const long long* I, uint32_t last, long long id_base, uint64_t* bound) {
  const long long id = I[last];
  *bound = id < 0 ? 1ull : make_key(D[last], static_cast<uint32_t>(id - id_base));
}
__global__ void fill_padding_kernel(float* D, long long* I, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    D[i] = kNegFltMax;
    I[i] = -1;
  }
}

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <typename T, int NQ, int CPL, int RB>
__global__ void __launch_bounds__(kScanThreads, 1) scan_small_kernel(const __grid_constant__ ScanSmallParams p) {
  constexpr int W = kScanConsumerWarps;
  constexpr int R = W * RB;       // rows per tile
  constexpr int V = NQ * RB;      // partial sums per lane before the butterfly
  static_assert(V <= 32 && (V & (V - 1)) == 0, "NQ*RB must be a power of two <= 32");
  constexpr int L = (V == 1) ? 0 : (V == 2) ? 1 : (V == 4) ? 2 : (V == 8) ? 3 : (V == 16) ? 4 : 5;

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* stage_base = smem;
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.n_stages) * p.stage_bytes);
  uint64_t* full_bar = lists + static_cast<size_t>(NQ) * W * p.kp;
  uint64_t* empty_bar = full_bar + kScanMaxStages;
  uint32_t* stage_tile = reinterpret_cast<uint32_t*>(empty_bar + kScanMaxStages);  // tile in each stage (producer -> consumers)
  // the queries, staged once per CTA: p.q may point into the peer's HBM (multi-GPU front) or into pinned HOST memory
  // (host-buffer searches skip the H2D copy) — one coalesced read per CTA instead of one per warp
  float* q_s = reinterpret_cast<float*>(stage_tile + kScanMaxStages);
  constexpr uint32_t kEndOfStream = 0xffffffffu;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t row_bytes = p.d * 2u;
  const uint32_t n_chunks = p.d >> 3;  // 16-byte chunks per row
  if (p.trace && tid == 0) p.trace[blockIdx.x * 4 + 0] = global_timer_ns();

  for (uint32_t i = tid; i < static_cast<uint32_t>(NQ) * W * p.kp; i += kScanThreads) lists[i] = 0ull;
  if (tid == 0) {
    for (uint32_t s = 0; s < p.n_stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], W);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if (warp == W) {
    // ------------------------------------------------------------ producer
    if (lane == 0) {
      const uint64_t pol = p.evict_first ? ptx::policy_evict_first() : ptx::policy_evict_last();
      const uint8_t* db = static_cast<const uint8_t*>(p.db);
      uint32_t it = 0;
      auto issue = [&](uint32_t tile) {
        const uint32_t s = it % p.n_stages, use = it / p.n_stages;
        if (use > 0) ptx::mbar_wait(&empty_bar[s], (use - 1) & 1);
        stage_tile[s] = tile;  // ordered before the consumers' wait by the arrive (release) below
        if (tile == kEndOfStream) {
          ptx::mbar_arrive(&full_bar[s]);
        } else {
          const uint32_t row0 = tile * R;
          const uint32_t rows = min(static_cast<uint32_t>(R), p.n_rows - row0);
          const uint32_t bytes = rows * row_bytes;
          ptx::mbar_expect_tx(&full_bar[s], bytes);
          ptx::bulk_g2s(stage_base + static_cast<size_t>(s) * p.stage_bytes,
                        db + static_cast<size_t>(row0) * row_bytes, bytes, &full_bar[s], pol);
        }
        ++it;
      };
      for (uint32_t tile = blockIdx.x; tile < p.n_static; tile += gridDim.x) issue(tile);
      while (true) {  // the shared tail: two tiles per claim (one atomic round trip per 64 KB streamed)
        const uint32_t t0 = p.n_static + atomicAdd(p.steal, 2u);
        if (t0 >= p.n_tiles) break;
        issue(t0);
        if (t0 + 1 < p.n_tiles) issue(t0 + 1);
      }
      issue(kEndOfStream);
    }
  } else {
    // ------------------------------------------------------------ consumers
    // query fragments, rounded to the storage dtype (same rounding the dense path applies)
    {
      const uint32_t n4 = (p.nq * p.d) >> 2;  // d is a multiple of 8
      if (p.q != nullptr) {
        const float4* src = reinterpret_cast<const float4*>(p.q);
        for (uint32_t i = tid; i < n4; i += W * 32) reinterpret_cast<float4*>(q_s)[i] = __ldg(src + i);
      } else {
        const float4* src = reinterpret_cast<const float4*>(p.q_inline);  // constant bank (grid-constant parameter)
        for (uint32_t i = tid; i < n4; i += W * 32) reinterpret_cast<float4*>(q_s)[i] = src[i];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(W * 32) : "memory");  // the consumer warps only: the producer is already streaming
    }
    uint32_t qf[NQ][CPL][4];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const uint32_t chunk = lane + 32 * c;
        if (qi < static_cast<int>(p.nq) && chunk < n_chunks) {
          const float4* src = reinterpret_cast<const float4*>(q_s + static_cast<size_t>(qi) * p.d + chunk * 8);
          const float4 a = src[0], b = src[1];
          qf[qi][c][0] = pack2<T>(a.x, a.y);
          qf[qi][c][1] = pack2<T>(a.z, a.w);
          qf[qi][c][2] = pack2<T>(b.x, b.y);
          qf[qi][c][3] = pack2<T>(b.z, b.w);
        } else {
          qf[qi][c][0] = qf[qi][c][1] = qf[qi][c][2] = qf[qi][c][3] = 0u;
        }
      }
    }
    float thr[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) thr[qi] = -INFINITY;
    const uint64_t bound = p.bound ? *p.bound : 0ull;

    for (uint32_t it = 0;; ++it) {
      const uint32_t s = it % p.n_stages, use = it / p.n_stages;
      ptx::mbar_wait_relaxed(&full_bar[s], use & 1, p.wait_ns);
      const uint32_t tile = stage_tile[s];
      if (tile == kEndOfStream) break;
      const uint32_t row0 = tile * R;
      const uint32_t rows = min(static_cast<uint32_t>(R), p.n_rows - row0);
      const uint8_t* sb = stage_base + static_cast<size_t>(s) * p.stage_bytes;

      uint4 v[RB][CPL];
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const uint32_t r = warp * RB + j;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const uint32_t chunk = lane + 32 * c;
          if (r < rows && chunk < n_chunks)
            v[j][c] = *reinterpret_cast<const uint4*>(sb + r * row_bytes + chunk * 16);
          else
            v[j][c] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      float acc[V];
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) {
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          float a = 0.f;
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            a = ptx::fma_mixed_lo<T>(v[j][c].x, qf[qi][c][0], a);
            a = ptx::fma_mixed_hi<T>(v[j][c].x, qf[qi][c][0], a);
            a = ptx::fma_mixed_lo<T>(v[j][c].y, qf[qi][c][1], a);
            a = ptx::fma_mixed_hi<T>(v[j][c].y, qf[qi][c][1], a);
            a = ptx::fma_mixed_lo<T>(v[j][c].z, qf[qi][c][2], a);
            a = ptx::fma_mixed_hi<T>(v[j][c].z, qf[qi][c][2], a);
            a = ptx::fma_mixed_lo<T>(v[j][c].w, qf[qi][c][3], a);
            a = ptx::fma_mixed_hi<T>(v[j][c].w, qf[qi][c][3], a);
          }
          acc[qi * RB + j] = a;
        }
      }
      // the stage's bytes are in registers now: hand the slot back to the producer
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty_bar[s]);

      // transposing butterfly: V values/lane -> 1 value/lane, then finish the all-reduce
#pragma unroll
      for (int step = 0; step < L; ++step) {
        const int off = 16 >> step;
        const int half = V >> (step + 1);
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const float keep = up ? acc[i + half] : acc[i];
          const float send = up ? acc[i] : acc[i + half];
          acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      float val = acc[0];
#pragma unroll
      for (int off = 16 >> L; off > 0; off >>= 1) val += __shfl_xor_sync(0xffffffffu, val, off);

      // lane holds the score of element e = lane >> (5-L):  query e / RB, row e % RB
      const int e = lane >> (5 - L);
      const int qi_me = e / RB, j_me = e % RB;
      const bool rep = (lane & ((1 << (5 - L)) - 1)) == 0;
      float t = thr[0];
#pragma unroll
      for (int qi = 1; qi < NQ; ++qi)
        if (qi_me == qi) t = thr[qi];
      const bool cand = rep && (warp * RB + j_me < rows) && (qi_me < static_cast<int>(p.nq)) && (val > t);
      uint32_t m = __ballot_sync(0xffffffffu, cand);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float sv = __shfl_sync(0xffffffffu, val, src);
        const int se = src >> (5 - L);
        const int sq = se / RB, sj = se % RB;
        const uint32_t id = row0 + warp * RB + sj;
        const uint64_t key = make_key(sv, id);
        if (bound != 0ull && key >= bound) continue;  // returned by an earlier pass (warp-uniform)
        uint64_t* Lq = lists + (static_cast<size_t>(sq) * W + warp) * p.kp;
        const uint64_t kth = warp_list_insert(Lq, static_cast<int>(p.k), key, lane);
        const float nt = (kth == 0ull) ? -INFINITY : key_score(kth);
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi)
          if (sq == qi) thr[qi] = nt;
      }
    }
  }

  // ---------------------------------------------------------------- CTA merge of the W warp lists
  __syncthreads();
  if (p.trace && tid == 0) p.trace[blockIdx.x * 4 + 1] = global_timer_ns();
  const uint32_t m = W * p.kp;  // power of two
  if (p.k <= 64u) {
    // short lists: warp qi merges the eight sorted warp lists of query qi head by head (k rounds of a 3-step
    // warp max) — ~1 us instead of the 28 barrier-separated stages of the block-wide sort below
    if (warp < static_cast<int>(p.nq)) {
      uint64_t* dst = p.partial + (static_cast<size_t>(warp) * gridDim.x + blockIdx.x) * p.k;
      warp_multiway_merge<1>(lists + static_cast<size_t>(warp) * m, W, p.kp, p.k, lane,
                             [=](uint32_t r, uint64_t key) { dst[r] = key; });
    }
  } else {
    for (uint32_t qi = 0; qi < p.nq; ++qi) {
      uint64_t* Lq = lists + static_cast<size_t>(qi) * m;
      block_bitonic_sort_desc(Lq, m);
      uint64_t* dst = p.partial + (static_cast<size_t>(qi) * gridDim.x + blockIdx.x) * p.k;
      for (uint32_t i = tid; i < p.k; i += kScanThreads) dst[i] = Lq[i];
    }
  }
  if (p.trace) {
    __syncthreads();
    if (tid == 0) p.trace[blockIdx.x * 4 + 2] = p.trace[blockIdx.x * 4 + 3] = global_timer_ns();
  }
  if (!p.fused) return;

  // ---------------------------------------------------------------- fused grid-level merge
  // "last block done": every CTA publishes its lists, bumps a counter; the CTA that sees the final
  // count stages all lists in shared memory (the stage ring is free now) and one warp per query
  // runs the multiway merge.  Saves the second kernel launch on the latency-critical small scans.
  __shared__ uint32_t s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(p.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  uint64_t* all = reinterpret_cast<uint64_t*>(smem);
  const uint32_t per_q = gridDim.x * p.k;
  for (uint32_t i = tid; i < p.nq * per_q; i += kScanThreads) all[i] = __ldcg(p.partial + i);
  __syncthreads();
  if (warp < static_cast<int>(p.nq)) {
    float* D = p.D;
    long long* I = p.I;
    const size_t o0 = static_cast<size_t>(warp) * p.k;
    const long long base = p.id_base;
    if (p.x_n == 0) {
      warp_multiway_merge<kMergeMaxLpl>(all + static_cast<size_t>(warp) * per_q, gridDim.x, p.k, p.k, lane,
                                        [=](uint32_t r, uint64_t key) { store_answer(D, I, o0 + r, key, base); });
    } else {
      const ScanSmallParams* pp = &p;  // grid-constant: lives in the constant bank
      warp_multiway_merge<kMergeMaxLpl>(all + static_cast<size_t>(warp) * per_q, gridDim.x, p.k, p.k, lane,
                                        [=](uint32_t r, uint64_t key) {
                                          for (uint32_t j = 0; j < pp->x_n; ++j)
                                            store_answer(pp->xD[j], pp->xI[j], o0 + r, key, base);
                                        });
      __threadfence_system();  // every lane that stored (lane 0, and the padding loop's lanes) orders its stores
    }
  }
  if (p.x_n != 0) {
    __syncthreads();  // all queries' answers are out (and fenced) before any flag goes up
    if (tid < static_cast<int>(p.x_n)) st_release_sys(p.xF[tid], p.x_epoch);
  }
  if (tid == 0) {
    *p.counter = 0u;
    *p.steal = 0u;
  }
  if (p.trace) {
    __syncthreads();
    if (tid == 0) p.trace[blockIdx.x * 4 + 3] = global_timer_ns();
  }
}

}  // namespace sgic
