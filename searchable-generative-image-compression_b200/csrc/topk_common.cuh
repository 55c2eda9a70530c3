// Shared top-k building blocks: warp-cooperative sorted insertion, block-wide bitonic
// sort of 64-bit candidate keys, and the final merge kernels (per-CTA partial lists
// -> answer; per-shard answers -> global answer = K5 in SURVEY.md §2.1).
#pragma once
#include "ptx.cuh"

namespace sgic {

constexpr float kNegFltMax = -3.4028234663852886e38f;  // faiss pads missing slots with -FLT_MAX / -1

__host__ __device__ __forceinline__ uint32_t next_pow2_u32(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Sort `n` (power of two) keys in shared memory, largest first.  All threads of the
// block call it; ends with a __syncthreads().
__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t* keys, uint32_t n) {
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  for (uint32_t size = 2; size <= n; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = tid; i < (n >> 1); i += nt) {
        const uint32_t lo = 2 * i - (i & (stride - 1));
        const uint32_t hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t a = keys[lo], b = keys[hi];
        if ((a < b) == desc) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
}

// Insert `key` into the warp-private list L[0..k) (sorted, largest first, empty slots
// are 0).  All 32 lanes call it with the same arguments.  Returns the list's new k-th
// key (0 while the list is not full).
__device__ __forceinline__ uint64_t warp_list_insert(uint64_t* L, int k, uint64_t key, int lane) {
  int p = 0;
  for (int base = 0; base < k; base += 32) {
    const int i = base + lane;
    const bool g = (i < k) && (L[i] > key);
    const uint32_t m = __ballot_sync(0xffffffffu, g);
    p += __popc(m);
    if (m != 0xffffffffu) break;  // sorted: first chunk that is not entirely greater ends the scan
  }
  if (p < k) {
    for (int base = ((k - 1) >> 5) << 5; base >= 0 && base + 31 > p; base -= 32) {
      const int i = base + lane;
      const bool mv = (i > p) && (i < k);
      uint64_t t = 0;
      if (mv) t = L[i - 1];
      __syncwarp();
      if (mv) L[i] = t;
      __syncwarp();
    }
    if (lane == 0) L[p] = key;
    __syncwarp();
  }
  return L[k - 1];
}

// Writes one answer slot from a key (0 = empty slot -> faiss padding).
__device__ __forceinline__ void store_answer(float* D, long long* I, size_t o, uint64_t key, long long id_base) {
  if (key == 0ull) {
    D[o] = kNegFltMax;
    I[o] = -1;
  } else {
    D[o] = key_score(key);
    I[o] = id_base + static_cast<long long>(key_id(key));
  }
}

// One warp merges `n_lists` (<= 32*MAXLPL) lists, each sorted largest-first and `len` keys long
// (list l at lists[l*len]), into the `k_out` largest keys overall, handed to emit(rank, key) by
// lane 0 in rank order.  A multiway merge over the list heads: per output one register max per
// lane, a 5-step warp max, and one list advance by the owning lane — k_out short rounds instead
// of sorting n_lists*len candidates.  Keys are unique (the id is part of the key).
template <int MAXLPL, typename Emit>
__device__ __forceinline__ void warp_multiway_merge(const uint64_t* lists, uint32_t n_lists, uint32_t len,
                                                    uint32_t k_out, int lane, Emit emit) {
  uint64_t head[MAXLPL];
  uint32_t pos[MAXLPL];
#pragma unroll
  for (int j = 0; j < MAXLPL; ++j) {
    const uint32_t l = lane + 32 * j;
    head[j] = (l < n_lists && len > 0) ? lists[static_cast<size_t>(l) * len] : 0ull;
    pos[j] = 0;
  }
  for (uint32_t r = 0; r < k_out; ++r) {
    uint64_t best = head[0];
    int bj = 0;
#pragma unroll
    for (int j = 1; j < MAXLPL; ++j)
      if (head[j] > best) {
        best = head[j];
        bj = j;
      }
    // warp max of a 64-bit key with two REDUX.MAX (score word, then the id word among the lanes that hold the
    // best score) instead of a 5-step shuffle butterfly: the rounds are a latency chain, this halves it.
    // A real key never has a zero id word (ids stop at 2^32-2), so 0 stands for "not a contender".
    const uint32_t bh = static_cast<uint32_t>(best >> 32);
    const uint32_t mh = __reduce_max_sync(0xffffffffu, bh);
    const uint32_t ml = __reduce_max_sync(0xffffffffu, bh == mh ? static_cast<uint32_t>(best) : 0u);
    const uint64_t w = (static_cast<uint64_t>(mh) << 32) | ml;
    if (lane == 0) emit(r, w);
    if (w == 0ull) {  // every list exhausted: the rest is padding
      for (uint32_t r2 = r + 1 + lane; r2 < k_out; r2 += 32) emit(r2, 0ull);
      break;
    }
    if (best == w) {
#pragma unroll
      for (int j = 0; j < MAXLPL; ++j)
        if (bj == j) {
          const uint32_t l = lane + 32 * j;
          ++pos[j];
          head[j] = (pos[j] < len) ? lists[static_cast<size_t>(l) * len + pos[j]] : 0ull;
        }
    }
  }
}
constexpr int kMergeMaxLpl = 5;  // lists per lane: up to 160 lists (one per SM of a 148-SM grid)

// ------------------------------------------------------------------ final merge (per-CTA partials)
// partial: [nq][n_lists][k] keys (score, local id).  One CTA per query.  Keys are pulled
// through shared memory in chunks of at most `chunk` (power of two) keys; each chunk is
// sorted together with the best k carried from the previous one.
struct MergeKeysParams {
  const uint64_t* partial;
  uint32_t n_lists;
  uint32_t k;
  uint32_t chunk;   // smem capacity in keys (power of two, >= 2*k_pow2)
  float* D;         // [nq][k]
  long long* I;     // [nq][k]
  long long id_base;
};

__global__ void __launch_bounds__(1024, 1) merge_keys_kernel(MergeKeysParams p) {
  extern __shared__ __align__(16) uint64_t mk_smem[];
  const uint32_t q = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const uint64_t* src = p.partial + static_cast<size_t>(q) * p.n_lists * p.k;
  const uint32_t n_total = p.n_lists * p.k;
  uint32_t carry = 0, pos = 0;
  while (pos < n_total) {
    const uint32_t take = min(p.chunk - carry, n_total - pos);
    const uint32_t m = max(next_pow2_u32(carry + take), 2u);
    for (uint32_t i = tid; i < m - carry; i += nt) mk_smem[carry + i] = (i < take) ? src[pos + i] : 0ull;
    __syncthreads();
    block_bitonic_sort_desc(mk_smem, m);
    carry = min(p.k, m);
    pos += take;
  }
  for (uint32_t i = tid; i < p.k; i += nt) {
    const uint64_t key = (i < carry) ? mk_smem[i] : 0ull;
    const size_t o = static_cast<size_t>(q) * p.k + i;
    if (key == 0ull) {
      p.D[o] = kNegFltMax;
      p.I[o] = -1;
    } else {
      p.D[o] = key_score(key);
      p.I[o] = p.id_base + static_cast<long long>(key_id(key));
    }
  }
}

// Fast path of the above when n_lists <= 160 and all lists fit in shared memory: stage the lists
// with one coalesced pass, then one warp runs the multiway merge (k short rounds).
__global__ void __launch_bounds__(128, 1) merge_keys_small_kernel(MergeKeysParams p) {
  extern __shared__ __align__(16) uint64_t ms_smem[];
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  const uint64_t* src = p.partial + static_cast<size_t>(q) * p.n_lists * p.k;
  const uint32_t n_total = p.n_lists * p.k;
  for (uint32_t i = tid; i < n_total; i += blockDim.x) ms_smem[i] = __ldcg(src + i);
  __syncthreads();
  if (tid < 32) {
    float* D = p.D;
    long long* I = p.I;
    const size_t o0 = static_cast<size_t>(q) * p.k;
    const long long base = p.id_base;
    warp_multiway_merge<kMergeMaxLpl>(ms_smem, p.n_lists, p.k, p.k, static_cast<int>(tid),
                                      [=](uint32_t r, uint64_t key) { store_answer(D, I, o0 + r, key, base); });
  }
}

// ------------------------------------------------------------------ K5: merge of per-shard answers
// D_lists / I_lists: [n_lists][nq][k] as an all-gather of each rank's (nq,k) answer lays
// them out.  Ordering is (score desc, global id asc) so that the G-GPU answer is identical to
// the 1-GPU answer.  tie_by_position == 0: ids < 2^32, the id itself is the low key word.
// tie_by_position == 1 (ids may exceed 32 bits): shards must hold ascending, disjoint row
// ranges; each list is already ordered (score desc, id asc), so the position g*k+j orders ties
// exactly like the global id does.
struct MergeListsParams {
  const float* D_lists;
  const long long* I_lists;
  uint32_t n_lists;
  uint32_t nq;
  uint32_t k;
  uint32_t chunk;
  uint32_t tie_by_position;
  float* D;
  long long* I;
  size_t list_stride = 0;            // elements between two lists (0: nq * k, the all-gather layout)
  // peer exchange (K5x): list g is complete once wait_flags[g * kXchgFlagStride] >= wait_epoch (written by rank g
  // over NVLink with a system-scope release after its candidates); nullptr: the lists are already there
  const uint32_t* wait_flags = nullptr;
  uint32_t wait_epoch = 0;
  uint32_t* err = nullptr;           // set to 1 if a flag did not arrive within kXchgTimeoutNs
};

constexpr uint32_t kXchgFlagStride = 32;                     // one flag per 128-byte line
constexpr unsigned long long kXchgTimeoutNs = 20000000000ull;  // a peer that is 20 s late is gone: do not hang the GPU

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// K5x push: this rank's (nq, k) answer -> slot `rank` of EVERY rank's gather buffer (its own included), written
// straight into the peers' HBM over NVLink, then one release-flag per destination.  One CTA per destination.
struct XchgPushParams {
  const float* D;          // [n] this rank's scores (n = nq * k)
  const long long* I;      // [n] global row numbers
  uint32_t n;
  uint32_t epoch;
  float* dstD[16];         // per destination rank: slot `rank` of its D_lists / I_lists / flags (peer mappings)
  long long* dstI[16];
  uint32_t* dstFlag[16];
};

__global__ void __launch_bounds__(256, 1) xchg_push_kernel(const XchgPushParams p) {
  const uint32_t dst = blockIdx.x;
  float* D = p.dstD[dst];
  long long* I = p.dstI[dst];
  for (uint32_t i = threadIdx.x; i < p.n; i += blockDim.x) {
    D[i] = p.D[i];
    I[i] = p.I[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(p.dstFlag[dst], p.epoch);
}

__global__ void __launch_bounds__(1024, 1) merge_lists_kernel(MergeListsParams p) {
  extern __shared__ __align__(16) uint64_t ml_smem[];
  const uint32_t q = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const uint32_t n_total = p.n_lists * p.k;
  const size_t stride = p.list_stride ? p.list_stride : static_cast<size_t>(p.nq) * p.k;
  if (p.wait_flags != nullptr) {
    if (tid < p.n_lists) {
      const uint32_t* f = p.wait_flags + static_cast<size_t>(tid) * kXchgFlagStride;
      const unsigned long long t0 = global_timer_ns();
      while (static_cast<int32_t>(ld_acquire_sys(f) - p.wait_epoch) < 0) {
        if (global_timer_ns() - t0 > kXchgTimeoutNs) {
          if (p.err) *p.err = 1u;
          break;
        }
        __nanosleep(100);
      }
    }
    __syncthreads();
  }
  if (n_total <= p.chunk && p.n_lists <= 32u * kMergeMaxLpl) {
    // fast path: all candidates fit in shared memory -> stage the keys, one warp merges the heads
    for (uint32_t e = tid; e < n_total; e += nt) {
      const uint32_t g = e / p.k, j = e - g * p.k;
      const size_t o = static_cast<size_t>(g) * stride + static_cast<size_t>(q) * p.k + j;
      const long long id = __ldcg(p.I_lists + o);
      ml_smem[e] = (id >= 0) ? make_key(__ldcg(p.D_lists + o), p.tie_by_position ? e : static_cast<uint32_t>(id)) : 0ull;
    }
    __syncthreads();
    if (tid < 32) {
      const MergeListsParams pp = p;
      warp_multiway_merge<kMergeMaxLpl>(ml_smem, p.n_lists, p.k, p.k, static_cast<int>(tid),
                                        [=](uint32_t r, uint64_t key) {
                                          const size_t out = static_cast<size_t>(q) * pp.k + r;
                                          if (key == 0ull) {
                                            pp.D[out] = kNegFltMax;
                                            pp.I[out] = -1;
                                          } else if (pp.tie_by_position) {
                                            const uint32_t e = key_id(key), g = e / pp.k, j = e - g * pp.k;
                                            const size_t o = static_cast<size_t>(g) * stride + static_cast<size_t>(q) * pp.k + j;
                                            pp.D[out] = __ldcg(pp.D_lists + o);
                                            pp.I[out] = __ldcg(pp.I_lists + o);
                                          } else {
                                            pp.D[out] = key_score(key);
                                            pp.I[out] = static_cast<long long>(key_id(key));
                                          }
                                        });
    }
    return;
  }
  uint32_t carry = 0, pos = 0;
  while (pos < n_total) {
    const uint32_t take = min(p.chunk - carry, n_total - pos);
    const uint32_t m = max(next_pow2_u32(carry + take), 2u);
    for (uint32_t i = tid; i < m - carry; i += nt) {
      uint64_t key = 0ull;
      if (i < take) {
        const uint32_t e = pos + i, g = e / p.k, j = e - g * p.k;
        const size_t o = static_cast<size_t>(g) * stride + static_cast<size_t>(q) * p.k + j;
        const long long id = __ldcg(p.I_lists + o);
        if (id >= 0) key = make_key(__ldcg(p.D_lists + o), p.tie_by_position ? e : static_cast<uint32_t>(id));
      }
      ml_smem[carry + i] = key;
    }
    __syncthreads();
    block_bitonic_sort_desc(ml_smem, m);
    carry = min(p.k, m);
    pos += take;
  }
  for (uint32_t i = tid; i < p.k; i += nt) {
    const uint64_t key = (i < carry) ? ml_smem[i] : 0ull;
    const size_t out = static_cast<size_t>(q) * p.k + i;
    if (key == 0ull) {
      p.D[out] = kNegFltMax;
      p.I[out] = -1;
    } else if (p.tie_by_position) {
      const uint32_t e = key_id(key), g = e / p.k, j = e - g * p.k;
      const size_t o = static_cast<size_t>(g) * stride + static_cast<size_t>(q) * p.k + j;
      p.D[out] = __ldcg(p.D_lists + o);
      p.I[out] = __ldcg(p.I_lists + o);
    } else {
      p.D[out] = key_score(key);
      p.I[out] = static_cast<long long>(key_id(key));
    }
  }
}

}  // namespace sgic
