// Micro-benchmark: HBM streaming rate of TMA tensor loads by box shape (which request pattern should the dense
// kernel's HBM regime use?).  Row-major [N][512] fp16 database, 148 persistent CTAs, ring of stages, no compute.
//   mode 0: 2-D box {64 elems, R rows}, one K chunk per stage (what scan_dense_kernel does today)
//   mode 1: 3-D view (64, rows, kc) box {64, R, 8}: R full rows per instruction, rows-major traversal, smem layout
//           = 8 K-chunk tiles of [R][128 B] (directly usable by UMMA)
//   mode 2: 3-D view (64, kc, rows) box {64, 8, R}: contiguous 1 KB per row (upper bound; layout not UMMA-usable)
//   mode 3: 1-D bulk copy of R contiguous rows (K3's pattern)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/tma_stream_bench scripts/micro/tma_stream_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void tma2(void* dst, const void* tm, int c0, int c1, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
               ::"r"(s32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(s32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma3(void* dst, const void* tm, int c0, int c1, int c2, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
               ::"r"(s32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk1(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)), "l"(pol) : "memory");
}

struct P { const uint8_t* db; uint32_t n_rows, R, stages, mode, kc_per_stage; uint32_t stage_bytes; unsigned long long* sink; };

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* st = smem + ((1024u - (s32(smem) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(st + (size_t)p.stages * p.stage_bytes);
  uint64_t* empty = full + p.stages;
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t n_blocks = p.n_rows / p.R;   // row blocks of R rows
  const uint32_t kc_total = 8;                // d = 512
  const uint32_t steps_per_block = (p.mode == 0) ? kc_total / p.kc_per_stage : 1;
  if (threadIdx.x == 0) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    uint32_t it = 0;
    for (uint32_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
      for (uint32_t stp = 0; stp < steps_per_block; ++stp, ++it) {
        const uint32_t s = it % p.stages, u = it / p.stages;
        if (u > 0) mbar_wait(&empty[s], (u - 1) & 1);
        uint8_t* dst = st + (size_t)s * p.stage_bytes;
        mbar_expect(&full[s], p.stage_bytes);
        if (p.mode == 0) {
          for (uint32_t j = 0; j < p.kc_per_stage; ++j)
            tma2(dst + (size_t)j * p.R * 128, &tm, (int)((stp * p.kc_per_stage + j) * 64), (int)(b * p.R), &full[s], pol);
        } else if (p.mode == 1) {
          tma3(dst, &tm, 0, (int)(b * p.R), 0, &full[s], pol);
        } else if (p.mode == 2) {
          tma3(dst, &tm, 0, 0, (int)(b * p.R), &full[s], pol);
        } else {
          bulk1(dst, p.db + (size_t)b * p.R * 1024, p.R * 1024, &full[s], pol);
        }
      }
    }
  } else if (threadIdx.x == 32) {
    uint32_t it = 0;
    unsigned long long acc = 0;
    for (uint32_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
      for (uint32_t stp = 0; stp < steps_per_block; ++stp, ++it) {
        const uint32_t s = it % p.stages;
        mbar_wait(&full[s], (it / p.stages) & 1);
        acc += *reinterpret_cast<volatile uint32_t*>(st + (size_t)s * p.stage_bytes);
        mbar_arrive(&empty[s]);
      }
    }
    if (acc == 0x1234567u) *p.sink = acc;
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const uint32_t n_rows = argc > 1 ? atoi(argv[1]) : 10000000;
  const uint32_t d = 512;
  void* db;
  CK(cudaMalloc(&db, (size_t)n_rows * d * 2));
  CK(cudaMemset(db, 1, (size_t)n_rows * d * 2));
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
  EncFn enc = (EncFn)fp;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  struct Cfg { int mode; uint32_t R, stages, kcps; const char* name; };
  std::vector<Cfg> cfgs = {
      {0, 256, 4, 1, "2D {64,256} x1 per stage, 4 stages of 32 KB (current)"},
      {0, 256, 6, 1, "2D {64,256} x1 per stage, 6 stages of 32 KB"},
      {0, 128, 8, 1, "2D {64,128} x1 per stage, 8 stages of 16 KB"},
      {0, 64, 3, 8, "2D {64,64} x8 per stage, 3 stages of 64 KB (full rows per stage)"},
      {0, 64, 2, 8, "2D {64,64} x8 per stage, 2 stages of 64 KB"},
      {0, 32, 4, 8, "2D {64,32} x8 per stage, 4 stages of 32 KB"},
      {0, 32, 6, 8, "2D {64,32} x8 per stage, 6 stages of 32 KB"},
      {1, 64, 3, 8, "3D (64,rows,kc) {64,64,8}, 3 stages of 64 KB"},
      {1, 32, 4, 8, "3D (64,rows,kc) {64,32,8}, 4 stages of 32 KB"},
      {1, 32, 6, 8, "3D (64,rows,kc) {64,32,8}, 6 stages of 32 KB"},
      {2, 64, 3, 8, "3D (64,kc,rows) {64,8,64}, 3 stages of 64 KB (contiguous)"},
      {2, 32, 4, 8, "3D (64,kc,rows) {64,8,32}, 4 stages of 32 KB (contiguous)"},
      {3, 32, 4, 8, "1D bulk 32 rows, 4 stages of 32 KB (K3)"},
      {3, 64, 3, 8, "1D bulk 64 rows, 3 stages of 64 KB"},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap tm;
    CUresult r;
    if (c.mode == 0) {
      cuuint64_t gd[2] = {d, n_rows};
      cuuint64_t gs[1] = {(cuuint64_t)d * 2};
      cuuint32_t box[2] = {64, c.R};
      cuuint32_t es[2] = {1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (c.mode == 1) {
      cuuint64_t gd[3] = {64, n_rows, 8};
      cuuint64_t gs[2] = {(cuuint64_t)d * 2, 128};
      cuuint32_t box[3] = {64, c.R, 8};
      cuuint32_t es[3] = {1, 1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, db, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t gd[3] = {64, 8, n_rows};
      cuuint64_t gs[2] = {128, (cuuint64_t)d * 2};
      cuuint32_t box[3] = {64, 8, c.R};
      cuuint32_t es[3] = {1, 1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, db, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("%-70s  encode failed (%d)\n", c.name, (int)r); continue; }
    P p;
    p.db = (const uint8_t*)db; p.n_rows = n_rows; p.R = c.R; p.stages = c.stages; p.mode = c.mode; p.kc_per_stage = c.kcps;
    p.stage_bytes = c.R * 128 * c.kcps; p.sink = sink;
    const size_t smem = 1024 + (size_t)c.stages * p.stage_bytes + 2 * c.stages * 8 + 64;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f, sum = 0;
    for (int rep = 0; rep < 7; ++rep) {
      CK(cudaEventRecord(e0));
      stream_kernel<<<148, 64, smem>>>(tm, p);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep >= 2) { best = ms < best ? ms : best; sum += ms; }
    }
    const double bytes = (double)(n_rows / c.R) * c.R * d * 2;
    printf("%-70s  best %.3f ms %7.0f GB/s   mean %7.0f GB/s\n", c.name, best, bytes / best / 1e6, bytes / (sum / 5) / 1e6);
  }
  return 0;
}
