// K1 / K2 — device-side loaders that pack embeddings into the row-major 16-bit database.
//   K2 pack_f32_kernel : fp32 rows -> fp16/bf16 rows (IndexFlatIP.add; reference
//                        src/build.py:94, src/compress.py:107 — no renormalisation).
//   K1 dequant_u8_kernel: u8 rows -> ((q/255)*2-1) -> L2-normalise -> fp16/bf16 rows
//                        (reference src/search.py:20-22 dequantize_clip_u8 + :16-18 l2n;
//                        same fp32 operation order, eps = 1e-9).
//   unpack_rows_kernel : 16-bit rows -> fp32 (write_index / reconstruct).
// All three are HBM/PCIe-bound byte movers: 128-bit accesses, grid-stride.
#pragma once
#include "ptx.cuh"

namespace sgic {

template <typename T>
__device__ __forceinline__ uint32_t pack2_rn(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2_rn<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <>
__device__ __forceinline__ uint32_t pack2_rn<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <typename T>
__device__ __forceinline__ float2 unpack2(uint32_t v);
template <>
__device__ __forceinline__ float2 unpack2<__half>(uint32_t v) {
  return __half22float2(*reinterpret_cast<__half2*>(&v));
}
template <>
__device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t v) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}

// n8 = number of 8-element groups (n_rows * d / 8)
template <typename T>
__global__ void __launch_bounds__(256) pack_f32_kernel(const float* __restrict__ src, void* __restrict__ dst,
                                                       size_t n8) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 o;
    o.x = pack2_rn<T>(a.x, a.y);
    o.y = pack2_rn<T>(a.z, a.w);
    o.z = pack2_rn<T>(b.x, b.y);
    o.w = pack2_rn<T>(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) unpack_rows_kernel(const void* __restrict__ src, float* __restrict__ dst,
                                                          size_t n8) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
    const float2 a = unpack2<T>(v.x), b = unpack2<T>(v.y), c = unpack2<T>(v.z), e = unpack2<T>(v.w);
    reinterpret_cast<float4*>(dst)[2 * i] = make_float4(a.x, a.y, b.x, b.y);
    reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(c.x, c.y, e.x, e.y);
  }
}

// One warp per row.  d multiple of 8.  Each lane handles 8-byte groups of u8.
template <typename T>
__global__ void __launch_bounds__(256) dequant_u8_kernel(const uint8_t* __restrict__ src, void* __restrict__ dst,
                                                         uint32_t n_rows, uint32_t d) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t n8 = d >> 3;
  for (uint32_t row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < n_rows;
       row += gridDim.x * warps_per_block) {
    const uint2* s = reinterpret_cast<const uint2*>(src + static_cast<size_t>(row) * d);
    uint4* o = reinterpret_cast<uint4*>(static_cast<uint8_t*>(dst) + static_cast<size_t>(row) * d * 2);
    float ss = 0.f;
    for (uint32_t g = lane; g < n8; g += 32) {
      const uint2 w = __ldg(s + g);
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const uint32_t byte = ((b < 4 ? w.x : w.y) >> (8 * (b & 3))) & 0xffu;
        const float z = __fsub_rn(__fmul_rn(__fdiv_rn(static_cast<float>(byte), 255.0f), 2.0f), 1.0f);
        ss = __fmaf_rn(z, z, ss);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float nrm = fmaxf(__fsqrt_rn(ss), 1e-9f);
    for (uint32_t g = lane; g < n8; g += 32) {
      const uint2 w = __ldg(s + g);
      float z[8];
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const uint32_t byte = ((b < 4 ? w.x : w.y) >> (8 * (b & 3))) & 0xffu;
        z[b] = __fdiv_rn(__fsub_rn(__fmul_rn(__fdiv_rn(static_cast<float>(byte), 255.0f), 2.0f), 1.0f), nrm);
      }
      uint4 v;
      v.x = pack2_rn<T>(z[0], z[1]);
      v.y = pack2_rn<T>(z[2], z[3]);
      v.z = pack2_rn<T>(z[4], z[5]);
      v.w = pack2_rn<T>(z[6], z[7]);
      o[g] = v;
    }
  }
}

}  // namespace sgic
