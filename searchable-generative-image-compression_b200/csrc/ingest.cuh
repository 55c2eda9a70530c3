// K1 / K2 — device-side loaders that pack embeddings into the row-major 16-bit database.
//   K2 pack_f32_kernel : fp32 rows -> fp16/bf16 rows (IndexFlatIP.add; reference
//                        src/build.py:94, src/compress.py:107 — no renormalisation).
//   K1 dequant_u8_kernel: u8 rows -> ((q/255)*2-1) -> L2-normalise -> fp16/bf16 rows
//                        (reference src/search.py:20-22 dequantize_clip_u8 + :16-18 l2n;
//                        same fp32 operation order, eps = 1e-9).
//   unpack_rows_kernel : 16-bit rows -> fp32 (write_index / reconstruct).
// All three are HBM/PCIe-bound byte movers: 128-bit accesses, grid-stride.
//   K0 zstd_lit_decode_kernel: clip_stream zstd frames -> u8 rows on the device (replaces the
//                        ZstdDecompressor().decompress of src/search.py:35 for frames in the profile of
//                        zstd_lit.cuh), feeding K1.
#pragma once
#include "ptx.cuh"
#include "zstd_lit.cuh"

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

// ---------------------------------------------------------------------------------------------- K0
constexpr uint32_t kZlMaxFrame = 2560;      // frame bytes staged per warp (d <= 2048 raw worst case + headers)
constexpr int kZlWarpsPerBlock = 4;

struct __align__(16) ZlWarpSmem {
  uint8_t frame[kZlMaxFrame];
  uint8_t lit[2048];   // literals of a frame that also has sequences
  uint8_t out[2048];   // its regenerated row, before the cooperative copy to global memory
  uint16_t tab[1u << zl::kHufMaxLog];
  uint8_t w[256];
  uint8_t nb[256];
  zl::FseTable ft;
  zl::SeqTables seq;
  uint32_t info[16];
};

// One warp per frame.  The frame is staged in shared memory with 16-byte loads; lane 0 parses the headers and
// decodes the Huffman tree description (FSE-compressed weights: inherently serial, ~100 symbols); all lanes
// fill the 2^max_bits decoding table; lanes 0..3 decode the four backward bitstreams in parallel — straight
// into the u8 row when the block has no sequences, otherwise into shared memory, where lane 0 then decodes and
// executes the (few) match sequences and the warp copies the finished row out.  status[f] = zl::ZL_OK or the
// reason the frame has to go back to libzstd on the host.
__global__ void __launch_bounds__(kZlWarpsPerBlock * 32)
zstd_lit_decode_kernel(const uint8_t* __restrict__ frames, const ZlDesc* __restrict__ desc, uint32_t n_frames,
                       uint32_t d, uint8_t* __restrict__ rows, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t zl_smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ZlWarpSmem& S = reinterpret_cast<ZlWarpSmem*>(zl_smem_raw)[warp];
  const uint32_t n_warps = gridDim.x * kZlWarpsPerBlock;
  for (uint32_t f = blockIdx.x * kZlWarpsPerBlock + warp; f < n_frames; f += n_warps) {
    const ZlDesc de = desc[f];
    if (de.len > kZlMaxFrame || d > 2048u) {  // the host never sends one; defensive
      if (lane == 0) status[f] = zl::ZL_HOST;
      continue;
    }
    for (uint32_t i = lane * 16; i < de.len; i += 32 * 16)
      *reinterpret_cast<uint4*>(S.frame + i) = __ldg(reinterpret_cast<const uint4*>(frames + de.off + i));
    __syncwarp();
    if (lane == 0) {
      zl::FrameInfo fi;
      int rc = zl::parse_frame(S.frame, de.len, fi);
      if (rc == zl::ZL_OK && fi.content_size != d) rc = zl::ZL_HOST;
      uint32_t td = 0, nsym = 0, max_bits = 0, kind = 0, pay = 0, has_seq = 0, regen = d;
      if (rc == zl::ZL_OK) {
        if (fi.block_type != 2) {
          kind = fi.block_type;  // 0 raw, 1 RLE block
          pay = fi.block_off;
        } else {
          kind = fi.lit_type;    // literals: 0 raw, 1 RLE, 2 Huffman
          pay = fi.lit_off;
          regen = fi.regen;
          has_seq = S.frame[fi.seq_off] != 0 ? 1u : 0u;
          if (kind == 2) {
            td = zl::huf_read_lengths(S.frame + fi.lit_off, fi.comp, S.w, S.nb, nsym, max_bits, S.ft);
            if (td == 0) rc = zl::ZL_CORRUPT;
          }
        }
      }
      S.info[0] = static_cast<uint32_t>(rc);
      S.info[1] = kind;
      S.info[2] = pay;
      S.info[3] = fi.comp;
      S.info[4] = fi.n_streams;
      S.info[5] = td;
      S.info[6] = nsym;
      S.info[7] = max_bits;
      S.info[8] = has_seq;
      S.info[9] = regen;
      S.info[10] = fi.seq_off;
      S.info[11] = fi.seq_size;
    }
    __syncwarp();
    const int rc = static_cast<int>(S.info[0]);
    if (rc != zl::ZL_OK) {
      if (lane == 0) status[f] = rc;
      __syncwarp();
      continue;
    }
    uint8_t* dst = rows + static_cast<size_t>(de.row) * d;
    const uint32_t kind = S.info[1], pay = S.info[2], has_seq = S.info[8], regen = S.info[9];
    uint8_t* lit = has_seq ? S.lit : dst;  // without sequences the literals are the row
    bool ok = true;
    if (kind == 0) {
      for (uint32_t i = lane; i < regen; i += 32) lit[i] = S.frame[pay + i];
    } else if (kind == 1) {
      const uint8_t b = S.frame[pay];
      for (uint32_t i = lane; i < regen; i += 32) lit[i] = b;
    } else {
      const uint32_t comp = S.info[3], n_streams = S.info[4], td = S.info[5], nsym = S.info[6], max_bits = S.info[7];
      ok = zl::huf_fill_table(S.nb, nsym, max_bits, S.tab, lane, 32);
      __syncwarp();
      uint32_t soff[4], slen[4], scnt[4];
      const uint8_t* ss = S.frame + pay + td;
      ok = ok && zl::huf_stream_layout(ss, comp - td, n_streams, regen, soff, slen, scnt);
      if (ok && lane < n_streams) {
        uint32_t o = 0;
        for (uint32_t i = 0; i < lane; ++i) o += scnt[i];
        ok = zl::huf_decode_stream(S.tab, max_bits, ss + soff[lane], slen[lane], lit + o, scnt[lane]);
      }
    }
    bool all_ok = __all_sync(0xffffffffu, ok);
    int final_rc = all_ok ? zl::ZL_OK : zl::ZL_CORRUPT;
    if (all_ok && has_seq) {
      __syncwarp();  // the literals are complete in shared memory
      int src = zl::ZL_OK;
      if (lane == 0) src = zl::decode_sequences(S.frame + S.info[10], S.info[11], S.lit, regen, S.out, d, S.seq);
      final_rc = __shfl_sync(0xffffffffu, src, 0);
      __syncwarp();
      if (final_rc == zl::ZL_OK)
        for (uint32_t i = lane; i < d / 8; i += 32)
          reinterpret_cast<uint2*>(dst)[i] = reinterpret_cast<const uint2*>(S.out)[i];
    }
    if (lane == 0) status[f] = final_rc;
    __syncwarp();  // the next frame overwrites this warp's shared memory
  }
}

}  // namespace sgic
