// Thin inline-PTX wrappers for sm_100a: mbarrier, bulk async copy (TMA), tcgen05.
// Hand-written; bit layouts cross-checked against the PTX ISA 8.8 text and the
// CuTe headers vendored in the image (cute/arch/{copy_sm90_desc,mma_sm100_desc}.hpp).
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include "tmem_ld.cuh"

namespace sgic {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// one lane of the (converged) warp: elect.sync.  The MMA issuers run their loops on the whole warp with
// warp-uniform operands (descriptors, ring slots, TMEM addresses then live in uniform registers) and issue under
// this predicate — a loop run by `lane == 0` alone makes the compiler shuttle every operand from vector to uniform
// registers with an ELECT / R2UR.BROADCAST / BRA.U.ANY sequence in front of every UTCHMMA.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// The same for warps that have slack (consumers of an HBM-bound stream, epilogue warps behind the tensor pipe):
// a failed try sleeps `ns` nanoseconds instead of polling again at once.  On a part that sits at its power cap in
// every regime the issue slots a spinning warp burns are clock taken from everybody else.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns) {
  while (!mbar_try_wait(bar, parity)) {
    if (ns) __nanosleep(ns);
  }
}

// ---------------------------------------------------------------- L2 policies
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---------------------------------------------------------------- TMA (bulk async copy)
// 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// 2-D tiled tensor-map copy global -> shared (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int32_t c0, int32_t c1, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---------------------------------------------------------------- tcgen05 (5th-gen tensor cores, TMEM)
// One full warp allocates `ncols` (power of two >= 32) TMEM columns; the base address lands in smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed (SASS UTCBAR).
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, M x N x 16 (16-bit inputs), fp32 accumulate (SASS UTCHMMA).
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128-byte-swizzled operand tile in shared memory (rows of 64 16-bit elements = 128 B,
// 8-row swizzle atoms of 1024 B; exactly what a TMA box {64, rows} with SWIZZLE_128B writes).
// Field layout per cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor): start>>4 [0,14),
// LBO>>4 [16,30) (=1 for swizzled K-major), SBO>>4 [32,46) (=1024 B between 8-row groups),
// version=1 [46,48), layout_type=SWIZZLE_128B(2) [61,64).  Advancing K by 16 elements inside the
// 128 B row = +32 B on the start address.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// UMMA::InstrDescriptor for kind::f16: c_format F32 (1) [4,6), a/b format [7,10)/[10,13) (0 = F16,
// 1 = BF16), a/b K-major (0) [15],[16], N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N, uint32_t is_bf16) {
  return (1u << 4) | (is_bf16 << 7) | (is_bf16 << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---------------------------------------------------------------- mixed-precision FMA (sm_100+)
// d = a*b + c with a,b 16-bit and c,d fp32, single rounding (SASS: FHFMA).  The
// product of two fp16/bf16 values is exact in fp32, so this equals an fp32 FMA on
// the up-cast operands.
template <typename T>
__device__ __forceinline__ float fma_mixed_lo(uint32_t a2, uint32_t b2, float c);
template <typename T>
__device__ __forceinline__ float fma_mixed_hi(uint32_t a2, uint32_t b2, float c);

template <>
__device__ __forceinline__ float fma_mixed_lo<__half>(uint32_t a2, uint32_t b2, float c) {
  float d;
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
      "mov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
      "fma.rn.f32.f16 %0, al, bl, %3;\n\t}"
      : "=f"(d)
      : "r"(a2), "r"(b2), "f"(c));
  return d;
}
template <>
__device__ __forceinline__ float fma_mixed_hi<__half>(uint32_t a2, uint32_t b2, float c) {
  float d;
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
      "mov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
      "fma.rn.f32.f16 %0, ah, bh, %3;\n\t}"
      : "=f"(d)
      : "r"(a2), "r"(b2), "f"(c));
  return d;
}
template <>
__device__ __forceinline__ float fma_mixed_lo<__nv_bfloat16>(uint32_t a2, uint32_t b2, float c) {
  float d;
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
      "mov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
      "fma.rn.f32.bf16 %0, al, bl, %3;\n\t}"
      : "=f"(d)
      : "r"(a2), "r"(b2), "f"(c));
  return d;
}
template <>
__device__ __forceinline__ float fma_mixed_hi<__nv_bfloat16>(uint32_t a2, uint32_t b2, float c) {
  float d;
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
      "mov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
      "fma.rn.f32.bf16 %0, ah, bh, %3;\n\t}"
      : "=f"(d)
      : "r"(a2), "r"(b2), "f"(c));
  return d;
}

}  // namespace ptx

// ---------------------------------------------------------------- candidate keys
// A candidate (score, local row id) is one 64-bit key whose unsigned order is
// (score descending-is-larger, id ascending-is-larger):  hi = order-preserving map
// of the fp32 score, lo = ~id.  key 0 is the "empty slot" sentinel (it would decode
// to a NaN score, which no finite dot product produces).
__host__ __device__ __forceinline__ uint32_t score_to_ord(float s) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(s);
#else
  union { float f; uint32_t u; } c; c.f = s; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord_to_score(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float s, uint32_t id) {
  return (static_cast<uint64_t>(score_to_ord(s)) << 32) | static_cast<uint64_t>(~id);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return ord_to_score(static_cast<uint32_t>(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t k) { return ~static_cast<uint32_t>(k); }

}  // namespace sgic
