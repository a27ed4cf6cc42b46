// tc_common.cuh -- tcgen05 / TMEM building blocks shared by the fused graph-GRU kernel and the split-fp16 GEMM:
// UMMA shared-memory / instruction descriptors, MMA issue + commit, TMEM loads, the hand-written SWIZZLE_128B
// K-major operand layout and the fp32 -> fp16 (hi, lo) operand split.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace stmp {

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) (unused: swizzled K-major) | SBO>>4 [32,46) = 1024 B |
  // version [46,48) = 1 (sm_100) | layout_type [61,64) = 2 (SWIZZLE_128B)
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ constexpr uint32_t umma_idesc_f16(int m, int n) {
  // cute::UMMA::InstrDescriptor: c_format [4,6) = 1 (F32); a/b_format = 0 (F16); a/b_major = 0 (K); n>>3 at [17,23); m>>4 at [24,29)
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
      "%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
template <int CW> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[CW]);
template <> __device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t (&v)[8]) { tmem_ld8(taddr, v); }
template <> __device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld32(taddr, v); }
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld16(taddr, v); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// byte offset of element (row, kin) inside a K-panel (kin in [0,64))
__device__ __forceinline__ int sw128(int row, int kin) { return row * 128 + ((((kin >> 3) ^ (row & 7))) << 4) + ((kin & 7) << 1); }

__device__ __forceinline__ uint32_t pack_h2(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// split 4 floats into fp16 hi/lo and store them (8 B each) at k offset `kin` (multiple of 4) of `row`
__device__ __forceinline__ void store_split4(unsigned char* a_hi, unsigned char* a_lo, int row, int kin, float4 v) {
  const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
  const int off = sw128(row, kin);
  *reinterpret_cast<uint2*>(a_hi + off) = make_uint2(pack_h2(h01), pack_h2(h23));
  *reinterpret_cast<uint2*>(a_lo + off) = make_uint2(pack_h2(l01), pack_h2(l23));
}

}  // namespace stmp
