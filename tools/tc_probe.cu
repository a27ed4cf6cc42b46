// tc_probe.cu -- stand-alone validation of the tcgen05 building block planned for the fused DCRNN kernel:
// D[256 x N] (fp32, TMEM) = A[256 x 112] * B[N x 112]^T with fp32 inputs split into fp16 hi/lo halves
// (3 MMAs: hi*hi + lo*hi + hi*lo => ~22-bit mantissa), operands hand-written into 128B-swizzled K-major shared
// memory panels, accumulators read back with tcgen05.ld.  Compares against an fp64 host reference.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>

constexpr int M_ROWS = 207, M_PAD = 256, K = 112, KP = 128, NB = 96;  // K padded to 2 panels of 64
constexpr int PANEL_BYTES_A = M_PAD * 128;                            // one K-panel (64 fp16) of A: 256 rows x 128 B
constexpr int PANEL_BYTES_B = NB * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row, k) inside a [rows x 64 fp16] panel stored K-major with the 128-byte swizzle:
// row pitch 128 B, 16-byte chunk index XORed with (row % 8); 8-row groups are 1024 B apart.
__device__ __host__ __forceinline__ int sw128_off(int row, int k) {
  const int chunk = (k >> 3) ^ (row & 7);
  return row * 128 + chunk * 16 + (k & 7) * 2;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version [46,48)=1,
  // base_offset [49,52)=0, layout_type [61,64): 2 = SWIZZLE_128B.  K-major swizzled: SBO = 1024 B (8 rows x 128 B), LBO unused.
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(0) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  // InstrDescriptor: c_format [4,6)=1 (F32), a_format [7,10)=0 (F16), b_format [10,13)=0 (F16), a_major [15]=0 (K), b_major [16]=0 (K),
  // n_dim [17,23) = N>>3, m_dim [24,29) = M>>4
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(256, 1) tc_gemm(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D1,
                                                  float* __restrict__ D2, int passes) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_hi = smem;                          // 2 panels
  unsigned char* a_lo = a_hi + 2 * PANEL_BYTES_A;
  unsigned char* b_hi = a_lo + 2 * PANEL_BYTES_A;
  unsigned char* b_lo = b_hi + 2 * PANEL_BYTES_B;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- operands: fp32 -> fp16 hi/lo, swizzled K-major panels ----------------------------------------------
  for (int idx = tid; idx < M_PAD * KP; idx += 256) {
    const int r = idx / KP, k = idx % KP;
    const float v = (r < M_ROWS && k < K) ? A[r * K + k] : 0.f;
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const int off = (k >> 6) * PANEL_BYTES_A + sw128_off(r, k & 63);
    *reinterpret_cast<__half*>(a_hi + off) = h;
    *reinterpret_cast<__half*>(a_lo + off) = l;
  }
  for (int idx = tid; idx < NB * KP; idx += 256) {
    const int r = idx / KP, k = idx % KP;
    const float v = (k < K) ? B[r * K + k] : 0.f;
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const int off = (k >> 6) * PANEL_BYTES_B + sw128_off(r, k & 63);
    *reinterpret_cast<__half*>(b_hi + off) = h;
    *reinterpret_cast<__half*>(b_lo + off) = l;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy operand writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  // ---- MMAs: one thread issues everything ---------------------------------------------------------------------
  if (tid == 0) {
    const uint32_t id1 = make_idesc(128, 64), id2 = make_idesc(128, 32);
    for (int tile = 0; tile < 2; ++tile) {
      for (int g = 0; g < 2; ++g) {  // g=0: N=64 (B rows 0..63) -> cols [64*tile, +64) ; g=1: N=32 (B rows 64..95) -> cols [128 + 32*tile, +32)
        const uint32_t dcol = g == 0 ? 64 * tile : 128 + 32 * tile;
        uint32_t acc = 0;
        for (int pass = 0; pass < passes; ++pass) {
          const unsigned char* ap = (pass == 1) ? a_lo : a_hi;   // hi*hi, lo*hi, hi*lo
          const unsigned char* bp = (pass == 2) ? b_lo : b_hi;
          for (int ks = 0; ks < K / 16; ++ks) {
            const int panel = ks >> 2, kin = (ks & 3) * 16;
            const uint32_t aaddr = smem_u32(ap + panel * PANEL_BYTES_A + tile * 128 * 128) + kin * 2;
            const uint32_t baddr = smem_u32(bp + panel * PANEL_BYTES_B + (g ? 64 * 128 : 0)) + kin * 2;
            mma_f16(tmem + dcol, make_desc(aaddr), make_desc(baddr), g == 0 ? id1 : id2, acc);
            acc = 1;
          }
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // ---- wait, read back: warp w reads TMEM lanes 32*(w%4).., tile = w/4 ---------------------------------------
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int tile = warp >> 2, q = warp & 3;
  const int row = tile * 128 + q * 32 + lane;
  const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
  uint32_t v[32];
  for (int part = 0; part < 3; ++part) {  // part 0,1: the 64 GEMM1 columns of this tile; part 2: the 32 GEMM2 columns
    const uint32_t col = part < 2 ? 64 * tile + 32 * part : 128 + 32 * tile;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(lane_addr + col));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row < M_ROWS) {
      for (int c = 0; c < 32; ++c) {
        if (part < 2) D1[row * 64 + 32 * part + c] = __uint_as_float(v[c]);
        else D2[row * 32 + c] = __uint_as_float(v[c]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main() {
  std::vector<float> A(M_ROWS * K), B(NB * K);
  srand(1);
  for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 4.0f;
  for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 0.5f;
  for (int i = 0; i < 50; ++i) A[rand() % A.size()] *= 1e-4f;  // exercise the subnormal-lo regime
  float *dA, *dB, *dD1, *dD2;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD1, M_ROWS * 64 * 4); cudaMalloc(&dD2, M_ROWS * 32 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 4 * PANEL_BYTES_A + 4 * PANEL_BYTES_B;
  cudaFuncSetAttribute(tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int passes : {1, 3}) {
    cudaMemset(dD1, 0, M_ROWS * 64 * 4); cudaMemset(dD2, 0, M_ROWS * 32 * 4);
    tc_gemm<<<1, 256, smem>>>(dA, dB, dD1, dD2, passes);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("passes=%d CUDA error: %s\n", passes, cudaGetErrorString(e)); return 1; }
    std::vector<float> D1(M_ROWS * 64), D2(M_ROWS * 32);
    cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
    double maxabs = 0, maxrel = 0, maxabs32 = 0;
    for (int r = 0; r < M_ROWS; ++r)
      for (int n = 0; n < NB; ++n) {
        double ref = 0; float ref32 = 0.f;
        for (int k = 0; k < K; ++k) { ref += (double)A[r * K + k] * B[n * K + k]; ref32 = fmaf(A[r * K + k], B[n * K + k], ref32); }
        const float got = n < 64 ? D1[r * 64 + n] : D2[r * 32 + n - 64];
        const double err = fabs(got - ref);
        if (err > maxabs) maxabs = err;
        if (fabs(ref) > 1e-3 && err / fabs(ref) > maxrel) maxrel = err / fabs(ref);
        if (fabs(ref32 - ref) > maxabs32) maxabs32 = fabs(ref32 - ref);
      }
    printf("passes=%d  max|err|=%.3e  max rel=%.3e   (fp32 FMA chain vs fp64: %.3e)   D[0][0..3]= %f %f %f %f  D[206][95]=%f\n", passes, maxabs,
           maxrel, maxabs32, D1[0], D1[1], D1[2], D1[3], D2[206 * 32 + 31]);
  }
  return 0;
}
