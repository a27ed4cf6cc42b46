// gemm_probe.cu -- isolates the fused kernel's contraction loop (S[N x LD] @ W[LD x 64], fp32 FFMA) to find
// the register/shared-memory blocking that gets closest to the FMA-pipe peak on sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

constexpr int N = 207, LD = 108, OUT = 32, WLD = 96, KG = LD / 4;

// MODE 0: scalar FFMA, loads at loop top (current v2 code)
// MODE 1: FFMA2, loads at loop top (current v3 code)
// MODE 2: scalar FFMA, explicit register double buffering of A and B
// MODE 3: FFMA2, explicit register double buffering
// MODE 4: scalar FFMA, K-step of 2 float4 (8 k) per iteration
template <int RT, int NW, int MODE>
__global__ void __launch_bounds__(NW * 32, 1) probe(float* out, long long* cyc, int reps) {
  extern __shared__ __align__(128) float sm[];
  float* S = sm;
  float* W = sm + N * LD;
  constexpr int NT = NW * 32, CG = OUT / 4, RQ = 32 / CG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < N * LD; i += NT) S[i] = 0.001f * (i % 97);
  for (int i = tid; i < LD * WLD; i += NT) W[i] = 0.002f * (i % 89);
  __syncthreads();
  const int cg = lane % CG, rq = lane / CG, c0 = cg * 4;
  const int row0 = warp * (RQ * RT) + rq;
  int soff[RT];
#pragma unroll
  for (int i = 0; i < RT; ++i) { int r = row0 + RQ * i; soff[i] = (r < N ? r : N - 1) * LD; }
  float total = 0.f;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if (MODE == 0 || MODE == 2 || MODE == 4) {
      float accz[RT][4], accr[RT][4];
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) { accz[i][c] = 0.f; accr[i][c] = 0.f; }
      if (MODE == 0) {
#pragma unroll 1
        for (int kg = 0; kg < KG; ++kg) {
          float4 a[RT];
#pragma unroll
          for (int i = 0; i < RT; ++i) a[i] = ld4(S + soff[i] + 4 * kg);
          const float* wrow = W + (4 * kg) * WLD + c0;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float4 bz = ld4(wrow + kk * WLD), br = ld4(wrow + kk * WLD + OUT);
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
              accz[i][0] = fmaf(av, bz.x, accz[i][0]); accz[i][1] = fmaf(av, bz.y, accz[i][1]);
              accz[i][2] = fmaf(av, bz.z, accz[i][2]); accz[i][3] = fmaf(av, bz.w, accz[i][3]);
              accr[i][0] = fmaf(av, br.x, accr[i][0]); accr[i][1] = fmaf(av, br.y, accr[i][1]);
              accr[i][2] = fmaf(av, br.z, accr[i][2]); accr[i][3] = fmaf(av, br.w, accr[i][3]);
            }
          }
        }
      } else if (MODE == 2) {
        float4 a[RT], bz[4], br[4];
#pragma unroll
        for (int i = 0; i < RT; ++i) a[i] = ld4(S + soff[i]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) { bz[kk] = ld4(W + kk * WLD + c0); br[kk] = ld4(W + kk * WLD + c0 + OUT); }
#pragma unroll 1
        for (int kg = 0; kg < KG; ++kg) {
          float4 an[RT], bzn[4], brn[4];
          const int kn = kg + 1 < KG ? kg + 1 : kg;
#pragma unroll
          for (int i = 0; i < RT; ++i) an[i] = ld4(S + soff[i] + 4 * kn);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) { bzn[kk] = ld4(W + (4 * kn + kk) * WLD + c0); brn[kk] = ld4(W + (4 * kn + kk) * WLD + c0 + OUT); }
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
              accz[i][0] = fmaf(av, bz[kk].x, accz[i][0]); accz[i][1] = fmaf(av, bz[kk].y, accz[i][1]);
              accz[i][2] = fmaf(av, bz[kk].z, accz[i][2]); accz[i][3] = fmaf(av, bz[kk].w, accz[i][3]);
              accr[i][0] = fmaf(av, br[kk].x, accr[i][0]); accr[i][1] = fmaf(av, br[kk].y, accr[i][1]);
              accr[i][2] = fmaf(av, br[kk].z, accr[i][2]); accr[i][3] = fmaf(av, br[kk].w, accr[i][3]);
            }
          }
#pragma unroll
          for (int i = 0; i < RT; ++i) a[i] = an[i];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) { bz[kk] = bzn[kk]; br[kk] = brn[kk]; }
        }
      } else {
#pragma unroll 1
        for (int kg = 0; kg + 1 < KG; kg += 2) {
          float4 a[RT], a2[RT];
#pragma unroll
          for (int i = 0; i < RT; ++i) { a[i] = ld4(S + soff[i] + 4 * kg); a2[i] = ld4(S + soff[i] + 4 * kg + 4); }
          const float* wrow = W + (4 * kg) * WLD + c0;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const float4 bz = ld4(wrow + kk * WLD), br = ld4(wrow + kk * WLD + OUT);
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              const float4 aa = kk < 4 ? a[i] : a2[i];
              const float av = (kk & 3) == 0 ? aa.x : ((kk & 3) == 1 ? aa.y : ((kk & 3) == 2 ? aa.z : aa.w));
              accz[i][0] = fmaf(av, bz.x, accz[i][0]); accz[i][1] = fmaf(av, bz.y, accz[i][1]);
              accz[i][2] = fmaf(av, bz.z, accz[i][2]); accz[i][3] = fmaf(av, bz.w, accz[i][3]);
              accr[i][0] = fmaf(av, br.x, accr[i][0]); accr[i][1] = fmaf(av, br.y, accr[i][1]);
              accr[i][2] = fmaf(av, br.z, accr[i][2]); accr[i][3] = fmaf(av, br.w, accr[i][3]);
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) total += accz[i][c] + accr[i][c];
    } else {
      float2 accz[RT][2], accr[RT][2];
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int c = 0; c < 2; ++c) { accz[i][c] = make_float2(0.f, 0.f); accr[i][c] = make_float2(0.f, 0.f); }
      if (MODE == 1) {
#pragma unroll 1
        for (int kg = 0; kg < KG; ++kg) {
          float4 a[RT];
#pragma unroll
          for (int i = 0; i < RT; ++i) a[i] = ld4(S + soff[i] + 4 * kg);
          const float* wrow = W + (4 * kg) * WLD + c0;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float4 bz = ld4(wrow + kk * WLD), br = ld4(wrow + kk * WLD + OUT);
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
              const float2 aa = make_float2(av, av);
              accz[i][0] = ffma2(aa, make_float2(bz.x, bz.y), accz[i][0]);
              accz[i][1] = ffma2(aa, make_float2(bz.z, bz.w), accz[i][1]);
              accr[i][0] = ffma2(aa, make_float2(br.x, br.y), accr[i][0]);
              accr[i][1] = ffma2(aa, make_float2(br.z, br.w), accr[i][1]);
            }
          }
        }
      } else {
        float4 a[RT], bz[4], br[4];
#pragma unroll
        for (int i = 0; i < RT; ++i) a[i] = ld4(S + soff[i]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) { bz[kk] = ld4(W + kk * WLD + c0); br[kk] = ld4(W + kk * WLD + c0 + OUT); }
#pragma unroll 1
        for (int kg = 0; kg < KG; ++kg) {
          float4 an[RT], bzn[4], brn[4];
          const int kn = kg + 1 < KG ? kg + 1 : kg;
#pragma unroll
          for (int i = 0; i < RT; ++i) an[i] = ld4(S + soff[i] + 4 * kn);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) { bzn[kk] = ld4(W + (4 * kn + kk) * WLD + c0); brn[kk] = ld4(W + (4 * kn + kk) * WLD + c0 + OUT); }
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
              const float2 aa = make_float2(av, av);
              accz[i][0] = ffma2(aa, make_float2(bz[kk].x, bz[kk].y), accz[i][0]);
              accz[i][1] = ffma2(aa, make_float2(bz[kk].z, bz[kk].w), accz[i][1]);
              accr[i][0] = ffma2(aa, make_float2(br[kk].x, br[kk].y), accr[i][0]);
              accr[i][1] = ffma2(aa, make_float2(br[kk].z, br[kk].w), accr[i][1]);
            }
          }
#pragma unroll
          for (int i = 0; i < RT; ++i) a[i] = an[i];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) { bz[kk] = bzn[kk]; br[kk] = brn[kk]; }
        }
      }
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int c = 0; c < 2; ++c) total += accz[i][c].x + accz[i][c].y + accr[i][c].x + accr[i][c].y;
    }
    __syncthreads();
  }
  long long t1 = clock64();
  out[blockIdx.x * NT + tid] = total;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int RT, int NW, int MODE>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int smem = (N * LD + LD * WLD) * 4;
  cudaFuncSetAttribute(probe<RT, NW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int reps = 200;
  probe<RT, NW, MODE><<<148, NW * 32, smem>>>(out, cyc, 5);
  probe<RT, NW, MODE><<<148, NW * 32, smem>>>(out, cyc, reps);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  double c = (double)h[0] / reps;
  double fma = (double)N * LD * 64;  // useful FMAs per GEMM1
  printf("%-40s RT=%d NW=%2d  cycles/GEMM1=%8.0f  useful FMA/clk/SM=%6.1f  %s\n", name, RT, NW, c, fma / c, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<7, 8, 0>("scalar FFMA, loads at top");
  run<4, 16, 0>("scalar FFMA, loads at top");
  run<7, 8, 1>("FFMA2, loads at top");
  run<4, 16, 1>("FFMA2, loads at top");
  run<7, 8, 2>("scalar FFMA, reg double-buffer");
  run<4, 16, 2>("scalar FFMA, reg double-buffer");
  run<7, 8, 3>("FFMA2, reg double-buffer");
  run<4, 16, 3>("FFMA2, reg double-buffer");
  run<7, 8, 4>("scalar FFMA, 8-k step");
  run<4, 16, 4>("scalar FFMA, 8-k step");
  return 0;
}
