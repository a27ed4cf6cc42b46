// pipe_probe.cu -- microbenchmark of sm_100a issue rates that decide the fused kernel's GEMM design:
// scalar FFMA, packed FFMA2, FFMA with a constant-bank operand, legacy mma.sync (tf32 / bf16).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe tools/pipe_probe.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__constant__ float cW[1024];

template <int MODE>
__global__ void probe(float* out, long long* cycles, int iters, float seed) {
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = seed * i;
  float a0 = seed + threadIdx.x, a1 = seed * 2.f, b0 = seed * 3.f, b1 = seed * 5.f;
  unsigned ua[4] = {__float_as_uint(a0), __float_as_uint(a1), __float_as_uint(b0), __float_as_uint(b1)};
  unsigned ub[2] = {__float_as_uint(b0), __float_as_uint(b1)};
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // scalar FFMA, 32 independent chains, operands rotate
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fmaf(a0, (i & 1) ? b0 : b1, acc[i]);
    } else if (MODE == 1) {  // FFMA2: 16 pairs
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        asm volatile("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rc, {%0, %1};\n\t"
                     "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
                     : "+f"(acc[2 * i]), "+f"(acc[2 * i + 1]) : "f"(a0), "f"(b0), "f"(b1));
      }
    } else if (MODE == 2) {  // FFMA with constant-bank operand (compile-time offsets)
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fmaf(a0, cW[i], acc[i]);
    } else if (MODE == 3) {  // mma.sync m16n8k8 tf32: 8 independent accumulator tiles
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[4 * i]), "+f"(acc[4 * i + 1]), "+f"(acc[4 * i + 2]), "+f"(acc[4 * i + 3])
                     : "r"(ua[0]), "r"(ua[1]), "r"(ua[2]), "r"(ua[3]), "r"(ub[0]), "r"(ub[1]));
      }
    } else if (MODE == 4) {  // mma.sync m16n8k16 bf16
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[4 * i]), "+f"(acc[4 * i + 1]), "+f"(acc[4 * i + 2]), "+f"(acc[4 * i + 3])
                     : "r"(ua[0]), "r"(ua[1]), "r"(ua[2]), "r"(ua[3]), "r"(ub[0]), "r"(ub[1]));
      }
    } else if (MODE == 5) {  // scalar FFMA where both multiplicands change per instruction (no reuse)
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[(i + 7) & 31], (i & 1) ? b0 : b1, acc[i]);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int inst_per_iter, double flops_per_inst_per_warp) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int warps_per_smsp : {1, 2, 4, 8}) {
    int threads = 32 * 4 * warps_per_smsp;
    int iters = 2000;
    probe<MODE><<<148, threads>>>(out, cyc, 10, 1.0f);
    probe<MODE><<<148, threads>>>(out, cyc, iters, 1.0f);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    double c = (double)h[0];
    double inst_per_smsp = (double)iters * inst_per_iter * warps_per_smsp;
    printf("%-34s warps/SMSP=%d  cycles=%9.0f  inst/cycle/SMSP=%.3f  FMA-lanes/clk/SM=%.1f %s\n", name, warps_per_smsp, c, inst_per_smsp / c,
           inst_per_smsp / c * flops_per_inst_per_warp * 4 / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  float w[1024];
  for (int i = 0; i < 1024; ++i) w[i] = 1.0f / (i + 1);
  cudaMemcpyToSymbol(cW, w, sizeof(w));
  run<0>("FFMA 3-reg (a reused)", 32, 64);
  run<5>("FFMA 3-reg (no reuse)", 32, 64);
  run<1>("FFMA2 (f32x2, scalar a)", 16, 128);
  run<2>("FFMA const-bank operand", 32, 64);
  run<3>("mma.sync m16n8k8 tf32", 8, 2.0 * 16 * 8 * 8);
  run<4>("mma.sync m16n8k16 bf16", 8, 2.0 * 16 * 8 * 16);
  return 0;
}
