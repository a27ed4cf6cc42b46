// dcrnn_bwd.cu -- backward of the fused DCRNN sequence for graphs that fit one SM's shared memory (K = 2, 32 hidden).
//
// What autograd would replay for BatchedDCRNN.forward (torch_geometric_temporal/nn/recurrent/dcrnn.py:429-475, cell
// :172-219) is a reverse-time recurrence on dL/dH.  Two kernels:
//
//   k_dcrnn_bwd_basis  one CTA per (t, window): rebuilds S1 = [U | P_o U | P_i U] for U = [X_t | H_{t-1}] and
//                      S2 for U = [X_t | H_{t-1} * R_t] from the forward outputs and the gate stash -- the
//                      operands of the weight-gradient contractions, which do NOT depend on the recurrence
//                      and are therefore done for all steps at once (two GEMMs over all (t,b,n) rows).
//   k_dcrnn_bwd_seq    one CTA per window, persistent over the T steps in reverse: per step
//                        dS2 = dpre_h @ Wh^T            (FFMA, operands in shared memory)
//                        dU2 = adjoint of the basis     (transposed-CSR gather out of shared memory)
//                        gate derivatives (z, r)        (pointwise, stash read from global)
//                        dS1 = dpre_zr @ Wzr^T ; dU1 = adjoint ; dH_{t-1}, dX_t
//                      dL/dH never leaves shared memory; d pre-activations are streamed out for the weight GEMMs.
//
// Everything here is fp32 FFMA: the per-window GEMMs are 207x32x102 / 207x64x102, the problem is latency-bound on
// the serial recurrence with 64 windows per step (the reference's batch size), not throughput-bound.
#include "common.cuh"

namespace stmp {
namespace {

constexpr int kCo = 32;          // hidden size served by these kernels
constexpr int kDpPitch = 2 * kCo + 1;   // odd pitch: lane = row reads of dp[row][k] are conflict-free

// ------------------------------------------------------------------------------------------------------------------
struct BasisParams {
  const int* rp[2]; const int2* cv[2];      // forward CSRs (by destination): y[i] = sum_k val_k x[col_k]
  int N, Ci, B, T, ld;
  const float* x; long long x_bs, x_ts;     // X[b,t] = x + b*x_bs + t*x_ts, (N, Ci) dense
  const float* out; const float* h0; const float* stash;
  float* S1; float* S2;                     // (T*B, N, ld), ld >= 3*(Ci+Co)
};

__global__ void __launch_bounds__(256) k_dcrnn_bwd_basis(BasisParams p) {
  extern __shared__ float sm[];
  const int N = p.N, Ci = p.Ci, C = p.Ci + kCo;
  float* U1 = sm;
  float* U2 = sm + N * C;
  const int q = blockIdx.x, t = q / p.B, b = q - t * p.B;
  const long long bt = (long long)b * p.T + t;
  float* s1 = p.S1 + (long long)q * N * p.ld;
  float* s2 = p.S2 + (long long)q * N * p.ld;
  for (int idx = threadIdx.x; idx < N * C; idx += blockDim.x) {
    const int n = idx / C, c = idx - n * C;
    float v1, v2;
    if (c < Ci) {
      v1 = v2 = __ldg(p.x + b * p.x_bs + t * p.x_ts + n * Ci + c);
    } else {
      const int cc = c - Ci;
      float h = 0.f;
      if (t > 0) h = __ldg(p.out + ((bt - 1) * N + n) * kCo + cc);
      else if (p.h0) h = __ldg(p.h0 + ((long long)b * N + n) * kCo + cc);
      const float r = __ldg(p.stash + ((bt * 3 + 1) * N + n) * kCo + cc);
      v1 = h; v2 = h * r;
    }
    U1[idx] = v1; U2[idx] = v2;
    s1[(long long)n * p.ld + c] = v1;
    s2[(long long)n * p.ld + c] = v2;
  }
  __syncthreads();
#pragma unroll
  for (int op = 0; op < 2; ++op) {
    const int* rp = p.rp[op];
    const int2* cv = p.cv[op];
    for (int idx = threadIdx.x; idx < N * C; idx += blockDim.x) {
      const int n = idx / C, c = idx - n * C;
      float a1 = 0.f, a2 = 0.f;
      const int k1 = __ldg(rp + n + 1);
      for (int k = __ldg(rp + n); k < k1; ++k) {     // same multiply / add order as k_spmm (bit-identical basis)
        const int2 e = __ldg(cv + k);
        const float w = __int_as_float(e.y);
        a1 = __fadd_rn(a1, __fmul_rn(w, U1[e.x * C + c]));
        a2 = __fadd_rn(a2, __fmul_rn(w, U2[e.x * C + c]));
      }
      s1[(long long)n * p.ld + (1 + op) * C + c] = a1;
      s2[(long long)n * p.ld + (1 + op) * C + c] = a2;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
struct BwdParams {
  const int* rp[2]; const int2* cv[2];      // transposed CSRs (by source): (A^T y)[j] = sum over edges leaving j
  int N, Ci, B, T;
  const float* gout; const float* out; const float* h0; const float* stash;
  const float* whsT; const float* wzrT;     // (Co, 3C) and (2Co, 3C), row-major
  float* dph_all; float* dpzr_all;          // (T,B,N,Co), (T,B,N,2Co)
  float* dx;                                // (B,T,N,Ci) or null
  float* dh0;                               // (B,N,Co)
};

// buf[row][half*HALF .. +HALF) = sum_k dp[row][k] * W[k][half*HALF ..]: one thread per (row, column half).
template <int HALF, int KD>
__device__ __forceinline__ void gemm_rows(const float* __restrict__ dp, const float* __restrict__ W, float* __restrict__ buf, int N) {
  constexpr int NCOL = 2 * HALF;
  const int tid = threadIdx.x;
  if (tid >= 2 * N) return;
  const int half = tid >= N ? 1 : 0, row = tid - half * N;
  float acc[HALF];
#pragma unroll
  for (int j = 0; j < HALF; ++j) acc[j] = 0.f;
  const float* a = dp + row * kDpPitch;
  const float4* w = reinterpret_cast<const float4*>(W + half * HALF);
#pragma unroll 2
  for (int k = 0; k < KD; ++k) {
    const float av = a[k];
#pragma unroll
    for (int j = 0; j < HALF / 4; ++j) {
      const float4 w4 = w[k * (NCOL / 4) + j];        // same address across the warp: broadcast
      acc[4 * j] = fmaf(av, w4.x, acc[4 * j]);
      acc[4 * j + 1] = fmaf(av, w4.y, acc[4 * j + 1]);
      acc[4 * j + 2] = fmaf(av, w4.z, acc[4 * j + 2]);
      acc[4 * j + 3] = fmaf(av, w4.w, acc[4 * j + 3]);
    }
  }
  float4* o = reinterpret_cast<float4*>(buf + row * NCOL + half * HALF);
#pragma unroll
  for (int j = 0; j < HALF / 4; ++j) o[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
}

// dU[n][c] = dS[n][c] + sum_op sum_{edges of row n of A_op^T} val * dS[col][(1+op)*C + c]   (adjoint of U -> [U|P_oU|P_iU])
template <int NCOL>
__device__ __forceinline__ float adjoint_at(const BwdParams& p, const float* __restrict__ buf, int n, int c, int C) {
  float v = buf[n * NCOL + c];
#pragma unroll
  for (int op = 0; op < 2; ++op) {
    const int k1 = __ldg(p.rp[op] + n + 1);
    const float* src = buf + (1 + op) * C + c;
    for (int k = __ldg(p.rp[op] + n); k < k1; ++k) {
      const int2 e = __ldg(p.cv[op] + k);
      v = fmaf(__int_as_float(e.y), src[e.x * NCOL], v);
    }
  }
  return v;
}

template <int HALF>
__global__ void __launch_bounds__(512, 1) k_dcrnn_bwd_seq(BwdParams p) {
  constexpr int NCOL = 2 * HALF;
  extern __shared__ __align__(16) float sm[];
  const int N = p.N, Ci = p.Ci, C = p.Ci + kCo, T = p.T, b = blockIdx.x, tid = threadIdx.x;
  float* Wh = sm;                            // [Co][NCOL]
  float* Wzr = Wh + kCo * NCOL;              // [2Co][NCOL]
  float* buf = Wzr + 2 * kCo * NCOL;         // [N][NCOL]
  float* dp = buf + N * NCOL;                // [N][kDpPitch]
  float* G = dp + N * kDpPitch;              // [N][Co]   dL/dH_t (open) -> partial dL/dH_{t-1}
  float* dXp = G + N * kCo;                  // [N][4]    dU2[:, :Ci] waiting for dU1
  // ---- weights, zero padded to NCOL columns
  for (int i = tid; i < 3 * kCo * NCOL; i += blockDim.x) {
    const int r = i / NCOL, c = i - r * NCOL;
    float v = 0.f;
    if (c < 3 * C) v = r < kCo ? __ldg(p.whsT + r * 3 * C + c) : __ldg(p.wzrT + (r - kCo) * 3 * C + c);
    sm[i] = v;
  }
  // ---- open step T-1
  const long long bT = (long long)b * T;
  for (int i = tid; i < N * kCo; i += blockDim.x) {
    const int n = i / kCo, cc = i - n * kCo;
    const long long bt = bT + (T - 1);
    const float g = __ldg(p.gout + bt * N * kCo + i);
    const float z = __ldg(p.stash + (bt * 3 + 0) * N * kCo + i), ht = __ldg(p.stash + (bt * 3 + 2) * N * kCo + i);
    const float d = g * (1.f - z) * (1.f - ht * ht);
    G[i] = g;
    dp[n * kDpPitch + cc] = d;
    p.dph_all[(((long long)(T - 1) * p.B + b) * N) * kCo + i] = d;
  }
  __syncthreads();
#pragma unroll 1
  for (int t = T - 1; t >= 0; --t) {
    const long long bt = bT + t;
    const float* st = p.stash + bt * 3 * N * kCo;
    const float* hprev = t > 0 ? p.out + (bt - 1) * N * kCo : (p.h0 ? p.h0 + (long long)b * N * kCo : nullptr);
    // dS2 = dpre_h @ Wh^T
    gemm_rows<HALF, kCo>(dp, Wh, buf, N);
    __syncthreads();
    // dU2 = adjoint; d pre-activations of z and r; partial carry  g*Z + dHR*R
    float* dpzr = p.dpzr_all + (((long long)t * p.B + b) * N) * 2 * kCo;
    for (int idx = tid; idx < N * C; idx += blockDim.x) {
      const int n = idx / C, c = idx - n * C;
      const float v = adjoint_at<NCOL>(p, buf, n, c, C);
      if (c < Ci) {
        dXp[n * 4 + c] = v;
      } else {
        const int cc = c - Ci, i = n * kCo + cc;
        const float hp = hprev ? __ldg(hprev + i) : 0.f;
        const float z = __ldg(st + i), r = __ldg(st + N * kCo + i), ht = __ldg(st + 2 * N * kCo + i);
        const float g = G[i];
        const float dpz = g * (hp - ht) * z * (1.f - z);
        const float dpr = v * hp * r * (1.f - r);
        dp[n * kDpPitch + cc] = dpz;
        dp[n * kDpPitch + kCo + cc] = dpr;
        dpzr[n * 2 * kCo + cc] = dpz;
        dpzr[n * 2 * kCo + kCo + cc] = dpr;
        G[i] = g * z + v * r;
      }
    }
    __syncthreads();
    // dS1 = dpre_zr @ Wzr^T
    gemm_rows<HALF, 2 * kCo>(dp, Wzr, buf, N);
    __syncthreads();
    // dU1 = adjoint; dX_t; dL/dH_{t-1}; open step t-1
    const float* stn = st - 3 * N * kCo;     // stash of step t-1 (only dereferenced when t > 0)
    for (int idx = tid; idx < N * C; idx += blockDim.x) {
      const int n = idx / C, c = idx - n * C;
      const float v = adjoint_at<NCOL>(p, buf, n, c, C);
      if (c < Ci) {
        if (p.dx) p.dx[(bt * N + n) * Ci + c] = dXp[n * 4 + c] + v;
      } else {
        const int cc = c - Ci, i = n * kCo + cc;
        const float dh = G[i] + v;
        if (t > 0) {
          const float g = __ldg(p.gout + (bt - 1) * N * kCo + i) + dh;
          const float z = __ldg(stn + i), ht = __ldg(stn + 2 * N * kCo + i);
          const float d = g * (1.f - z) * (1.f - ht * ht);
          G[i] = g;
          dp[n * kDpPitch + cc] = d;
          p.dph_all[(((long long)(t - 1) * p.B + b) * N) * kCo + i] = d;
        } else {
          p.dh0[(long long)b * N * kCo + i] = dh;
        }
      }
    }
    __syncthreads();
  }
}

inline int ncol_for(int cin) { return (3 * (cin + kCo) + 7) / 8 * 8; }
inline size_t seq_smem(int N, int cin) {
  const int ncol = ncol_for(cin);
  return sizeof(float) * ((size_t)3 * kCo * ncol + (size_t)N * (ncol + kDpPitch + kCo + 4));
}
inline bool bwd_supported(const stmp_plan* plan, long long cin, long long cout, long long K) {
  if (!plan || plan->flavor != STMP_FLAVOR_DCONV || plan->n_ops != 2) return false;
  if (K != 2 || cout != kCo || cin < 1 || cin > 4) return false;
  const int ncol = ncol_for((int)cin);
  if (ncol != 104 && ncol != 112) return false;
  return 2 * plan->n <= 512 && seq_smem(plan->n, (int)cin) <= 227 * 1024;
}

}  // namespace
}  // namespace stmp

using namespace stmp;

extern "C" int stmp_dcrnn_bwd_supported(const stmp_plan* plan, int64_t cin, int64_t cout, int64_t K) {
  return bwd_supported(plan, cin, cout, K) ? 1 : 0;
}

extern "C" int stmp_dcrnn_bwd_basis(const stmp_plan* plan, int64_t B, int64_t T, int64_t cin, int64_t cout, const float* x,
                                    int64_t x_bstride, int64_t x_tstride, const float* out, const float* h0, const float* stash,
                                    float* S1, float* S2, int64_t ld, void* stream) {
  STMP_REQUIRE(plan != nullptr, STMP_EINVAL, "stmp_dcrnn_bwd_basis: plan is NULL");
  STMP_REQUIRE(bwd_supported(plan, cin, cout, 2), STMP_EUNSUPPORTED, "stmp_dcrnn_bwd_basis: configuration not served (K=2, cout=32, cin<=4, small graph)");
  STMP_REQUIRE(x && out && stash && S1 && S2, STMP_EINVAL, "stmp_dcrnn_bwd_basis: NULL tensor");
  STMP_REQUIRE(B >= 0 && T > 0 && ld >= 3 * (cin + cout), STMP_ESHAPE, "stmp_dcrnn_bwd_basis: bad sizes");
  if (B == 0) return STMP_OK;
  BasisParams p;
  for (int o = 0; o < 2; ++o) { p.rp[o] = plan->fwd[o].rowptr; p.cv[o] = plan->fwd[o].cv; }
  p.N = plan->n; p.Ci = (int)cin; p.B = (int)B; p.T = (int)T; p.ld = (int)ld;
  p.x = x; p.x_bs = x_bstride; p.x_ts = x_tstride; p.out = out; p.h0 = h0; p.stash = stash; p.S1 = S1; p.S2 = S2;
  const size_t smem = sizeof(float) * 2 * (size_t)plan->n * (cin + cout);
  STMP_REQUIRE(smem <= 100 * 1024, STMP_EUNSUPPORTED, "stmp_dcrnn_bwd_basis: graph too large for the shared-memory tile");
  STMP_CUDA_OK(cudaFuncSetAttribute(k_dcrnn_bwd_basis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_dcrnn_bwd_basis<<<(unsigned)(B * T), 256, smem, (cudaStream_t)stream>>>(p);
  STMP_LAUNCH_OK("k_dcrnn_bwd_basis");
  return STMP_OK;
}

extern "C" int stmp_dcrnn_bwd_seq(const stmp_plan* plan, int64_t B, int64_t T, int64_t cin, int64_t cout, const float* gout,
                                  const float* out, const float* h0, const float* stash, const float* whsT, const float* wzrT,
                                  float* dph_all, float* dpzr_all, float* dx, float* dh0, void* stream) {
  STMP_REQUIRE(plan != nullptr, STMP_EINVAL, "stmp_dcrnn_bwd_seq: plan is NULL");
  STMP_REQUIRE(bwd_supported(plan, cin, cout, 2), STMP_EUNSUPPORTED, "stmp_dcrnn_bwd_seq: configuration not served (K=2, cout=32, cin<=4, small graph)");
  STMP_REQUIRE(gout && out && stash && whsT && wzrT && dph_all && dpzr_all && dh0, STMP_EINVAL, "stmp_dcrnn_bwd_seq: NULL tensor");
  STMP_REQUIRE(B >= 0 && T > 0, STMP_ESHAPE, "stmp_dcrnn_bwd_seq: bad sizes");
  if (B == 0) return STMP_OK;
  BwdParams p;
  for (int o = 0; o < 2; ++o) { p.rp[o] = plan->bwd[o].rowptr; p.cv[o] = plan->bwd[o].cv; }
  p.N = plan->n; p.Ci = (int)cin; p.B = (int)B; p.T = (int)T;
  p.gout = gout; p.out = out; p.h0 = h0; p.stash = stash; p.whsT = whsT; p.wzrT = wzrT;
  p.dph_all = dph_all; p.dpzr_all = dpzr_all; p.dx = dx; p.dh0 = dh0;
  const size_t smem = seq_smem(plan->n, (int)cin);
  if (ncol_for((int)cin) == 104) {
    STMP_CUDA_OK(cudaFuncSetAttribute(k_dcrnn_bwd_seq<52>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_dcrnn_bwd_seq<52><<<(unsigned)B, 512, smem, (cudaStream_t)stream>>>(p);
  } else {
    STMP_CUDA_OK(cudaFuncSetAttribute(k_dcrnn_bwd_seq<56>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_dcrnn_bwd_seq<56><<<(unsigned)B, 512, smem, (cudaStream_t)stream>>>(p);
  }
  STMP_LAUNCH_OK("k_dcrnn_bwd_seq");
  return STMP_OK;
}
