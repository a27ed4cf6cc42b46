// tgcn_attn.cu -- fused temporal-attention + GCN-GRU kernel: A3TGCN / A3TGCN2 (and a TGCN / TGCN2 cell as the one-period case)
// for graphs of ANY size (the tcgen05 graph-GRU kernel stops at 207 nodes; PEMS-BAY has 325).
//
// Reference (nn/recurrent/attentiontemporalgcn.py:130-157, temporalgcn.py:187-233): for every period t
//     G_g = GCNConv_g(X[..., t])            three GCNConvs = gcn_norm + Linear(in,out) + propagate over `out` channels
//     Z = sigma(L_z [G_z | H]),  R = sigma(L_r [G_r | H]),  H~ = tanh(L_h [G_h | H*R]),  H_t = Z*H + (1-Z)*H~
//     out += softmax(attention)[t] * H_t                                     (the SAME H enters every period)
// Because A^(X W) = (A^ X) W the graph part only ever touches the `in` channels: one gather per node yields A^X for ALL
// periods at once (a row of in*periods floats: lane = (feature, period)), and GCNConv's Linear and the gate Linear fold into
//     pre_g = (A^X_t) A_g + H' B_g + c_g          A_g (in x out), B_g (out x out), c_g (out)   [host: TGCN._packed3]
// One warp per (batch row, node), lane = output channel.  H' B is an out x out mat-vec per node with the weight column in
// registers and H' handed round by shuffles; H B_z, H B_r are shared by all periods, only (H*R_t) B_h is per period.
// X[b] (N x in*periods floats, 31 KB at the PEMS-BAY shape) is staged into shared memory with ONE TMA bulk copy per CTA; the
// gather then runs out of shared memory.  Output and nothing else goes back to HBM: algorithmic bytes per (row, node) =
// 4*in*periods (X) + 4*out (H) + 4*out (out).
#include <cuda_runtime.h>

#include "common.cuh"

namespace stmp {
namespace {

struct TgcnArgs {
  const int* rowptr;
  const int2* cv;
  int N, FIN, P, FP;      // FP = FIN * P floats per node of X[b]
  long long B;
  const float* x;         // [B][N][FIN][P] contiguous
  const float* h;         // [B][N][32] (h_bstride) or null
  long long h_bstride;
  const float* A;         // [FIN][96]   columns z | r | h
  const float* Bm;        // [32][96]
  const float* c;         // [96]
  const float* probs;     // [P] or null (one period, weight 1)
  float* out;             // [B][N][32]
  int stage;              // 1: X[b] staged in shared memory by TMA; 0: gathered from global memory (huge graphs)
};

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return 2.0f * __fdividef(1.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }

constexpr int kNodesPerBlock = 64;

template <int NQ, bool HAS_H>
__global__ void __launch_bounds__(256) k_tgcn_attn(const TgcnArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* Xs = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  const int n0 = blockIdx.x * kNodesPerBlock;
  const float* xb = a.x + b * (long long)a.N * a.FP;
  if (a.stage) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
      const uint32_t bytes = (uint32_t)a.N * a.FP * 4u;
      mbar_arrive_expect_tx(&bar, bytes);
      tma_bulk_g2s(Xs, xb, bytes, &bar);
    }
  }
  // weights of my output channel (lane) while the copy is in flight
  float Az[4], Ar[4], Ah[4];
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    Az[f] = f < a.FIN ? __ldg(a.A + f * 96 + lane) : 0.f;
    Ar[f] = f < a.FIN ? __ldg(a.A + f * 96 + 32 + lane) : 0.f;
    Ah[f] = f < a.FIN ? __ldg(a.A + f * 96 + 64 + lane) : 0.f;
  }
  const float cz = __ldg(a.c + lane), cr = __ldg(a.c + 32 + lane), ch = __ldg(a.c + 64 + lane);
  float Bz[HAS_H ? 32 : 1], Br[HAS_H ? 32 : 1], Bh[HAS_H ? 32 : 1];
  if (HAS_H) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      Bz[k] = __ldg(a.Bm + k * 96 + lane);
      Br[k] = __ldg(a.Bm + k * 96 + 32 + lane);
      Bh[k] = __ldg(a.Bm + k * 96 + 64 + lane);
    }
  }
  if (a.stage) {
    __syncthreads();          // barrier initialised before anyone waits on it
    mbar_wait(&bar, 0);
  }
  const float* Xg = a.stage ? Xs : xb;
  const int FP = a.FP, P = a.P;
  const int nend = min(n0 + kNodesPerBlock, a.N);
  for (int n = n0 + warp; n < nend; n += 8) {
    // ---- A^X for all periods: lane holds entries (q*32 + lane) of the node's in*periods row, reference's edge order --------
    float ax[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) ax[q] = 0.f;
    const int beg = __ldg(a.rowptr + n), end = __ldg(a.rowptr + n + 1);
    for (int k = beg; k < end; ++k) {
      const int2 e = __ldg(a.cv + k);
      const float w = __int_as_float(e.y);
      const float* xr = Xg + (long long)e.x * FP;
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        if (q * 32 + lane < FP) ax[q] = __fadd_rn(ax[q], __fmul_rn(w, a.stage ? xr[q * 32 + lane] : __ldg(xr + q * 32 + lane)));
    }
    float h = 0.f, hz = cz, hr = cr;
    if (HAS_H) {
      h = __ldg(a.h + b * a.h_bstride + (long long)n * 32 + lane);
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float hk = __shfl_sync(0xffffffffu, h, k);
        hz = fmaf(hk, Bz[k], hz);
        hr = fmaf(hk, Br[k], hr);
      }
    }
    float acc = 0.f;
    for (int t = 0; t < P; ++t) {
      float pz = hz, pr = hr, ph = ch;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        if (f < a.FIN) {
          const int idx = f * P + t;
          float src = ax[0];
#pragma unroll
          for (int q = 1; q < NQ; ++q) src = (idx >> 5) == q ? ax[q] : src;
          const float v = __shfl_sync(0xffffffffu, src, idx & 31);
          pz = fmaf(v, Az[f], pz);
          pr = fmaf(v, Ar[f], pr);
          ph = fmaf(v, Ah[f], ph);
        }
      }
      const float Z = sigmoid_f(pz);
      float hn;
      if (HAS_H) {
        const float hrr = h * sigmoid_f(pr);
#pragma unroll
        for (int k = 0; k < 32; ++k) ph = fmaf(__shfl_sync(0xffffffffu, hrr, k), Bh[k], ph);
        hn = Z * h + (1.0f - Z) * tanh_f(ph);
      } else {
        hn = (1.0f - Z) * tanh_f(ph);       // H = 0: Z*H vanishes, R is irrelevant
      }
      acc = a.probs ? fmaf(__ldg(a.probs + t), hn, acc) : hn;
    }
    a.out[(b * a.N + n) * 32 + lane] = acc;
  }
}

// ---- backward (H = None: the training configuration of the reference's A3TGCN2 example) ---------------------------------------
// What autograd records for attentiontemporalgcn.py:130-157 with H = 0: per period  Z = sigma(pz), H~ = tanh(ph), H_t = (1 - Z) H~,
// out = sum_t probs[t] H_t with pz = (A^X_t) A_z + c_z, ph = (A^X_t) A_h + c_h (the r gate multiplies H = 0 and has no gradient).
// Given g = dL/dout the kernel recomputes A^X (same gather as the forward) and the gates, and reduces
//     dA_z[f] = sum ax[f,t] dpz,  dA_h[f] = sum ax[f,t] dph,  dc_z = sum dpz,  dc_h = sum dph,  dprobs[t] = sum_j g_j H_t,j
// with dH_t = probs[t] g, dpz = -dH_t H~ Z (1 - Z), dph = dH_t (1 - Z)(1 - H~^2), over all (batch row, node, period).  One warp per
// (row, node), lane = output channel; per-CTA partials (fixed order inside the CTA) and a second launch that sums them in launch order:
// deterministic.  The gradient w.r.t. X is not produced (callers that need it take the op-for-op path).
struct TgcnBwdArgs {
  const int* rowptr;
  const int2* cv;
  int N, FIN, P, FP;
  const float* x; const float* A; const float* c; const float* probs; const float* gout;
  float* partial;         // [gridDim.y * gridDim.x][10 * 32 + 128]: dA_z[4][32] | dA_h[4][32] | dc_z[32] | dc_h[32] | dprobs[128]
  int stage;
};
constexpr int kBwdPartial = 10 * 32 + 128;

template <int NQ>
__global__ void __launch_bounds__(256) k_tgcn_attn_bwd(const TgcnBwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* Xs = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bar;
  __shared__ float red[8][kBwdPartial];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  const int n0 = blockIdx.x * kNodesPerBlock;
  const float* xb = a.x + b * (long long)a.N * a.FP;
  if (a.stage && threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    const uint32_t bytes = (uint32_t)a.N * a.FP * 4u;
    mbar_arrive_expect_tx(&bar, bytes);
    tma_bulk_g2s(Xs, xb, bytes, &bar);
  }
  float Az[4], Ah[4];
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    Az[f] = f < a.FIN ? __ldg(a.A + f * 96 + lane) : 0.f;
    Ah[f] = f < a.FIN ? __ldg(a.A + f * 96 + 64 + lane) : 0.f;
  }
  const float cz = __ldg(a.c + lane), ch = __ldg(a.c + 64 + lane);
  if (a.stage) {
    __syncthreads();
    mbar_wait(&bar, 0);
  }
  const float* Xg = a.stage ? Xs : xb;
  const int FP = a.FP, P = a.P;
  const int nend = min(n0 + kNodesPerBlock, a.N);
  float dAz[4] = {0.f, 0.f, 0.f, 0.f}, dAh[4] = {0.f, 0.f, 0.f, 0.f}, dcz = 0.f, dch = 0.f, dpr[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n = n0 + warp; n < nend; n += 8) {
    float ax[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) ax[q] = 0.f;
    const int beg = __ldg(a.rowptr + n), end = __ldg(a.rowptr + n + 1);
    for (int k = beg; k < end; ++k) {
      const int2 e = __ldg(a.cv + k);
      const float w = __int_as_float(e.y);
      const float* xr = Xg + (long long)e.x * FP;
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        if (q * 32 + lane < FP) ax[q] = __fadd_rn(ax[q], __fmul_rn(w, a.stage ? xr[q * 32 + lane] : __ldg(xr + q * 32 + lane)));
    }
    const float g = __ldg(a.gout + (b * a.N + n) * 32 + lane);
    for (int t = 0; t < P; ++t) {
      float pz = cz, ph = ch, v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        if (f < a.FIN) {
          const int idx = f * P + t;
          float src = ax[0];
#pragma unroll
          for (int q = 1; q < NQ; ++q) src = (idx >> 5) == q ? ax[q] : src;
          v[f] = __shfl_sync(0xffffffffu, src, idx & 31);
          pz = fmaf(v[f], Az[f], pz);
          ph = fmaf(v[f], Ah[f], ph);
        }
      }
      const float Z = sigmoid_f(pz), Ht = tanh_f(ph), hn = (1.0f - Z) * Ht;
      const float dh = (a.probs ? __ldg(a.probs + t) : 1.0f) * g;
      const float dpz = -dh * Ht * Z * (1.0f - Z), dph = dh * (1.0f - Z) * (1.0f - Ht * Ht);
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        dAz[f] = fmaf(v[f], dpz, dAz[f]);
        dAh[f] = fmaf(v[f], dph, dAh[f]);
      }
      dcz += dpz;
      dch += dph;
      float s = g * hn;                                   // dprobs[t] += sum over channels
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == (t & 31)) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if ((t >> 5) == q) dpr[q] += s;
      }
    }
  }
  // ---- CTA reduction in warp order, one partial per CTA ----------------------------------------------------------------------
  float* r = red[warp];
#pragma unroll
  for (int f = 0; f < 4; ++f) { r[f * 32 + lane] = dAz[f]; r[128 + f * 32 + lane] = dAh[f]; }
  r[256 + lane] = dcz;
  r[288 + lane] = dch;
#pragma unroll
  for (int q = 0; q < 4; ++q) r[320 + q * 32 + lane] = dpr[q];
  __syncthreads();
  float* out = a.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kBwdPartial;
  for (int i = threadIdx.x; i < kBwdPartial; i += 256) {
    float s = red[0][i];
#pragma unroll
    for (int w = 1; w < 8; ++w) s += red[w][i];
    out[i] = s;
  }
}

// dA (FIN x 96, r columns zero), dc (96), dprobs (P): sums of the per-CTA partials, 8 sub-sums per output in a fixed association
__global__ void __launch_bounds__(256) k_tgcn_attn_bwd_reduce(int parts, int FIN, int P, const float* __restrict__ partial, float* __restrict__ dA,
                                                              float* __restrict__ dc, float* __restrict__ dprobs) {
  __shared__ float sub[8][32];
  const int x = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + x;                     // index into the partial layout
  const int per = (parts + 7) / 8, q0 = w * per, q1 = (q0 + per < parts) ? q0 + per : parts;
  float s = 0.f;
  if (i < kBwdPartial)
    for (int q = q0; q < q1; ++q) s += partial[(size_t)q * kBwdPartial + i];
  sub[w][x] = s;
  __syncthreads();
  if (w != 0 || i >= kBwdPartial) return;
  float t = sub[0][x];
#pragma unroll
  for (int k = 1; k < 8; ++k) t += sub[k][x];
  if (i < 128) { const int f = i >> 5; if (f < FIN) dA[f * 96 + (i & 31)] = t; }
  else if (i < 256) { const int f = (i - 128) >> 5; if (f < FIN) dA[f * 96 + 64 + (i & 31)] = t; }
  else if (i < 288) dc[i - 256] = t;
  else if (i < 320) dc[64 + i - 288] = t;
  else if (i - 320 < P && dprobs) dprobs[i - 320] = t;
}

template <int NQ>
int launch_nq(const TgcnArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  if (a.h) {
    STMP_CUDA_OK(cudaFuncSetAttribute(k_tgcn_attn<NQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tgcn_attn<NQ, true><<<grid, 256, smem, st>>>(a);
  } else {
    STMP_CUDA_OK(cudaFuncSetAttribute(k_tgcn_attn<NQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tgcn_attn<NQ, false><<<grid, 256, smem, st>>>(a);
  }
  STMP_LAUNCH_OK("k_tgcn_attn");
  return STMP_OK;
}

}  // namespace
}  // namespace stmp

using namespace stmp;

extern "C" int stmp_tgcn_attn_fwd(const stmp_plan* plan, int64_t B, int64_t fin, int64_t periods, const float* x, const float* h,
                                  int64_t h_bstride, const float* A, const float* Bm, const float* c, const float* probs,
                                  float* out, void* stream) {
  STMP_REQUIRE(plan != nullptr, STMP_EINVAL, "stmp_tgcn_attn_fwd: plan is NULL");
  STMP_REQUIRE(plan->n_ops >= 1, STMP_EINVAL, "stmp_tgcn_attn_fwd: plan has no operator");
  STMP_REQUIRE(x && A && Bm && c && out, STMP_EINVAL, "stmp_tgcn_attn_fwd: NULL tensor");
  STMP_REQUIRE(B >= 0 && periods >= 1, STMP_EINVAL, "stmp_tgcn_attn_fwd: bad B/periods");
  if (fin < 1 || fin > 4 || fin * periods > 128)
    return set_error(STMP_EUNSUPPORTED, "fused TGCN-attention kernel takes in_channels <= 4 and in_channels*periods <= 128 (got %lld x %lld)",
                     (long long)fin, (long long)periods);
  if (B == 0) return STMP_OK;
  STMP_REQUIRE(B < 65536, STMP_ESHAPE, "stmp_tgcn_attn_fwd: batch too large for one launch");
  TgcnArgs a;
  a.rowptr = plan->fwd[0].rowptr; a.cv = plan->fwd[0].cv;
  a.N = plan->n; a.FIN = (int)fin; a.P = (int)periods; a.FP = (int)(fin * periods); a.B = B;
  a.x = x; a.h = h; a.h_bstride = h_bstride; a.A = A; a.Bm = Bm; a.c = c; a.probs = probs; a.out = out;
  const size_t bytes = (size_t)a.N * a.FP * 4;
  a.stage = (bytes <= 160 * 1024 && bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0 && ((size_t)a.N * a.FP * 4) % 16 == 0) ? 1 : 0;
  const size_t smem = a.stage ? bytes : 0;
  dim3 grid((unsigned)((a.N + kNodesPerBlock - 1) / kNodesPerBlock), (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
  const int nq = (a.FP + 31) / 32;
  switch (nq) {
    case 1: return launch_nq<1>(a, grid, smem, st);
    case 2: return launch_nq<2>(a, grid, smem, st);
    case 3: return launch_nq<3>(a, grid, smem, st);
    default: return launch_nq<4>(a, grid, smem, st);
  }
}

extern "C" int64_t stmp_tgcn_attn_bwd_workspace_bytes(const stmp_plan* plan, int64_t B) {
  if (!plan || B < 0) return 0;
  return (int64_t)B * ((plan->n + kNodesPerBlock - 1) / kNodesPerBlock) * kBwdPartial * 4;
}

extern "C" int stmp_tgcn_attn_bwd(const stmp_plan* plan, int64_t B, int64_t fin, int64_t periods, const float* x, const float* A,
                                  const float* c, const float* probs, const float* gout, void* workspace, float* dA, float* dc,
                                  float* dprobs, void* stream) {
  STMP_REQUIRE(plan != nullptr, STMP_EINVAL, "stmp_tgcn_attn_bwd: plan is NULL");
  STMP_REQUIRE(plan->n_ops >= 1, STMP_EINVAL, "stmp_tgcn_attn_bwd: plan has no operator");
  STMP_REQUIRE(x && A && c && gout && workspace && dA && dc, STMP_EINVAL, "stmp_tgcn_attn_bwd: NULL tensor");
  STMP_REQUIRE(B >= 1 && periods >= 1, STMP_EINVAL, "stmp_tgcn_attn_bwd: bad B/periods");
  if (fin < 1 || fin > 4 || fin * periods > 128)
    return set_error(STMP_EUNSUPPORTED, "fused TGCN-attention backward takes in_channels <= 4 and in_channels*periods <= 128 (got %lld x %lld)",
                     (long long)fin, (long long)periods);
  STMP_REQUIRE(B < 65536, STMP_ESHAPE, "stmp_tgcn_attn_bwd: batch too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  TgcnBwdArgs a;
  a.rowptr = plan->fwd[0].rowptr; a.cv = plan->fwd[0].cv;
  a.N = plan->n; a.FIN = (int)fin; a.P = (int)periods; a.FP = (int)(fin * periods);
  a.x = x; a.A = A; a.c = c; a.probs = probs; a.gout = gout; a.partial = reinterpret_cast<float*>(workspace);
  const size_t bytes = (size_t)a.N * a.FP * 4;
  a.stage = (bytes <= 128 * 1024 && bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0) ? 1 : 0;
  const size_t smem = a.stage ? bytes : 0;
  dim3 grid((unsigned)((a.N + kNodesPerBlock - 1) / kNodesPerBlock), (unsigned)B);
  STMP_CUDA_OK(cudaMemsetAsync(dA, 0, (size_t)fin * 96 * 4, st));          // the r-gate columns have no gradient when H = 0
  STMP_CUDA_OK(cudaMemsetAsync(dc, 0, 96 * 4, st));
  const int nq = (a.FP + 31) / 32;
#define STMP_TGCN_BWD(NQ)                                                                                                     \
  do {                                                                                                                        \
    STMP_CUDA_OK(cudaFuncSetAttribute(k_tgcn_attn_bwd<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
    k_tgcn_attn_bwd<NQ><<<grid, 256, smem, st>>>(a);                                                                          \
  } while (0)
  switch (nq) {
    case 1: STMP_TGCN_BWD(1); break;
    case 2: STMP_TGCN_BWD(2); break;
    case 3: STMP_TGCN_BWD(3); break;
    default: STMP_TGCN_BWD(4); break;
  }
#undef STMP_TGCN_BWD
  STMP_LAUNCH_OK("k_tgcn_attn_bwd");
  k_tgcn_attn_bwd_reduce<<<(kBwdPartial + 31) / 32, 256, 0, st>>>((int)(grid.x * grid.y), (int)fin, (int)periods, a.partial, dA, dc, dprobs);
  STMP_LAUNCH_OK("k_tgcn_attn_bwd_reduce");
  return STMP_OK;
}
