// dcrnn_seq_tc.cu -- fused DCRNN recurrence with the dense contraction on the 5th-gen tensor cores.
//
// Same contract as k_dcrnn_seq (dcrnn_seq.cu) for K=2, Cout=32, Cin<=4, N<=255; what changes is WHERE the
// S @ [Wz|Wr] and S @ Wh contractions run: `tcgen05.mma.kind::f16` with accumulators in TMEM instead of
// FFMA in registers (the FFMA contraction is 58% of a step in the FFMA kernel, profiles/r01_dcrnn_seq_v3).
//
// fp32 accuracy on fp16 tensor cores: every fp32 operand v is split on the fly into hi = fp16(v) and
// lo = fp16(v - hi) (22 mantissa bits together; products of fp16 pairs are exact in the fp32 accumulator) and
// the product is formed as lo*hi + hi*lo + hi*hi -- three MMAs.  Validated stand-alone in tools/tc_probe.cu:
// max |err| 6e-6 vs fp64 on K=112 dot products, the same order as an fp32 FMA chain (3e-6).
//
// Shared memory (<= 227 KB), all operands written by hand in the canonical K-major SWIZZLE_128B layout
// (row pitch 128 B = 64 fp16, 16-byte chunk index XOR (row % 8), 8-row atoms of 1024 B):
//   A_hi / A_lo : 2 K-panels x 208 rows   k = [H|H*R (32) , P_o H (32)] , [P_i H (32), X(4) P_oX(4) P_iX(4) 0(4)]
//   B_hi / B_lo : 2 K-panels x 96 rows    rows = output channels of z | r | h, same k order (7 k-steps of 16)
//   U  fp32 [N][36] = [H | X] : the gather source of the diffusion (tensor cores only see the fp16 split)
//   graph (padded / pre-scaled / length-sorted, dcrnn_common.cuh), biases, 4 mbarriers.
// TMEM (256 columns): z|r accumulators of the two 128-row tiles at columns [0,64) [64,128), candidate at
// [128,160) [160,192).  TMEM lane == row, so thread (warp w, lane l) owns row 128*(w/4) + 32*(w%4) + l for the
// whole step: it reads its 64+32 accumulator columns with tcgen05.ld, applies the gates, keeps Z and H in
// registers, and writes H*R / H_t back as fp32 (U), as fp16 hi/lo (A panel) and to HBM.  16 warps: each row is
// shared by two threads (channel halves), which also doubles the warps available to hide the gather latency.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "dcrnn_common.cuh"
#include "tc_common.cuh"

namespace stmp {
namespace {

constexpr int kMaxSmemTc = 232448;
constexpr int TC_NT_MAX = 512;             // 16 warps: 2 row tiles x 4 lane quadrants x 2 channel halves (HALVES=2)
constexpr int TC_UP = 36;                  // U row pitch (floats): [H(32) | X(4)]
constexpr int TC_AROWS = 208;              // rows stored per A panel (tile 1 over-reads into the next buffer: harmless)
constexpr int TC_PANEL_A = TC_AROWS * 128;
constexpr int TC_PANEL_B = 96 * 128;
constexpr int TC_TMEM_COLS = 256;

struct TcParams {
  int N, CIN, T;
  long long B;
  const int* rowptr[2];
  const int2* cv[2];
  const float* x;
  const long long* win_start;
  long long x_bstride, x_tstride;
  const float* w[3];
  const float* bias[3];
  const float* h0;
  long long h0_bstride;   // elements between windows' H0 (0: every window starts from the same H0)
  const float* wcat;      // optional prepacked fp32 weights [96][112] in the kernel's k order (else: DConv weights p.w[])
  const float* bcat;      // with wcat: biases [96]
  int n_ops;              // 2: DConv (P_o, P_i); 1: single operator (ChebConv K=2 / GCN); 0: no propagation
  const void* gimg;       // plan's prebuilt shared-memory graph image (TMA bulk source) or null
  int gimg_bytes;
  const void* wimage;     // prebuilt B-operand image (fp16 hi/lo, swizzled, + biases) or null
  float* out;
  float* stash;
  int off_A, off_B, off_U, off_gstart, off_order, off_ce, off_bias, off_bar;
};

// split 32 floats (one row's H-like vector) into panel 0, k 0..31 : 4 swizzled 16-byte chunks for hi and for lo
__device__ __forceinline__ void store_split_row32(unsigned char* a_hi, unsigned char* a_lo, int row, const float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = v[8 * c + 2 * j], b = v[8 * c + 2 * j + 1];
      const __half2 h = __floats2half2_rn(a, b);
      const float2 f = __half22float2(h);
      hw[j] = pack_h2(h);
      lw[j] = pack_h2(__floats2half2_rn(a - f.x, b - f.y));
    }
    const int off = row * 128 + ((c ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

// same for one channel half (16 floats -> chunks 2*half, 2*half+1 of panel 0)
__device__ __forceinline__ void store_split_row16(unsigned char* a_hi, unsigned char* a_lo, int row, int half, const float (&v)[16]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = v[8 * c + 2 * j], b = v[8 * c + 2 * j + 1];
      const __half2 h = __floats2half2_rn(a, b);
      const float2 f = __half22float2(h);
      hw[j] = pack_h2(h);
      lw[j] = pack_h2(__floats2half2_rn(a - f.x, b - f.y));
    }
    const int off = row * 128 + (((2 * half + c) ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

// CW = 32 (whole row) or 16 (channel half `half`): chunks [half*CW/8, +CW/8) of panel 0
template <int CW>
__device__ __forceinline__ void store_split_row(unsigned char* a_hi, unsigned char* a_lo, int row, int half, const float (&v)[CW]) {
#pragma unroll
  for (int c = 0; c < CW / 8; ++c) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = v[8 * c + 2 * j], b = v[8 * c + 2 * j + 1];
      const __half2 h = __floats2half2_rn(a, b);
      const float2 f = __half22float2(h);
      hw[j] = pack_h2(h);
      lw[j] = pack_h2(__floats2half2_rn(a - f.x, b - f.y));
    }
    const int off = row * 128 + (((half * (CW / 8) + c) ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

constexpr int TC_WIMAGE_BYTES = 4 * TC_PANEL_B + 96 * 4;   // B hi (2 panels) | B lo (2 panels) | 96 biases

// fp32 weight of output row n (gate*32 + channel) at k position kk (kernel k order, 0..111)
__device__ __forceinline__ float tc_weight_value(const float* wcat, const float* w0, const float* w1, const float* w2, int CIN, int n,
                                                 int kk) {
  if (wcat) return wcat[n * 112 + kk];
  const int C = 32 + CIN;
  const int gte = n >> 5, o = n & 31;
  const int panel = kk >= 64 ? 1 : 0, kin = kk - 64 * panel;
  int blk, ch;  // blk 0: U, 1: P_o, 2: P_i ; ch = reference channel index or -1
  if (panel == 0) { blk = kin < 32 ? 0 : 1; ch = CIN + (kin & 31); }
  else if (kin < 32) { blk = 2; ch = CIN + kin; }
  else { const int qq = kin - 32; blk = qq >> 2; const int c = qq & 3; ch = (blk < 3 && c < CIN) ? c : -1; }
  if (ch < 0) return 0.f;
  const float* wg = gte == 0 ? w0 : (gte == 1 ? w1 : w2);
  if (blk == 0) return wg[((0 * 2 + 0) * C + ch) * 32 + o] + wg[((1 * 2 + 0) * C + ch) * 32 + o];
  return wg[(((blk - 1) * 2 + 1) * C + ch) * 32 + o];
}

// Builds the B-operand image once per weight update (stmp_*_pack_weights); the fused kernel then TMA-copies it.
__global__ void k_pack_weight_image(const float* wcat, const float* bcat, const float* w0, const float* w1, const float* w2,
                                    const float* b0, const float* b1, const float* b2, int CIN, unsigned char* image) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < 96 * 128) {
    const int n = idx >> 7, kk = idx & 127;
    const float v = kk < 112 ? tc_weight_value(wcat, w0, w1, w2, CIN, n, kk) : 0.f;
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const int off = (kk >> 6) * TC_PANEL_B + sw128(n, kk & 63);
    *reinterpret_cast<__half*>(image + off) = h;
    *reinterpret_cast<__half*>(image + 2 * TC_PANEL_B + off) = l;
  } else if (idx < 96 * 128 + 96) {
    const int j = idx - 96 * 128, gte = j >> 5;
    const float* bg = gte == 0 ? b0 : (gte == 1 ? b1 : b2);
    reinterpret_cast<float*>(image + 4 * TC_PANEL_B)[j] = bcat ? bcat[j] : (bg ? bg[j & 31] : 0.f);
  }
}

// fast, accurate-enough gates (abs err ~2e-7): ex2.approx + rcp.approx
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 2.0f * __fdividef(1.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }

// One diffusion round: gather fp32 rows of U, write the products as fp16 hi/lo into the A panels.
//   op 0 (P_o): H part -> panel 0 k 32..63, X part -> panel 1 k 36..39
//   op 1 (P_i): H part -> panel 1 k 0..31,  X part -> panel 1 k 40..43
template <int TC_NT>
__device__ __forceinline__ void diffuse_tc(const float* U, const GraphSmem g, int N, unsigned char* a_hi, unsigned char* a_lo,
                                           bool with_x, int tid, int s_beg, int s_end) {
  const int j = tid & 7;
  for (int slot = s_beg + (tid >> 3); slot < s_end; slot += TC_NT / 8) {
    const int task = g.order[slot];
    const int op = task >= N ? 1 : 0;
    const int i = task - op * N;
    const float4 acc = gather_row(U + 4 * j, g.ce, g.gstart[task], g.gstart[task + 1]);
    const int pb = op ? TC_PANEL_A : 0;
    store_split4(a_hi + pb, a_lo + pb, i, (op ? 0 : 32) + 4 * j, acc);
  }
  if (with_x) {
    for (int slot = s_beg + tid; slot < s_end; slot += TC_NT) {
      const int task = g.order[slot];
      const int op = task >= N ? 1 : 0;
      const int i = task - op * N;
      const float4 acc = gather_row(U + 32, g.ce, g.gstart[task], g.gstart[task + 1]);
      store_split4(a_hi + TC_PANEL_A, a_lo + TC_PANEL_A, i, op ? 40 : 36, acc);
    }
  }
}

template <int HALVES>
__global__ void __launch_bounds__(256 * HALVES, 1) k_dcrnn_seq_tc(const TcParams p) {
  constexpr int TC_NT = 256 * HALVES;
  constexpr int CW = 32 / HALVES;   // channels per thread
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, CIN = p.CIN, T = p.T;
  unsigned char* a_hi = smem + p.off_A;
  unsigned char* a_lo = a_hi + 2 * TC_PANEL_A;
  unsigned char* b_hi = smem + p.off_B;
  unsigned char* b_lo = b_hi + 2 * TC_PANEL_B;
  float* U = reinterpret_cast<float*>(smem + p.off_U);
  int* s_gstart = reinterpret_cast<int*>(smem + p.off_gstart);
  int* s_order = reinterpret_cast<int*>(smem + p.off_order);
  int2* s_ce = reinterpret_cast<int2*>(smem + p.off_ce);
  float* Bs = reinterpret_cast<float*>(smem + p.off_bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);   // [4]: gemm1 tile0/1, gemm2 tile0/1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  if (blockIdx.x >= p.B) return;

  // ---- one-time per CTA ------------------------------------------------------------------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // bars[0..3]: MMA completion (gemm1 tile0/1, gemm2 tile0/1); bars[4]: prologue TMA copies
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    const uint32_t tx = (p.gimg ? (uint32_t)p.gimg_bytes : 0u) + (p.wimage ? (uint32_t)TC_WIMAGE_BYTES : 0u);
    if (tx) {   // graph image and weight image arrive by TMA bulk copies while the CTA zeroes its panels
      mbar_arrive_expect_tx(&bars[4], tx);
      if (p.gimg) tma_bulk_g2s(smem + p.off_gstart, p.gimg, (uint32_t)p.gimg_bytes, &bars[4]);
      if (p.wimage) {
        tma_bulk_g2s(b_hi, p.wimage, 4u * TC_PANEL_B, &bars[4]);
        tma_bulk_g2s(Bs, reinterpret_cast<const unsigned char*>(p.wimage) + 4 * TC_PANEL_B, 96u * 4u, &bars[4]);
      }
    }
  }
  {  // zero A (pad columns / rows must be finite); B too unless the TMA image overwrites all of it
    uint4* z = reinterpret_cast<uint4*>(a_hi);
    const int nz = (4 * TC_PANEL_A + (p.wimage ? 0 : 4 * TC_PANEL_B)) / 16;
    for (int i = tid; i < nz; i += TC_NT) z[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < N * TC_UP; i += TC_NT) U[i] = 0.f;
  }
  if (!p.gimg) stage_graph<TC_NT>(p.rowptr[0], p.rowptr[1], p.cv[0], p.cv[1], N, TC_UP, s_ce, s_gstart, s_order, tid, 1 << 30, p.n_ops, 2);
  __syncthreads();
  if (!p.wimage) {
    // weights -> B operand (fp16 hi/lo, swizzled).  Row n = gate*32 + out channel; k order as the A panels.
    for (int idx = tid; idx < 96 * 112; idx += TC_NT) {
      const int n = idx / 112, kk = idx - n * 112;
      const float v = tc_weight_value(p.wcat, p.w[0], p.w[1], p.w[2], CIN, n, kk);
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      const int off = (kk >> 6) * TC_PANEL_B + sw128(n, kk & 63);
      *reinterpret_cast<__half*>(b_hi + off) = h;
      *reinterpret_cast<__half*>(b_lo + off) = l;
    }
    for (int idx = tid; idx < 96; idx += TC_NT) {
      const int gte = idx >> 5;
      const float* bg = gte == 0 ? p.bias[0] : (gte == 1 ? p.bias[1] : p.bias[2]);
      Bs[idx] = p.bcat ? p.bcat[idx] : (bg ? bg[idx & 31] : 0.f);
    }
  }
  if (p.gimg || p.wimage) mbar_wait(&bars[4], 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const GraphSmem gs{s_ce, s_gstart, s_order};
  // thread = (row, channel half): warp = half*8 + tile*4 + q ; TMEM lane == row, a warp may only touch lanes 32*(warp%4)..
  const int half = HALVES == 2 ? (warp >> 3) : 0, tile = (warp >> 2) & 1, q = warp & 3;
  const int row = tile * 128 + q * 32 + lane;
  const int ch0 = CW * half;
  const bool live = row < N;
  const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
  const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), b_hi_s = smem_u32(b_hi), b_lo_s = smem_u32(b_lo);
  constexpr uint32_t ID64 = umma_idesc_f16(128, 64), ID32 = umma_idesc_f16(128, 32);
  uint32_t parity = 0;
  const bool two_tiles = N > 128;

  // issue the 3 x 7 MMAs of one (tile, gemm) and commit them to `bar`          (one thread)
  auto issue = [&](int tl, int gm, uint64_t* bar) {
    const uint32_t dcol = gm == 0 ? 64u * tl : 128u + 32u * tl;
    uint32_t acc = 0;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {           // lo*hi, hi*lo, hi*hi (small terms first)
      const uint32_t ab = (pass == 0 ? a_lo_s : a_hi_s) + tl * (128 * 128);
      const uint32_t bb = (pass == 1 ? b_lo_s : b_hi_s) + (gm ? 64 * 128 : 0);
#pragma unroll
      for (int ks = 0; ks < 7; ++ks) {
        const int panel = ks >> 2, kin = (ks & 3) * 16;
        umma_f16(tmem + dcol, umma_desc(ab + panel * TC_PANEL_A + kin * 2), umma_desc(bb + panel * TC_PANEL_B + kin * 2),
                 gm == 0 ? ID64 : ID32, acc);
        acc = 1;
      }
    }
    umma_commit(bar);
  };

  auto x_base = [&](long long b) -> const float* { return p.x + (p.win_start ? p.win_start[b] * p.x_tstride : b * p.x_bstride); };

  for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
    const float* xb = x_base(b);
    // ---- window prologue: H_0 and X_0 into U (fp32) and the A panels (fp16 hi/lo) -------------------------------
    float hreg[CW];
    float xn[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
#pragma unroll
      for (int c = 0; c < CW / 4; ++c) {
        const float4 h = p.h0 ? __ldg(reinterpret_cast<const float4*>(p.h0 + b * p.h0_bstride + row * 32 + ch0) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        hreg[4 * c] = h.x; hreg[4 * c + 1] = h.y; hreg[4 * c + 2] = h.z; hreg[4 * c + 3] = h.w;
        st4(U + row * TC_UP + ch0 + 4 * c, h);
      }
      store_split_row<CW>(a_hi, a_lo, row, half, hreg);
      if (half == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < CIN) xn[c] = __ldg(xb + row * CIN + c);
        const float4 xv = make_float4(xn[0], xn[1], xn[2], xn[3]);
        st4(U + row * TC_UP + 32, xv);
        store_split4(a_hi + TC_PANEL_A, a_lo + TC_PANEL_A, row, 32, xv);
      }
    }
    __syncthreads();

    for (int t = 0; t < T; ++t) {
      // prefetch X_{t+1} (consumed at the end of the step)
      if (live && half == 0 && t + 1 < T) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < CIN) xn[c] = __ldg(xb + (t + 1) * p.x_tstride + row * CIN + c);
      }
      // ---- round 1: diffuse [H | X_t] -----------------------------------------------------------------------
      diffuse_tc<TC_NT>(U, gs, N, a_hi, a_lo, true, tid, 0, p.n_ops * N);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (tid == 0) { issue(0, 0, &bars[0]); if (two_tiles) issue(1, 0, &bars[1]); }
      // ---- epilogue 1: z, r gates; H*R ------------------------------------------------------------------------
      if (tile == 0 || two_tiles) mbar_wait(&bars[tile], parity);   // tile 1 has no rows when N <= 128
      tc_fence_after();
      const long long obase = (b * T + t) * (long long)N;
      {
        uint32_t vr[CW];
        tmem_ld<CW>(trow + 64 * tile + 32 + ch0, vr);
        tmem_ld_wait();
        float hr[CW];
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          const float r = sigmoid_fast(__uint_as_float(vr[c]) + Bs[32 + ch0 + c]);
          hr[c] = hreg[c] * r;
          vr[c] = __float_as_uint(r);
        }
        if (live) {
#pragma unroll
          for (int c = 0; c < CW / 4; ++c) st4(U + row * TC_UP + ch0 + 4 * c, make_float4(hr[4 * c], hr[4 * c + 1], hr[4 * c + 2], hr[4 * c + 3]));
          store_split_row<CW>(a_hi, a_lo, row, half, hr);
          if (p.stash) {
            float* sp = p.stash + ((obase * 3) + row) * 32 + ch0;
#pragma unroll
            for (int c = 0; c < CW / 4; ++c) {
              st4(sp + (long long)N * 32 + 4 * c, make_float4(__uint_as_float(vr[4 * c]), __uint_as_float(vr[4 * c + 1]),
                                                              __uint_as_float(vr[4 * c + 2]), __uint_as_float(vr[4 * c + 3])));
            }
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      // ---- round 2: re-diffuse the H*R columns -----------------------------------------------------------------
      diffuse_tc<TC_NT>(U, gs, N, a_hi, a_lo, false, tid, 0, p.n_ops * N);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (tid == 0) { issue(0, 1, &bars[2]); if (two_tiles) issue(1, 1, &bars[3]); }
      // ---- epilogue 2: candidate, H_t ----------------------------------------------------------------------------
      if (tile == 0 || two_tiles) mbar_wait(&bars[2 + tile], parity);
      tc_fence_after();
      {
        // Z is recomputed from its accumulator, which stays in TMEM until the next step's GEMM 1: cheaper than keeping
        // CW registers alive (and spilling) across the second diffusion round
        uint32_t vh[CW], vz[CW];
        tmem_ld<CW>(trow + 128 + 32 * tile + ch0, vh);
        tmem_ld<CW>(trow + 64 * tile + ch0, vz);
        tmem_ld_wait();
        float ht[CW], zreg[CW];
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          zreg[c] = sigmoid_fast(__uint_as_float(vz[c]) + Bs[ch0 + c]);
          ht[c] = tanh_fast(__uint_as_float(vh[c]) + Bs[64 + ch0 + c]);
          hreg[c] = zreg[c] * hreg[c] + (1.0f - zreg[c]) * ht[c];   // dcrnn.py:190-192
        }
        if (live) {
          float* op = p.out + (obase + row) * 32 + ch0;
#pragma unroll
          for (int c = 0; c < CW / 4; ++c) {
            const float4 hv = make_float4(hreg[4 * c], hreg[4 * c + 1], hreg[4 * c + 2], hreg[4 * c + 3]);
            st4(U + row * TC_UP + ch0 + 4 * c, hv);
            st4(op + 4 * c, hv);
          }
          store_split_row<CW>(a_hi, a_lo, row, half, hreg);
          if (p.stash) {
            float* sp = p.stash + ((obase * 3) + row) * 32 + ch0;
#pragma unroll
            for (int c = 0; c < CW / 4; ++c) {
              st4(sp + 4 * c, make_float4(zreg[4 * c], zreg[4 * c + 1], zreg[4 * c + 2], zreg[4 * c + 3]));
              st4(sp + 2 * (long long)N * 32 + 4 * c, make_float4(ht[4 * c], ht[4 * c + 1], ht[4 * c + 2], ht[4 * c + 3]));
            }
          }
          if (half == 0 && t + 1 < T) {
            const float4 xv = make_float4(xn[0], xn[1], xn[2], xn[3]);
            st4(U + row * TC_UP + 32, xv);
            store_split4(a_hi + TC_PANEL_A, a_lo + TC_PANEL_A, row, 32, xv);
          }
        }
      }
      parity ^= 1u;
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS));
}


inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

bool tc_layout(const stmp_plan* plan, TcParams* p, int* smem_bytes) {
  const int N = plan->n;
  int off = 0;
  p->off_A = off; off += 4 * TC_PANEL_A;
  p->off_B = off; off += 4 * TC_PANEL_B;
  p->off_U = off; off += align_up(N * TC_UP * 4, 16);
  p->off_gstart = off; off += align_up((2 * N + 1) * 4, 16);
  p->off_order = off; off += align_up(2 * N * 4, 16);
  const int nnz1 = plan->n_ops > 1 ? plan->fwd[1].nnz : 0;
  p->off_ce = off; off += align_up((plan->fwd[0].nnz + nnz1 + 6 * N + 4) * 8, 16);
  p->off_bias = off; off += 96 * 4;
  p->off_bar = off; off += 64;
  *smem_bytes = off;
  return off <= kMaxSmemTc;
}

}  // namespace

bool gru_tc_supported(const stmp_plan* plan, long long cin) {
  if (!plan || cin < 1 || cin > 4) return false;
  if (plan->n > TC_AROWS - 1 || plan->n < 1) return false;
  TcParams p;
  int smem = 0;
  p.N = plan->n;
  return tc_layout(plan, &p, &smem);
}

bool dcrnn_tc_supported(const stmp_plan* plan, long long cin, long long cout, long long K) {
  if (!plan || plan->flavor != STMP_FLAVOR_DCONV || plan->n_ops != 2) return false;
  if (K != 2 || cout != 32 || cin < 1 || cin > 4) return false;
  if (plan->n > TC_AROWS - 1 || plan->n < 1) return false;
  TcParams p;
  int smem = 0;
  p.N = plan->n;
  return tc_layout(plan, &p, &smem);
}

static int tc_launch_params(const stmp_plan* plan, TcParams& p, cudaStream_t st);

int dcrnn_tc_launch(const stmp_plan* plan, long long B, long long T, long long cin, const float* x, const long long* win_start,
                    long long x_bstride, long long x_tstride, const float* w_z, const float* w_r, const float* w_h, const float* b_z,
                    const float* b_r, const float* b_h, const float* h0, float* out, float* stash, const void* wimage,
                    cudaStream_t st) {
  TcParams p;
  p.N = plan->n; p.CIN = (int)cin; p.T = (int)T; p.B = B;
  p.x = x; p.win_start = win_start; p.x_bstride = x_bstride; p.x_tstride = x_tstride;
  p.w[0] = w_z; p.w[1] = w_r; p.w[2] = w_h; p.bias[0] = b_z; p.bias[1] = b_r; p.bias[2] = b_h;
  p.h0 = h0; p.h0_bstride = (long long)plan->n * 32; p.wcat = nullptr; p.bcat = nullptr; p.n_ops = 2;
  p.out = out; p.stash = stash; p.wimage = wimage;
  return tc_launch_params(plan, p, st);
}

// generic graph-GRU: prepacked weights, any plan flavor with >= n_ops operators
int gru_tc_launch(const stmp_plan* plan, int n_ops, long long B, long long T, long long cin, const float* x, const long long* win_start,
                  long long x_bstride, long long x_tstride, const float* wcat, const float* bcat, const float* h0, long long h0_bstride,
                  float* out, float* stash, const void* wimage, cudaStream_t st) {
  TcParams p;
  p.N = plan->n; p.CIN = (int)cin; p.T = (int)T; p.B = B;
  p.x = x; p.win_start = win_start; p.x_bstride = x_bstride; p.x_tstride = x_tstride;
  for (int i = 0; i < 3; ++i) { p.w[i] = nullptr; p.bias[i] = nullptr; }
  p.h0 = h0; p.h0_bstride = h0_bstride; p.wcat = wcat; p.bcat = bcat; p.n_ops = n_ops;
  p.out = out; p.stash = stash; p.wimage = wimage;
  return tc_launch_params(plan, p, st);
}

static int tc_launch_params(const stmp_plan* plan, TcParams& p, cudaStream_t st) {
  int smem = 0;
  const long long B = p.B;
  if (!tc_layout(plan, &p, &smem)) return set_error(STMP_EUNSUPPORTED, "tcgen05 graph-GRU kernel needs %d B of shared memory", smem);
  for (int op = 0; op < 2; ++op) {
    const int src = op < plan->n_ops ? op : 0;
    p.rowptr[op] = plan->fwd[src].rowptr; p.cv[op] = plan->fwd[src].cv;
  }
  p.gimg = (p.n_ops >= 1 && p.n_ops <= 2) ? plan->gimg[p.n_ops] : nullptr;
  p.gimg_bytes = p.gimg ? plan->gimg_bytes[p.n_ops] : 0;
  if (p.n_ops == 0) { p.gimg = nullptr; p.gimg_bytes = 0; }
  int dev = 0, sms = 0;
  STMP_CUDA_OK(cudaGetDevice(&dev));
  STMP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = (int)(B < sms ? B : sms);
  static int halves = -1;
  if (halves < 0) {
    const char* v = getenv("STMP_DCRNN_TC_HALVES");
    halves = (v && atoi(v) == 1) ? 1 : 2;   // 2 (default): 16 warps, two channel halves per row; 1: 8 warps
  }
  if (halves == 2) {
    STMP_CUDA_OK(cudaFuncSetAttribute(k_dcrnn_seq_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_dcrnn_seq_tc<2><<<grid, 512, smem, st>>>(p);
  } else {
    STMP_CUDA_OK(cudaFuncSetAttribute(k_dcrnn_seq_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_dcrnn_seq_tc<1><<<grid, 256, smem, st>>>(p);
  }
  STMP_LAUNCH_OK("k_dcrnn_seq_tc");
  return STMP_OK;
}

int tc_pack_weight_image(const float* wcat, const float* bcat, const float* w0, const float* w1, const float* w2, const float* b0,
                         const float* b1, const float* b2, int cin, void* image, cudaStream_t st) {
  const int total = 96 * 128 + 96;
  k_pack_weight_image<<<(total + 255) / 256, 256, 0, st>>>(wcat, bcat, w0, w1, w2, b0, b1, b2, cin, reinterpret_cast<unsigned char*>(image));
  STMP_LAUNCH_OK("k_pack_weight_image");
  return STMP_OK;
}
int tc_weight_image_bytes() { return TC_WIMAGE_BYTES; }

}  // namespace stmp
