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
// Shared memory (~204 KB of 227), all operands written by hand in the canonical K-major SWIZZLE_128B layout
// (row pitch 128 B = 64 fp16, 16-byte chunk index XOR (row % 8), 8-row atoms of 1024 B):
//   A_hi / A_lo : 2 K-panels x 208 rows   k = [H|H*R (32) , P_o H (32)] , [P_i H (32), X(4) P_oX(4) P_iX(4) 0(4)]
//   B_hi / B_lo : 2 K-panels x 96 rows    rows = output channels of z | r | h, same k order (7 k-steps of 16)
//   U  fp32 [208][36] = H (or H*R) + 16 B of row padding: the gather source of the diffusion (tensor cores only see the fp16 split); row 207 = 0
//   graph image (graph_image.cuh: balanced warp-task lists, 8-bit source rows four per word, values four per 128 bits),
//   biases, 6 mbarriers, 2 arrival counters.
// TMEM (256 columns): z|r accumulators of the two 128-row tiles at columns [0,64) [64,128), candidate at
// [128,160) [160,192).  TMEM lane == row, so thread (warp w, lane l) owns row 128*(w/4) + 32*(w%4) + l for the
// whole step: it reads its 64+32 accumulator columns with tcgen05.ld, applies the gates, keeps H in
// registers, and writes H*R / H_t back as fp32 (U), as fp16 hi/lo (A panel) and to HBM.  16 warps: each row is
// shared by two threads (channel halves), which also doubles the warps available to hide the gather latency.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "dcrnn_common.cuh"
#include "graph_image.cuh"
#include "tc_common.cuh"

#ifndef STMP_TC_GUNROLL
#define STMP_TC_GUNROLL 2
#endif
#ifndef STMP_TC_PREFETCH
#define STMP_TC_PREFETCH 0
#endif
#define STMP_TC_PRAGMA(x) _Pragma(#x)
#define STMP_TC_UNROLL(n) STMP_TC_PRAGMA(unroll n)

namespace stmp {
int g_fwd_split = 1;      // 1: a CTA pair per window for small batches (stmp_set_option("dcrnn_fwd_split")): 64 windows 160.8 -> 132.2 us
namespace {

constexpr int kMaxSmemTc = 232448;
constexpr int TC_UP = 36;                  // U row pitch (floats): a 128-byte row of H (or of T*Cin X values in the window prologue) + 16 B, so that the
                                           // epilogue's row-strided 16-byte stores rotate through the banks (pitch 32: 8-way conflicts, measured -14 %)
constexpr int TC_UROWS = 208;              // rows of U; row 207 is the all-zero row that pad entries of the graph image point at
constexpr int TC_AROWS = 208;              // rows stored per A panel (tile 1 over-reads into the next buffer: harmless)
constexpr int TC_PANEL_A = TC_AROWS * 128;
constexpr int TC_PANEL_B = 96 * 128;
constexpr int TC_TMEM_COLS = 256;

struct TcParams {
  int N, CIN, T;
  long long B;
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
  const void* gimg;       // plan's prebuilt shared-memory graph image for n_ops operators (TMA bulk source)
  GraphImageLayout gl;    // its internal offsets
  const void* wimage;     // prebuilt B-operand image (fp16 hi/lo, swizzled, + biases) or null
  float* out;
  float* stash;
  float* ws;              // workspace [grid][N][2 operators][ws_pitch] for P_o X / P_i X of a window's steps, or null (park them in `out`)
  int ws_pitch;
  int off_A, off_B, off_U, off_img, off_bias, off_bar;
};

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

// ---- gather ------------------------------------------------------------------------------------------------------------
// Weighted sum of the source rows of one task (graph_image.cuh): quarter-warp lane j accumulates floats 4j..4j+3 of the row.
// Four edges per group: one broadcast 32-bit load carries their four 8-bit source rows, one 128-bit load their values; the
// next group's entries are fetched while the current group's four feature rows are in flight.  Summation order = CSR order
// (the reference's scatter order), products by FMA as in the round-1 kernel.
#if STMP_IMG_OFF16
typedef uint2 img_idx_t;
#else
typedef uint32_t img_idx_t;
#endif
static_assert(kImgRowPitchBytes == TC_UP * 4, "graph image offsets are pre-scaled by the gather buffer's row pitch");
static_assert(kImgZeroRow * kImgRowPitchBytes < 65536, "row offsets must fit 16 bits");
__device__ __forceinline__ float4 ld4_off(const float* base, uint32_t byte_off) {
  return *reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(base) + byte_off);
}
// (u, v) = the task's first group, already loaded (gather_segment fetches it while the previous task is being gathered)
__device__ __forceinline__ float4 gather_groups(const float* __restrict__ Uj, const img_idx_t* __restrict__ idx4,
                                                const float4* __restrict__ val4, int g0, int ng, img_idx_t u, float4 v) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  STMP_TC_UNROLL(STMP_TC_GUNROLL)    // 2: +6 % over the rolled loop (A/B on one box: 886 k -> 942 k snapshots/s) -- a task has 2-3 groups on average,
  for (int g = 1; g <= ng; ++g) {    // the loop branch and its convergence barrier were 10 % of the issued instructions
    const img_idx_t un = idx4[g0 + g];     // (one spare group at the end of the arrays)
    const float4 vn = val4[g0 + g];
    // (predicating the loads of pad entries off saves their wavefronts but costs more in compares / selects: measured -2 %, A/B on one box)
#if STMP_IMG_OFF16
    const float4 x0 = ld4_off(Uj, u.x & 0xffffu);
    const float4 x1 = ld4_off(Uj, u.x >> 16);
    const float4 x2 = ld4_off(Uj, u.y & 0xffffu);
    const float4 x3 = ld4_off(Uj, u.y >> 16);
#else
    const float4 x0 = ld4(Uj + (u & 0xffu) * TC_UP);
    const float4 x1 = ld4(Uj + ((u >> 8) & 0xffu) * TC_UP);
    const float4 x2 = ld4(Uj + ((u >> 16) & 0xffu) * TC_UP);
    const float4 x3 = ld4(Uj + (u >> 24) * TC_UP);
#endif
    fma4(acc, v.x, x0);
    fma4(acc, v.y, x1);
    fma4(acc, v.z, x2);
    fma4(acc, v.w, x3);
    u = un;
    v = vn;
  }
  return acc;
}
__device__ __forceinline__ float4 gather_groups(const float* __restrict__ Uj, const img_idx_t* __restrict__ idx4,
                                                const float4* __restrict__ val4, int g0, int ng) {
  return gather_groups(Uj, idx4, val4, g0, ng, idx4[g0], val4[g0]);
}

// Step anatomy (all 16 warps; T_k = MMA row tile k = rows [128k, 128k+128)):
//   round 1  tid 0 issues the H | X k-steps of GEMM1 for both tiles; all warps gather [P_o H | P_i H] of T_0's rows -> A panels
//            barrier; tid 0 issues T_0's P_o / P_i k-steps + commit -- the tensor core works on T_0 while the LSU gathers T_1's rows
//            barrier; tid 0 issues T_1's k-steps + commit
//   epi 1    wait GEMM1(own tile): R, H*R -> U, A panel                                                          barrier
//   round 2  same gathers over H*R, GEMM2 (candidate)
//   epi 2    wait GEMM2(own tile): Z (recomputed from TMEM), H~, H_t -> U, A panel, HBM;  X_{t+1} k-step           barrier
// The task lists of the gather are balanced over the warps when the plan is built (graph_image.cuh), which is what makes the extra
// barriers cheap (round 1 profile: 25 % of all warp time was barrier wait behind the warp that always drew the longest rows).
// X is never gathered per step: P_o X_t, P_i X_t of ALL steps of a window are produced by one gather pass over rows of
// T*Cin floats in the window prologue and parked in the window's own (not yet written) output rows out[b, t, :, 0:8].
// SPLIT = 2 (small batches: 2 B CTAs still fit the machine, N > 128): a window is served by a 2-CTA thread-block cluster.  CTA c owns MMA
// row tile c -- its gather tasks, its MMAs, its epilogue (all 16 warps: thread = (row, channel quarter)) -- and pushes the rows of H*R / H_t it
// produces into the partner's gather buffer U through distributed shared memory, so both gathers stay local.  Per round: "done reading U"
// is a relaxed cluster arrival right after the gather, waited for just before the epilogue overwrites U; the barrier that closes an epilogue is
// a release / acquire cluster barrier (the pushed rows are visible).
template <int CIN, int SPLIT>
__global__ void __launch_bounds__(512, 1) k_dcrnn_seq_tc(const TcParams p) {
  constexpr int CW = SPLIT == 2 ? 8 : 16;   // channels per thread (two / four threads share a row)
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, T = p.T;
  unsigned char* a_hi = smem + p.off_A;
  unsigned char* a_lo = a_hi + 2 * TC_PANEL_A;
  unsigned char* b_hi = smem + p.off_B;
  unsigned char* b_lo = b_hi + 2 * TC_PANEL_B;
  float* U = reinterpret_cast<float*>(smem + p.off_U);
  const unsigned char* img = smem + p.off_img;
  const uint16_t* s_wstart = reinterpret_cast<const uint16_t*>(img + p.gl.off_wstart);
  const uint16_t* s_wcount = reinterpret_cast<const uint16_t*>(img + p.gl.off_wcount);
  const uint32_t* s_wt = reinterpret_cast<const uint32_t*>(img + p.gl.off_wt);
  const img_idx_t* s_idx = reinterpret_cast<const img_idx_t*>(img + p.gl.off_idx);
  const float4* s_val = reinterpret_cast<const float4*>(img + p.gl.off_val);
  float* Bs = reinterpret_cast<float*>(smem + p.off_bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);   // [0..3] MMA done (gemm*2+tile), [4] prologue TMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int crank = SPLIT == 2 ? (int)cluster_rank() : 0;
  const long long cta = blockIdx.x / SPLIT, n_cta = gridDim.x / SPLIT;
  if (cta >= p.B) return;

  // ---- one-time per CTA ------------------------------------------------------------------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    const uint32_t tx = (p.n_ops ? (uint32_t)p.gl.bytes : 0u) + (p.wimage ? (uint32_t)TC_WIMAGE_BYTES : 0u);
    if (tx) {   // graph image and weight image arrive by TMA bulk copies while the CTA zeroes its panels
      mbar_arrive_expect_tx(&bars[4], tx);
      if (p.n_ops) tma_bulk_g2s(smem + p.off_img, p.gimg, (uint32_t)p.gl.bytes, &bars[4]);
      if (p.wimage) {
        tma_bulk_g2s(b_hi, p.wimage, 4u * TC_PANEL_B, &bars[4]);
        tma_bulk_g2s(Bs, reinterpret_cast<const unsigned char*>(p.wimage) + 4 * TC_PANEL_B, 96u * 4u, &bars[4]);
      }
    }
  }
  {  // zero A (pad columns / rows must be finite) and U (row 207 stays the zero row); B too unless the TMA image overwrites all of it
    uint4* z = reinterpret_cast<uint4*>(a_hi);
    const int nz = (4 * TC_PANEL_A + (p.wimage ? 0 : 4 * TC_PANEL_B)) / 16;
    for (int i = tid; i < nz; i += 512) z[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < TC_UROWS * TC_UP; i += 512) U[i] = 0.f;
  }
  __syncthreads();
  if (!p.wimage) {
    // weights -> B operand (fp16 hi/lo, swizzled).  Row n = gate*32 + out channel; k order as the A panels.
    for (int idx = tid; idx < 96 * 112; idx += 512) {
      const int n = idx / 112, kk = idx - n * 112;
      const float v = tc_weight_value(p.wcat, p.w[0], p.w[1], p.w[2], CIN, n, kk);
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      const int off = (kk >> 6) * TC_PANEL_B + sw128(n, kk & 63);
      *reinterpret_cast<__half*>(b_hi + off) = h;
      *reinterpret_cast<__half*>(b_lo + off) = l;
    }
    for (int idx = tid; idx < 96; idx += 512) {
      const int gte = idx >> 5;
      const float* bg = gte == 0 ? p.bias[0] : (gte == 1 ? p.bias[1] : p.bias[2]);
      Bs[idx] = p.bcat ? p.bcat[idx] : (bg ? bg[idx & 31] : 0.f);
    }
  }
  if (p.n_ops || p.wimage) mbar_wait(&bars[4], 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // thread = (row, channel half): warp = half*8 + tile*4 + q ; TMEM lane == row, a warp may only touch lanes 32*(warp%4)..
  // (cluster pair: thread = (row of my tile, channel quarter): warp = quarter*4 + q)
  const int half = SPLIT == 2 ? (warp >> 2) : (warp >> 3), tile = SPLIT == 2 ? crank : ((warp >> 2) & 1), q = warp & 3;
  const int row = tile * 128 + q * 32 + lane;
  const int ch0 = CW * half;
  const bool live = row < N;
  const bool owner = live && half == 0;      // the thread that feeds its row's X k-step
  const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
  const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), b_hi_s = smem_u32(b_hi), b_lo_s = smem_u32(b_lo);
  constexpr uint32_t ID64 = umma_idesc_f16(128, 64), ID32 = umma_idesc_f16(128, 32);
  uint32_t parity = 0;
  const bool two_tiles = N > 128;
  const int j = lane & 7, quarter = lane >> 3;
  const float* Uj = U + 4 * j;
  uint32_t peer_U = 0;
  if constexpr (SPLIT == 2) peer_U = map_to_peer(U, (uint32_t)(crank ^ 1));
  // store a float4 of my row into U -- and into the partner's U in the cluster-pair variant
  auto put_u = [&](int off, float4 v) {
    st4(U + off, v);
    if constexpr (SPLIT == 2) st4_cluster(peer_U + (uint32_t)off * 4u, v);
  };
  // barrier that closes an epilogue / the window prologue: operand + U stores of everybody visible to the MMAs and the next gather
  auto close_phase = [&]() {
    fence_proxy_async();
    tc_fence_before();
    if constexpr (SPLIT == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
  };

  // The 3 x 7 MMAs of one (tile, gemm) are issued in three groups, each as soon as its k-steps are in shared memory, each with
  // its own commit to the (tile, gemm) barrier (count 3):
  //   group 0: k-steps of H | H*R and X  -- complete when the round starts; its first MMA overwrites the accumulator
  //   group 1: k-steps of P_o H          -- after the last warp finished the tile's P_o tasks
  //   group 2: k-steps of P_i H          -- after the last warp finished the tile's P_i tasks
  // so the tensor core runs underneath the gather and only the last group's 6 MMAs are exposed at the end of a round.
  auto issue_group = [&](int tl, int gm, int grp) {
    const uint32_t dcol = gm == 0 ? 64u * tl : 128u + 32u * tl;
    const int ks0 = grp == 0 ? 0 : (grp == 1 ? 2 : 4);
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {           // lo*hi, hi*lo, hi*hi (small terms first)
      const uint32_t ab = (pass == 0 ? a_lo_s : a_hi_s) + tl * (128 * 128);
      const uint32_t bb = (pass == 1 ? b_lo_s : b_hi_s) + (gm ? 64 * 128 : 0);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (i == 2 && grp != 0) continue;
        const int ks = i == 2 ? 6 : ks0 + i;
        const int panel = ks >> 2, kin = (ks & 3) * 16;
        umma_f16(tmem + dcol, umma_desc(ab + panel * TC_PANEL_A + kin * 2), umma_desc(bb + panel * TC_PANEL_B + kin * 2),
                 gm == 0 ? ID64 : ID32, (grp == 0 && pass == 0 && i == 0) ? 0u : 1u);
      }
    }
  };

  // this warp's warp-tasks of segment `seg` = (MMA tile of the destination rows) * 2 + operator: results -> A panels as fp16 hi/lo
  //   op 0 (P_o): panel 0 k 32..63        op 1 (P_i): panel 1 k 0..31
  auto gather_segment = [&](int seg) {
    const int ws = s_wstart[warp * 4 + seg], wc = s_wcount[warp * 4 + seg];
    const int op = seg & 1;
    unsigned char* dh = a_hi + (op ? TC_PANEL_A : 0);
    unsigned char* dl = a_lo + (op ? TC_PANEL_A : 0);
    const int kcol = (op ? 0 : 32) + 4 * j;
#if STMP_TC_PREFETCH
    // the next task's descriptor and first edge group are fetched while the current task is gathered: the chain descriptor -> group ->
    // feature rows (three dependent shared-memory latencies per task) is taken off the critical path
    uint32_t d = wc > 0 ? s_wt[ws * 4 + quarter] : kImgNoTask;
    int g0 = d == kImgNoTask ? 0 : (int)(d >> 16);
    img_idx_t u0 = s_idx[g0];
    float4 v0 = s_val[g0];
    for (int i = 0; i < wc; ++i) {
      const uint32_t dn = i + 1 < wc ? s_wt[(ws + i + 1) * 4 + quarter] : kImgNoTask;
      const int g0n = dn == kImgNoTask ? 0 : (int)(dn >> 16);
      const img_idx_t un0 = s_idx[g0n];
      const float4 vn0 = s_val[g0n];
      if (d != kImgNoTask) {
        const float4 acc = gather_groups(Uj, s_idx, s_val, g0, (int)((d >> 9) & 0x7f), u0, v0);
        store_split4(dh, dl, (int)(d & 0xff), kcol, acc);
      }
      d = dn; g0 = g0n; u0 = un0; v0 = vn0;
    }
#else
    for (int i = 0; i < wc; ++i) {
      const uint32_t d = s_wt[(ws + i) * 4 + quarter];
      if (d != kImgNoTask) {
        const float4 acc = gather_groups(Uj, s_idx, s_val, (int)(d >> 16), (int)((d >> 9) & 0x7f));
        store_split4(dh, dl, (int)(d & 0xff), kcol, acc);
      }
    }
#endif
  };
  // One gather round.  tid 0 issues every MMA: the static group (H | X k-steps) of both tiles right away (the block barrier in front of
  // the round ordered those operand stores), tile 0's P_o / P_i groups behind the barrier that closes tile 0's tasks -- they run on the
  // tensor core while the LSU gathers tile 1 -- and tile 1's groups behind the closing barrier.  One commit per tile covers all of
  // that tile's MMAs (same issuing thread => in order).  The task lists are balanced (graph_image.cuh), so the two barriers are cheap;
  // the closing one also tells the epilogue that nobody reads U any more.
  auto gather_round = [&](int gm) {
    if constexpr (SPLIT == 2) {
      if (tid == 0) issue_group(crank, gm, 0);
      if (p.n_ops > 0) gather_segment(2 * crank);
      if (p.n_ops > 1) gather_segment(2 * crank + 1);
      cluster_arrive_relaxed();   // this CTA has finished reading U (waited for by the partner before its epilogue overwrites my U)
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (tid == 0) {
        issue_group(crank, gm, 1);
        issue_group(crank, gm, 2);
        umma_commit(&bars[2 * gm + crank]);
      }
      return;
    }
    if (tid == 0) {
      issue_group(0, gm, 0);
      if (two_tiles) issue_group(1, gm, 0);
    }
    if (p.n_ops > 0) gather_segment(0);
    if (p.n_ops > 1) gather_segment(1);
    fence_proxy_async();        // my generic-proxy stores to the A panels -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      issue_group(0, gm, 1);
      issue_group(0, gm, 2);
      umma_commit(&bars[2 * gm]);
    }
    if (two_tiles) {
      if (p.n_ops > 0) gather_segment(2);
      if (p.n_ops > 1) gather_segment(3);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0 && two_tiles) {
      issue_group(1, gm, 1);
      issue_group(1, gm, 2);
      umma_commit(&bars[2 * gm + 1]);
    }
  };

  auto x_base = [&](long long b) -> const float* { return p.x + (p.win_start ? p.win_start[b] * p.x_tstride : b * p.x_bstride); };
  // [X_t | P_o X_t | P_i X_t] of my row -> k 32..43 of panel 1 (weights of absent channels / operators are zero, but the
  // operand itself must be finite: everything not produced is written as 0 -- at store time, so the loads stay in flight)
  auto load_x = [&](const float* xb, long long b, int t, float4& xv, float4& po, float4& pi) {
    float xn[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < CIN; ++c) xn[c] = __ldg(xb + t * p.x_tstride + row * CIN + c);
    xv = make_float4(xn[0], xn[1], xn[2], xn[3]);
    po = pi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.ws) {             // per-CTA workspace rows [row][op][t*CIN + c]: rewritten every window, so they stay in L2
      const float* w0 = p.ws + (((long long)blockIdx.x * N + row) * 2) * p.ws_pitch + t * CIN;
      float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        if (p.n_ops >= 1) a0[c] = w0[c];
        if (p.n_ops >= 2) a1[c] = w0[p.ws_pitch + c];
      }
      po = make_float4(a0[0], a0[1], a0[2], a0[3]);
      pi = make_float4(a1[0], a1[1], a1[2], a1[3]);
      return;
    }
    const float* sc = p.out + ((b * T + t) * (long long)N + row) * 32;   // parked there by the window prologue (plain loads:
    if (p.n_ops >= 1) po = *reinterpret_cast<const float4*>(sc);         //  written by this CTA earlier in this launch)
    if (p.n_ops >= 2) pi = *reinterpret_cast<const float4*>(sc + 4);
  };
  auto mask_c = [&](float4 v) {
    if (CIN < 4) v.w = 0.f;
    if (CIN < 3) v.z = 0.f;
    if (CIN < 2) v.y = 0.f;
    return v;
  };
  auto store_x = [&](const float4& xv, const float4& po, const float4& pi) {
    store_split4(a_hi + TC_PANEL_A, a_lo + TC_PANEL_A, row, 32, xv);
    store_split4(a_hi + TC_PANEL_A, a_lo + TC_PANEL_A, row, 36, mask_c(po));
    store_split4(a_hi + TC_PANEL_A, a_lo + TC_PANEL_A, row, 40, mask_c(pi));
  };

  for (long long b = cta; b < p.B; b += n_cta) {
    const float* xb = x_base(b);
    // ---- window prologue A: P_o X_t, P_i X_t for every step of the window, TCH steps per gather pass ------------------------
    if (p.n_ops) {
      constexpr int TCH = 32 / CIN;
      for (int t0 = 0; t0 < T; t0 += TCH) {
        const int tn = (T - t0) < TCH ? (T - t0) : TCH;
        const int F = tn * CIN, NC = N * CIN;
        // U[n][tt*CIN + c] = X[b, t0+tt, n, c]: all loads of a thread are issued before the first store (one HBM/L2 latency)
        for (int base = 0; base < tn * NC; base += 512 * 8) {
          float xv8[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int idx = base + u * 512 + tid;
            xv8[u] = 0.f;
            if (idx < tn * NC) {
              const int tt = idx / NC, r = idx - tt * NC;
              xv8[u] = __ldg(xb + (long long)(t0 + tt) * p.x_tstride + r);
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int idx = base + u * 512 + tid;
            if (idx < tn * NC) {
              const int tt = idx / NC, r = idx - tt * NC;
              const int n = r / CIN, c = r - n * CIN;
              U[n * TC_UP + tt * CIN + c] = xv8[u];
            }
          }
        }
        __syncthreads();
        for (int seg = SPLIT == 2 ? 2 * crank : 0; seg < (SPLIT == 2 ? 2 * crank + 2 : 4); ++seg) {
          const int ws = s_wstart[warp * 4 + seg], wc = s_wcount[warp * 4 + seg];
          for (int i = 0; i < wc; ++i) {
            const uint32_t d = s_wt[(ws + i) * 4 + quarter];
            if (d != kImgNoTask && 4 * j < F) {
              const float4 acc = gather_groups(Uj, s_idx, s_val, (int)(d >> 16), (int)((d >> 9) & 0x7f));
              const float av[4] = {acc.x, acc.y, acc.z, acc.w};
              const int drow = d & 0xff, op = (d >> 8) & 1;
              float* orow = p.out + ((b * T + t0) * (long long)N + drow) * 32 + op * 4;
              const long long tstep = (long long)N * 32;
              if (p.ws) {               // floats 4j..4j+3 of the row are consecutive (t, c) pairs: one 16-byte store, full sectors
                float* wrow = p.ws + (((long long)blockIdx.x * N + drow) * 2 + op) * p.ws_pitch + t0 * CIN + 4 * j;
                if (4 * j + 4 <= F) {
                  *reinterpret_cast<float4*>(wrow) = acc;
                } else {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (4 * j + k < F) wrow[k] = av[k];
                }
              } else if (CIN == 2) {           // floats 4j..4j+3 = (t, c) = (2j,0) (2j,1) (2j+1,0) (2j+1,1): two 8-byte stores
                *reinterpret_cast<float2*>(orow + (2 * j) * tstep) = make_float2(av[0], av[1]);
                if (4 * j + 2 < F) *reinterpret_cast<float2*>(orow + (2 * j + 1) * tstep) = make_float2(av[2], av[3]);
              } else if (CIN == 4) {    // one timestep per lane: a 16-byte store
                *reinterpret_cast<float4*>(orow + j * tstep) = acc;
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int f = 4 * j + k;
                  if (f < F) {
                    const int tt = f / CIN, c = f - tt * CIN;
                    orow[tt * tstep + c] = av[k];
                  }
                }
              }
            }
          }
        }
        __syncthreads();
      }
    }
    // ---- window prologue B: H_0 into U (fp32) and the A panels (fp16 hi/lo); the X k-step of step 0 ---------------------------
    float hreg[CW];
    if constexpr (SPLIT == 2) {   // the partner has finished the X gather of its prologue: its U may be overwritten
      cluster_arrive_relaxed();
      cluster_wait();
    }
    if (live) {
#pragma unroll
      for (int c = 0; c < CW / 4; ++c) {
        const float4 h = p.h0 ? __ldg(reinterpret_cast<const float4*>(p.h0 + b * p.h0_bstride + row * 32 + ch0) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        hreg[4 * c] = h.x; hreg[4 * c + 1] = h.y; hreg[4 * c + 2] = h.z; hreg[4 * c + 3] = h.w;
        put_u(row * TC_UP + ch0 + 4 * c, h);
      }
      store_split_row<CW>(a_hi, a_lo, row, half, hreg);
      if (owner) {
        float4 xv, po, pi;
        load_x(xb, b, 0, xv, po, pi);
        store_x(xv, po, pi);
      }
    }
    close_phase();             // the first MMA group of step 0 is issued right behind this barrier

    for (int t = 0; t < T; ++t) {
      // ---- round 1: diffuse H ------------------------------------------------------------------------------------------
      gather_round(0);
      // ---- epilogue 1: r gate; H*R ---------------------------------------------------------------------------------------
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
        if constexpr (SPLIT == 2) cluster_wait();       // the partner is done gathering from its U
        if (live) {
#pragma unroll
          for (int c = 0; c < CW / 4; ++c) put_u(row * TC_UP + ch0 + 4 * c, make_float4(hr[4 * c], hr[4 * c + 1], hr[4 * c + 2], hr[4 * c + 3]));
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
      close_phase();
      // ---- round 2: re-diffuse H*R ---------------------------------------------------------------------------------------
      gather_round(1);
      // ---- epilogue 2: candidate, H_t --------------------------------------------------------------------------------------
      float4 xv, po, pi;
      const bool feed_x = owner && t + 1 < T;
      if (feed_x) load_x(xb, b, t + 1, xv, po, pi);      // in flight under the MMA wait
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
        if constexpr (SPLIT == 2) cluster_wait();
        if (live) {
          float* op = p.out + (obase + row) * 32 + ch0;
#pragma unroll
          for (int c = 0; c < CW / 4; ++c) {
            const float4 hv = make_float4(hreg[4 * c], hreg[4 * c + 1], hreg[4 * c + 2], hreg[4 * c + 3]);
            put_u(row * TC_UP + ch0 + 4 * c, hv);
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
          if (feed_x) store_x(xv, po, pi);
        }
      }
      parity ^= 1u;
      close_phase();
    }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS));
}


inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// shared-memory layout for `n_ops` operators of `plan`
bool tc_layout(const stmp_plan* plan, int n_ops, TcParams* p, int* smem_bytes) {
  int off = 0;
  p->off_A = off; off += 4 * TC_PANEL_A;
  p->off_B = off; off += 4 * TC_PANEL_B;
  p->off_U = off; off += TC_UROWS * TC_UP * 4;
  int nnz = 0;
  for (int op = 0; op < n_ops; ++op) nnz += plan->fwd[op].nnz;
  p->gl = graph_image_layout(n_ops * plan->n, nnz);
  p->off_img = off; off += n_ops ? p->gl.bytes : 0;
  p->off_bias = off; off += 96 * 4;
  p->off_bar = off; off += 64;
  *smem_bytes = off;
  return off <= kMaxSmemTc;
}

}  // namespace

int tc_ws_pitch(long long T, long long cin) { return (int)((T * cin + 7) / 8 * 8); }   // floats per (row, operator): whole 32-byte sectors

long long tc_workspace_bytes(const stmp_plan* plan, long long T, long long cin) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (long long)sms * plan->n * 2 * tc_ws_pitch(T, cin) * 4;
}

static bool tc_fits(const stmp_plan* plan, int n_ops) {
  if (!plan || plan->n > kImgMaxN || plan->n < 1) return false;
  if (n_ops > 0 && !plan->gimg[n_ops]) return false;      // no image: graph too large / too dense for the 8-bit format
  TcParams p;
  int smem = 0;
  return tc_layout(plan, n_ops, &p, &smem);
}

bool gru_tc_supported(const stmp_plan* plan, long long cin, int n_ops) {
  if (!plan || cin < 1 || cin > 4 || n_ops < 0 || n_ops > 2 || n_ops > plan->n_ops) return false;
  return tc_fits(plan, n_ops);
}

bool dcrnn_tc_supported(const stmp_plan* plan, long long cin, long long cout, long long K) {
  if (!plan || plan->flavor != STMP_FLAVOR_DCONV || plan->n_ops != 2) return false;
  if (K != 2 || cout != 32 || cin < 1 || cin > 4) return false;
  return tc_fits(plan, 2);
}

static int tc_launch_params(const stmp_plan* plan, TcParams& p, cudaStream_t st);

int dcrnn_tc_launch(const stmp_plan* plan, long long B, long long T, long long cin, const float* x, const long long* win_start,
                    long long x_bstride, long long x_tstride, const float* w_z, const float* w_r, const float* w_h, const float* b_z,
                    const float* b_r, const float* b_h, const float* h0, float* out, float* stash, const void* wimage,
                    void* workspace, cudaStream_t st) {
  TcParams p;
  p.N = plan->n; p.CIN = (int)cin; p.T = (int)T; p.B = B;
  p.x = x; p.win_start = win_start; p.x_bstride = x_bstride; p.x_tstride = x_tstride;
  p.w[0] = w_z; p.w[1] = w_r; p.w[2] = w_h; p.bias[0] = b_z; p.bias[1] = b_r; p.bias[2] = b_h;
  p.h0 = h0; p.h0_bstride = (long long)plan->n * 32; p.wcat = nullptr; p.bcat = nullptr; p.n_ops = 2;
  p.out = out; p.stash = stash; p.wimage = wimage; p.ws = reinterpret_cast<float*>(workspace);
  return tc_launch_params(plan, p, st);
}

// generic graph-GRU: prepacked weights, any plan flavor with >= n_ops operators
int gru_tc_launch(const stmp_plan* plan, int n_ops, long long B, long long T, long long cin, const float* x, const long long* win_start,
                  long long x_bstride, long long x_tstride, const float* wcat, const float* bcat, const float* h0, long long h0_bstride,
                  float* out, float* stash, const void* wimage, void* workspace, cudaStream_t st) {
  TcParams p;
  p.N = plan->n; p.CIN = (int)cin; p.T = (int)T; p.B = B;
  p.x = x; p.win_start = win_start; p.x_bstride = x_bstride; p.x_tstride = x_tstride;
  for (int i = 0; i < 3; ++i) { p.w[i] = nullptr; p.bias[i] = nullptr; }
  p.h0 = h0; p.h0_bstride = h0_bstride; p.wcat = wcat; p.bcat = bcat; p.n_ops = n_ops;
  p.out = out; p.stash = stash; p.wimage = wimage; p.ws = reinterpret_cast<float*>(workspace);
  return tc_launch_params(plan, p, st);
}

static int tc_launch_params(const stmp_plan* plan, TcParams& p, cudaStream_t st) {
  int smem = 0;
  const long long B = p.B;
  if (!tc_layout(plan, p.n_ops, &p, &smem)) return set_error(STMP_EUNSUPPORTED, "tcgen05 graph-GRU kernel needs %d B of shared memory", smem);
  p.gimg = p.n_ops ? plan->gimg[p.n_ops] : nullptr;
  if (p.n_ops && !p.gimg) return set_error(STMP_EUNSUPPORTED, "tcgen05 graph-GRU kernel: the plan has no shared-memory graph image");
  int dev = 0, sms = 0;
  STMP_CUDA_OK(cudaGetDevice(&dev));
  STMP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // small batches: a CTA pair (thread-block cluster) per window when both tiles exist and 2 B CTAs fit the machine
  const bool split = g_fwd_split != 0 && plan->n > 128 && 2 * B <= sms;
  const int grid = split ? (int)(2 * B) : (int)(B < sms ? B : sms);
  p.ws_pitch = tc_ws_pitch(p.T, p.CIN);
  switch (p.CIN) {
#define STMP_TC_CASE(C)                                                                                              \
  case C:                                                                                                            \
    if (split) {                                                                                                     \
      STMP_CUDA_OK(cudaFuncSetAttribute(k_dcrnn_seq_tc<C, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   \
      cudaLaunchConfig_t cfg = {};                                                                                   \
      cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st; \
      cudaLaunchAttribute at[1];                                                                                     \
      at[0].id = cudaLaunchAttributeClusterDimension;                                                                \
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;                            \
      cfg.attrs = at; cfg.numAttrs = 1;                                                                              \
      STMP_CUDA_OK(cudaLaunchKernelEx(&cfg, k_dcrnn_seq_tc<C, 2>, p));                                               \
    } else {                                                                                                         \
      STMP_CUDA_OK(cudaFuncSetAttribute(k_dcrnn_seq_tc<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   \
      k_dcrnn_seq_tc<C, 1><<<grid, 512, smem, st>>>(p);                                                              \
    }                                                                                                                \
    break;
    STMP_TC_CASE(1) STMP_TC_CASE(2) STMP_TC_CASE(3) STMP_TC_CASE(4)
#undef STMP_TC_CASE
    default: return set_error(STMP_EUNSUPPORTED, "tcgen05 graph-GRU kernel: cin %d not in 1..4", p.CIN);
  }
  STMP_LAUNCH_OK("k_dcrnn_seq_tc");
  if (split) { static const int slot2 = path_slot("k_dcrnn_seq_tc[cluster2]"); count_path(slot2); }
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
