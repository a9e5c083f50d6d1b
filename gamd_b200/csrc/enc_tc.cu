// Edge featurisation + 3-layer GELU edge encoder + LayerNorm on the tensor cores (K5 + K6).
//
// Restates calc_edge_feat (code/nn_module.py:603-634: unit vector, standardised distance, 40-centre RBF,
// optional bond flag) and edge_layer_norm(edge_encoder(feat)) (:598-600, :646) for 128-edge tiles:
//   A0 = features [128 x 64 (44|45 zero-padded)]  --tcgen05-->  GELU  -->  GELU  -->  + LayerNorm
// The A operand of every GEMM is written to TMEM by the epilogue threads (thread = edge row), the three
// weight matrices stay resident in shared memory as SWIZZLE_128B K-major bf16 images (hi and lo parts),
// accumulators live in TMEM.  Output: the bf16 hi/lo "blob" layout the message-passing kernel consumes
// ([tile][hi|lo][16 k-chunks][128 rows][16 B]).  Same CTA organisation as mp_tc.cu: 16 epilogue warps
// (2 tiles in flight x 4 TMEM lane quadrants x 2 column halves), one MMA-issue warp, one weight-loader warp.
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int TILE = 128;
constexpr int EPI_WARPS = 16;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int W0 = 16384, W1 = 32768;                 // bytes of one part of enc0 (K=64) / enc2, enc4 (K=128)
constexpr int OFF_ENC0 = 0, OFF_ENC2 = 2 * W0, OFF_ENC4 = 2 * W0 + 2 * W1, W_TOTAL = 2 * W0 + 4 * W1;

struct __align__(1024) SmemEnc {
  uint8_t w[W_TOTAL];          // enc0 hi, enc0 lo, enc2 hi, enc2 lo, enc4 hi, enc4 lo
  float bias[3][128];
  float ln_w[128], ln_b[128];
  float centers[GAMD_NRBF];
  float xch[2][128][2];        // LayerNorm partial sums exchanged between the two column-half warps of a row
  uint64_t w_full, a_ready[2], d_ready[2];
  uint32_t tmem_base;
};

struct EncTcArgs {
  const uint8_t* w_img;        // W_TOTAL bytes
  const float *bias, *ln_w, *ln_b, *centers;
  const float4* pos;           // feature positions (wrapped), .w unused here
  const int *col, *edst, *n_edges, *orig_id, *bond;
  uint8_t* e_blob;
  float length_mean, length_std;
  float box[3];
  int n_edge_in, use_bond, expand_edge, atoms_per_frame, exact;
  int dynbox;     // WaterMDDynamicBoxNet: rel = -(min-image of pos[center] - pos[neigh])  (nn_module.py:327)
};

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exact-erf GELU, erf by Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7): 2 MUFU + ~11 FMA-pipe ops
__device__ __forceinline__ float gelu_as(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p, e, 1.f);          // erf(|x|/sqrt2)
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs, hx);             // 0.5 x (1 + sign(x) erf_abs)
}
// the same for two elements on packed fp32x2 arithmetic (FMUL2 / FFMA2): 12 FMA-pipe + 4 ALU + 4 MUFU instructions per
// PAIR instead of ~15 + 2 per element - the encoder epilogues are issue-bound (ncu: 8 k issue cycles of 14 k per tile)
__device__ __forceinline__ f32x2 gelu_as2(f32x2 X) {
  const f32x2 W = mul2(X, X);                                             // x^2
  float x0, x1;
  unpk2(X, x0, x1);
  const f32x2 Z = mul2(pk2(fabsf(x0), fabsf(x1)), pk2(0.70710678118654752440f, 0.70710678118654752440f));
  const f32x2 D = fma2(Z, pk2(0.3275911f, 0.3275911f), pk2(1.f, 1.f));
  float d0, d1;
  unpk2(D, d0, d1);
  const f32x2 T = pk2(rcp_approx(d0), rcp_approx(d1));
  f32x2 P = fma2(T, pk2(1.061405429f, 1.061405429f), pk2(-1.453152027f, -1.453152027f));
  P = fma2(P, T, pk2(1.421413741f, 1.421413741f));
  P = fma2(P, T, pk2(-0.284496736f, -0.284496736f));
  P = fma2(P, T, pk2(0.254829592f, 0.254829592f));
  P = mul2(P, T);
  // exp(-z^2) = 2^(-log2(e)/2 * x^2)
  const f32x2 A = mul2(W, pk2(-0.72134752044448170368f, -0.72134752044448170368f));
  float a0, a1;
  unpk2(A, a0, a1);
  const f32x2 E = pk2(ex2_approx(a0), ex2_approx(a1));
  const f32x2 ERF = fma2(mul2(P, pk2(-1.f, -1.f)), E, pk2(1.f, 1.f));       // erf(|x| / sqrt 2)
  const f32x2 HX = mul2(X, pk2(0.5f, 0.5f));
  float h0, h1;
  unpk2(HX, h0, h1);
  return fma2(pk2(fabsf(h0), fabsf(h1)), ERF, HX);                         // 0.5 x (1 + sign(x) erf)
}
// split two fp32 values (packed) into bf16 hi / lo pairs
__device__ __forceinline__ void split2(f32x2 Y, uint32_t& hi, uint32_t& lo) {
  float y0, y1;
  unpk2(Y, y0, y1);
  hi = pack_bf16(y0, y1);
  const f32x2 R = fma2(pk2u(hi << 16, hi & 0xffff0000u), pk2(-1.f, -1.f), Y);
  float q0, q1;
  unpk2(R, q0, q1);
  lo = pack_bf16(q0, q1);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) k_edge_encode_tc(EncTcArgs a) {
  extern __shared__ __align__(1024) uint8_t raw[];
  SmemEnc& sm = *reinterpret_cast<SmemEnc*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int E = *a.n_edges;
  const int ntiles = (E + TILE - 1) / TILE;
  const int npairs = (ntiles + 1) / 2;
  const bool exact = a.exact != 0;

  if (warp == MMA_WARP) tmem_alloc(&sm.tmem_base, 512);
  if (tid == 0) {
    mbar_init(&sm.w_full, 1);
    for (int g = 0; g < 2; g++) {
      mbar_init(&sm.a_ready[g], 256);
      mbar_init(&sm.d_ready[g], 1);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 3 * 128; i += THREADS) (&sm.bias[0][0])[i] = a.bias[i];
  for (int i = tid; i < 128; i += THREADS) {
    sm.ln_w[i] = a.ln_w[i];
    sm.ln_b[i] = a.ln_b[i];
  }
  if (tid < GAMD_NRBF) sm.centers[tid] = a.expand_edge ? a.centers[tid] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = sm.tmem_base;

  if (warp < EPI_WARPS) {
    // ===== epilogue warps: thread = edge row; warp = (tile slot g, column half ch, lane quadrant wq) =====
    const int g = warp >> 3, ch = (warp >> 2) & 1, wq = warp & 3;
    const int r = wq * 32 + lane;
    const int col0 = ch * 64;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t Dc = tb + lane_base + g * 256 + col0;
    const uint32_t AH = tb + lane_base + g * 256 + 128, AL = AH + 64;      // whole-row A base (packed columns)
    const uint32_t bias_addr = smem_u32(&sm.bias[0][col0]);
    const int bar_id = 1 + g * 4 + wq;
    uint32_t d_par = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int tile = pair * 2 + g;
      if (tile >= ntiles) continue;
      const int e = tile * TILE + r;
      const bool valid = e < E;

      // ---- stage 0 operand: edge features, my 32 of the 64 (zero padded) columns -> TMEM A ----
      {
        float ux = 0.f, uy = 0.f, uz = 0.f, dh = 0.f, flag = 0.f;
        if (valid) {
          const int c = a.edst[e], n = a.col[e];
          const float4 pc = a.pos[c], pn = a.pos[n];
          // rel = pos[neigh] - pos[center]; remainder(rel + L/2, L) - L/2   (nn_module.py:615-621)
          float rr[3] = {pn.x - pc.x, pn.y - pc.y, pn.z - pc.z};
          if (a.dynbox) { rr[0] = pc.x - pn.x; rr[1] = pc.y - pn.y; rr[2] = pc.z - pn.z; }
#pragma unroll
          for (int d = 0; d < 3; d++) {
            const float half = 0.5f * a.box[d];
            const float t = __fadd_rn(rr[d], half);
            float m = fmodf(t, a.box[d]);
            if (m < 0.f) m = __fadd_rn(m, a.box[d]);
            rr[d] = __fsub_rn(m, half);
            if (a.dynbox) rr[d] = -rr[d];
          }
          const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rr[0], rr[0]), __fmul_rn(rr[1], rr[1])), __fmul_rn(rr[2], rr[2])));
          const float den = dist + 1e-8f;
          ux = rr[0] / den; uy = rr[1] / den; uz = rr[2] / den;
          dh = (dist - a.length_mean) / a.length_std;
          if (a.use_bond) {
            const int ic = a.orig_id ? a.orig_id[c] : c, in = a.orig_id ? a.orig_id[n] : n;
            if (ic / a.atoms_per_frame == in / a.atoms_per_frame) {
              const int lc = ic % a.atoms_per_frame, ln = in % a.atoms_per_frame;
#pragma unroll
              for (int k = 0; k < GAMD_MAX_BOND; k++) flag = (a.bond[lc * GAMD_MAX_BOND + k] == ln) ? 1.f : flag;
            }
          }
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int k = ch * 32 + j;                 // feature column
          float v = 0.f;
          if (k == 0) v = ux;
          else if (k == 1) v = uy;
          else if (k == 2) v = uz;
          else if (k == 3) v = dh;
          else if (k < 4 + GAMD_NRBF) {
            if (a.expand_edge) {
              const float q = dh - sm.centers[k - 4];
              // torch.exp(-40 * radial**2) (nn_module.py:261-263)
              v = valid ? ex2_approx(-57.70780163555854f * (q * q)) : 0.f;
            } else if (k == 4 && a.use_bond) {
              v = flag;
            }
          } else if (k == 4 + GAMD_NRBF && a.use_bond && a.expand_edge) {
            v = flag;
          }
          f[j] = v;
        }
        {
          uint32_t h[16], l[16];
#pragma unroll
          for (int j = 0; j < 16; j++) split_bf16(f[2 * j], f[2 * j + 1], h[j], l[j]);
          tmem_st16(AH + ch * 16, h);
          if (exact) tmem_st16(AL + ch * 16, l);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&sm.a_ready[g]);
      }

      // ---- stages 0 and 1: + bias, GELU, split -> next A (my 64 columns) ----
#pragma unroll 1
      for (int s = 0; s < 2; s++) {
        mbar_wait(&sm.d_ready[g], d_par);
        d_par ^= 1;
        tc_fence_after();
        uint32_t vbuf[2][16];
        tmem_ld16(Dc, vbuf[0]);
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          tmem_wait_ld();
          if (cc < 3) tmem_ld16(Dc + (cc + 1) * 16, vbuf[(cc + 1) & 1]);
          const uint32_t(&v)[16] = vbuf[cc & 1];
          uint32_t h[8], l[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 b = lds128(bias_addr + (s * 128 + cc * 16 + j4 * 4) * 4);
            split2(gelu_as2(add2(pk2u(v[4 * j4], v[4 * j4 + 1]), pk2(b.x, b.y))), h[2 * j4], l[2 * j4]);
            split2(gelu_as2(add2(pk2u(v[4 * j4 + 2], v[4 * j4 + 3]), pk2(b.z, b.w))), h[2 * j4 + 1], l[2 * j4 + 1]);
          }
          tmem_st8(AH + col0 / 2 + cc * 8, h);
          if (exact) tmem_st8(AL + col0 / 2 + cc * 8, l);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&sm.a_ready[g]);
      }

      // ---- stage 2: + bias, LayerNorm over the 128 columns (two warps per row exchange partial sums) ----
      {
        mbar_wait(&sm.d_ready[g], d_par);
        d_par ^= 1;
        tc_fence_after();
        float s1 = 0.f;
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          uint32_t v[16];
          tmem_ld16(Dc + cc * 16, v);
          tmem_wait_ld();
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 b = lds128(bias_addr + (2 * 128 + cc * 16 + j4 * 4) * 4);
            const f32x2 S = add2(add2(pk2u(v[4 * j4], v[4 * j4 + 1]), pk2(b.x, b.y)),
                                 add2(pk2u(v[4 * j4 + 2], v[4 * j4 + 3]), pk2(b.z, b.w)));
            float sa, sb;
            unpk2(S, sa, sb);
            s1 += sa + sb;
          }
        }
        sm.xch[g][r][ch] = s1;
        named_bar_sync(bar_id, 64);
        const float mean = (sm.xch[g][r][0] + sm.xch[g][r][1]) * (1.f / 128.f);
        named_bar_sync(bar_id, 64);
        float s2 = 0.f;
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          uint32_t v[16];
          tmem_ld16(Dc + cc * 16, v);
          tmem_wait_ld();
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 b = lds128(bias_addr + (2 * 128 + cc * 16 + j4 * 4) * 4);
            const f32x2 NM = pk2(-mean, -mean);
            const f32x2 D0 = add2(add2(pk2u(v[4 * j4], v[4 * j4 + 1]), pk2(b.x, b.y)), NM);
            const f32x2 D1 = add2(add2(pk2u(v[4 * j4 + 2], v[4 * j4 + 3]), pk2(b.z, b.w)), NM);
            const f32x2 Q = fma2(D1, D1, mul2(D0, D0));
            float qa, qb;
            unpk2(Q, qa, qb);
            s2 += qa + qb;
          }
        }
        sm.xch[g][r][ch] = s2;
        named_bar_sync(bar_id, 64);
        const float rstd = 1.f / sqrtf((sm.xch[g][r][0] + sm.xch[g][r][1]) * (1.f / 128.f) + 1e-5f);
        named_bar_sync(bar_id, 64);
        uint8_t* blob = a.e_blob + (size_t)tile * 65536;
        const uint32_t lnw_addr = smem_u32(&sm.ln_w[col0]), lnb_addr = smem_u32(&sm.ln_b[col0]);
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          uint32_t v[16];
          tmem_ld16(Dc + cc * 16, v);
          tmem_wait_ld();
          uint32_t hh[8], ll[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 b = lds128(bias_addr + (2 * 128 + cc * 16 + j4 * 4) * 4);
            const float4 w = lds128(lnw_addr + (cc * 16 + j4 * 4) * 4);
            const float4 o = lds128(lnb_addr + (cc * 16 + j4 * 4) * 4);
            const f32x2 NM = pk2(-mean, -mean), RS = pk2(rstd, rstd);
            // ((x + b - mean) * rstd) * w + o, rounded exactly like the scalar expression
            const f32x2 Y0 = fma2(mul2(add2(add2(pk2u(v[4 * j4], v[4 * j4 + 1]), pk2(b.x, b.y)), NM), RS), pk2(w.x, w.y),
                                  pk2(o.x, o.y));
            const f32x2 Y1 = fma2(mul2(add2(add2(pk2u(v[4 * j4 + 2], v[4 * j4 + 3]), pk2(b.z, b.w)), NM), RS),
                                  pk2(w.z, w.w), pk2(o.z, o.w));
            split2(Y0, hh[2 * j4], ll[2 * j4]);
            split2(Y1, hh[2 * j4 + 1], ll[2 * j4 + 1]);
          }
          if (valid) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int j = 0; j < 4; j++) {
                h[j] = hh[half * 4 + j];
                l[j] = ll[half * 4 + j];
              }
              const int kc = (col0 + cc * 16 + half * 8) >> 3;    // 8-wide k-chunk index
              *reinterpret_cast<uint4*>(blob + ((size_t)kc * 128 + r) * 16) = make_uint4(h[0], h[1], h[2], h[3]);
              if (exact)
                *reinterpret_cast<uint4*>(blob + 32768 + ((size_t)kc * 128 + r) * 16) = make_uint4(l[0], l[1], l[2], l[3]);
            }
          }
        }
        // D of this slot is free again once every thread has finished reading it: the next tile's stage-0
        // arrival on a_ready (after these loads, in program order) is what releases it to the MMA warp
      }
    }
  } else if (warp == MMA_WARP) {
    // the whole warp runs the issue loop on warp-uniform values; only the MMAs / commits are predicated on one
    // elected lane (tc_common.cuh: issuing from inside `if (lane == 0)` halves the MMA issue rate)
    {
      const uint32_t leader = elect_leader();
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      uint32_t a_par[2] = {0, 0};
      bool first = true;
      const uint32_t wbase = smem_u32(sm.w);
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        if (first) {
          mbar_wait(&sm.w_full, 0);
          first = false;
        }
        for (int s = 0; s < 3; s++) {
          const uint32_t off = s == 0 ? OFF_ENC0 : (s == 1 ? OFF_ENC2 : OFF_ENC4);
          const uint32_t part = s == 0 ? W0 : W1;
          const int nks = s == 0 ? 4 : 8;
          for (int g = 0; g < 2; g++) {
            if (pair * 2 + g >= ntiles) continue;
            mbar_wait(&sm.a_ready[g], a_par[g]);
            a_par[g] ^= 1;
            __syncwarp();
            tc_fence_after();
            const uint32_t d = tb + g * 256, ah = d + 128, al = d + 192;
            const int passes = exact ? 3 : 1;
            uint32_t accum = 0;
            for (int p = 0; p < passes; p++) {
              const uint32_t bb = wbase + off + (p == 2 ? part : 0);
              const uint32_t aa = (p == 1) ? al : ah;
              for (int ks = 0; ks < nks; ks++) {
                umma_ts_elect(d, aa + ks * 8, umma_desc_sw128(bb + (ks >> 2) * 16384 + (ks & 3) * 32), idesc, accum, leader);
                accum = 1;
              }
            }
            if (leader) umma_commit(&sm.d_ready[g]);
          }
        }
      }
    }
    __syncwarp();
  } else {
    if (lane == 0 && blockIdx.x < npairs) {
      mbar_arrive_expect_tx(&sm.w_full, W_TOTAL);
      for (int i = 0; i < W_TOTAL / 8192; i++) bulk_g2s(sm.w + i * 8192, a.w_img + i * 8192, 8192, &sm.w_full);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tb, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// Three tiles in flight (default): same arithmetic, the TMEM scheme of mp_tc3.cu / mp_tc2cta.cu - four blocks of 128
// columns = three tile homes + one floating block, activations written IN PLACE over the accumulator (K step j of the
// next operand at columns 16 j: 8 columns of bf16 hi pairs, 8 of lo pairs), a GEMM reads the home and writes the
// floating block, which becomes the new home (published in shared memory before the commit).  The encoder needs no
// gather staging and its 160 KB of weights are resident, so nothing couples the three tile chains.
// ---------------------------------------------------------------------------------------------------------------
constexpr int NSLOT3 = 3;
constexpr int EPI_WARPS3 = 8 * NSLOT3;
constexpr int MMA_WARP3 = EPI_WARPS3;
constexpr int THREADS3 = (EPI_WARPS3 + 2) * 32;

struct __align__(1024) SmemEnc3 {
  uint8_t w[W_TOTAL];          // enc0 hi, enc0 lo, enc2 hi, enc2 lo, enc4 hi, enc4 lo
  float bias[3][128];
  float ln_w[128], ln_b[128];
  float centers[GAMD_NRBF];
  float xch[NSLOT3][128][2];   // LayerNorm partial sums exchanged between the two column-half warps of a row
  uint64_t w_full, a_ready[NSLOT3], d_ready[NSLOT3];
  volatile uint32_t home[NSLOT3];
  uint32_t tmem_base;
};

// STATIC: the MMA warp serves the slots in a fixed round-robin order, so GEMM number n always writes TMEM block
// (3 + n) & 3 and an epilogue thread derives its accumulator block from its own GEMM count - no hand-over through shared
// memory, one barrier polled instead of three (the change that took 10 % off the message-passing kernel); operands are
// announced with one arrival per warp instead of one per thread.
template <bool STATIC>
__global__ void __launch_bounds__(THREADS3, 1) k_edge_encode_tc3(EncTcArgs a) {
  extern __shared__ __align__(1024) uint8_t raw[];
  SmemEnc3& sm = *reinterpret_cast<SmemEnc3*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0 && (smem_u32(raw) & 1023u)) __trap();
  const int E = *a.n_edges;
  const int ntiles = (E + TILE - 1) / TILE;
  const int ngroups = (ntiles + NSLOT3 - 1) / NSLOT3;
  const bool exact = a.exact != 0;

  if (warp == MMA_WARP3) tmem_alloc(&sm.tmem_base, 512);
  if (tid == 0) {
    mbar_init(&sm.w_full, 1);
    for (int g = 0; g < NSLOT3; g++) {
      mbar_init(&sm.a_ready[g], STATIC ? 8 : 256);
      mbar_init(&sm.d_ready[g], 1);
      sm.home[g] = g;
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 3 * 128; i += THREADS3) (&sm.bias[0][0])[i] = a.bias[i];
  for (int i = tid; i < 128; i += THREADS3) {
    sm.ln_w[i] = a.ln_w[i];
    sm.ln_b[i] = a.ln_b[i];
  }
  if (tid < GAMD_NRBF) sm.centers[tid] = a.expand_edge ? a.centers[tid] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = sm.tmem_base;

  if (warp < EPI_WARPS3) {
    // ===== epilogue warps: thread = edge row; warp = (tile slot g, column half ch, lane quadrant wq) =====
    const int g = warp >> 3, ch = (warp >> 2) & 1, wq = warp & 3;
    const int r = wq * 32 + lane;
    const int col0 = ch * 64;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    uint32_t Hb = tb + lane_base + g * 128;          // my lanes of the slot's current home block
    const uint32_t bias_addr = smem_u32(&sm.bias[0][col0]);
    const int bar_id = 1 + g * 4 + wq;
    uint32_t d_par = 0;
    uint32_t nseq = (uint32_t)g;      // STATIC: number of the next GEMM of my slot in the MMA warp's sequence
    // operand ready: every thread's tcgen05.st has completed; one arrival per thread, or per warp (STATIC)
    auto announce = [&]() {
      tmem_wait_st();
      tc_fence_before();
      if (STATIC) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.a_ready[g]);
      } else {
        mbar_arrive(&sm.a_ready[g]);
      }
    };
    auto next_block = [&](uint32_t nact) -> uint32_t {
      if (!STATIC) return sm.home[g];
      const uint32_t b = (3u + nseq) & 3u;
      nseq += nact;
      return b;
    };
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
      const int tile = grp * NSLOT3 + g;
      if (tile >= ntiles) {
        // absent slot of the tail group: under the fixed service order its last home block is written by the tail
        // group's GEMMs - tell the MMA warp that the previous tile's LayerNorm has read it
        if (STATIC) announce();
        continue;
      }
      const uint32_t nact = (uint32_t)min(NSLOT3, ntiles - grp * NSLOT3);
      const int e = tile * TILE + r;
      const bool valid = e < E;

      // ---- stage 0 operand: edge features, my 32 of the 64 (zero padded) columns = K steps 2 ch, 2 ch + 1 ----
      {
        float ux = 0.f, uy = 0.f, uz = 0.f, dh = 0.f, flag = 0.f;
        if (valid) {
          const int c = a.edst[e], n = a.col[e];
          const float4 pc = a.pos[c], pn = a.pos[n];
          // rel = pos[neigh] - pos[center]; remainder(rel + L/2, L) - L/2   (nn_module.py:615-621)
          float rr[3] = {pn.x - pc.x, pn.y - pc.y, pn.z - pc.z};
          if (a.dynbox) { rr[0] = pc.x - pn.x; rr[1] = pc.y - pn.y; rr[2] = pc.z - pn.z; }
#pragma unroll
          for (int d = 0; d < 3; d++) {
            const float half = 0.5f * a.box[d];
            const float t = __fadd_rn(rr[d], half);
            float m = fmodf(t, a.box[d]);
            if (m < 0.f) m = __fadd_rn(m, a.box[d]);
            rr[d] = __fsub_rn(m, half);
            if (a.dynbox) rr[d] = -rr[d];
          }
          const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rr[0], rr[0]), __fmul_rn(rr[1], rr[1])), __fmul_rn(rr[2], rr[2])));
          const float den = dist + 1e-8f;
          ux = rr[0] / den; uy = rr[1] / den; uz = rr[2] / den;
          dh = (dist - a.length_mean) / a.length_std;
          if (a.use_bond) {
            const int ic = a.orig_id ? a.orig_id[c] : c, in = a.orig_id ? a.orig_id[n] : n;
            if (ic / a.atoms_per_frame == in / a.atoms_per_frame) {
              const int lc = ic % a.atoms_per_frame, ln = in % a.atoms_per_frame;
#pragma unroll
              for (int k = 0; k < GAMD_MAX_BOND; k++) flag = (a.bond[lc * GAMD_MAX_BOND + k] == ln) ? 1.f : flag;
            }
          }
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int k = ch * 32 + j;                 // feature column
          float v = 0.f;
          if (k == 0) v = ux;
          else if (k == 1) v = uy;
          else if (k == 2) v = uz;
          else if (k == 3) v = dh;
          else if (k < 4 + GAMD_NRBF) {
            if (a.expand_edge) {
              const float q = dh - sm.centers[k - 4];
              // torch.exp(-40 * radial**2) (nn_module.py:261-263)
              v = valid ? ex2_approx(-57.70780163555854f * (q * q)) : 0.f;
            } else if (k == 4 && a.use_bond) {
              v = flag;
            }
          } else if (k == 4 + GAMD_NRBF && a.use_bond && a.expand_edge) {
            v = flag;
          }
          f[j] = v;
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {                // K step 2 ch + j: columns 16 (2 ch + j): [hi | lo]
          uint32_t h[8], l[8];
#pragma unroll
          for (int q = 0; q < 8; q++) split_bf16(f[16 * j + 2 * q], f[16 * j + 2 * q + 1], h[q], l[q]);
          tmem_st8(Hb + (2 * ch + j) * 16, h);
          if (exact) tmem_st8(Hb + (2 * ch + j) * 16 + 8, l);
        }
        announce();
      }

      // ---- stages 0 and 1: + bias, GELU, split -> next operand, in place (my 64 columns) ----
#pragma unroll 1
      for (int s = 0; s < 2; s++) {
        mbar_wait(&sm.d_ready[g], d_par);
        d_par ^= 1;
        tc_fence_after();
        Hb = tb + lane_base + next_block(nact) * 128u;
        const uint32_t Dc = Hb + col0;
        uint32_t vbuf[2][16];
        tmem_ld16(Dc, vbuf[0]);
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          tmem_wait_ld();
          if (cc < 3) tmem_ld16(Dc + (cc + 1) * 16, vbuf[(cc + 1) & 1]);
          const uint32_t(&v)[16] = vbuf[cc & 1];
          uint32_t h[8], l[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 b = lds128(bias_addr + (s * 128 + cc * 16 + j4 * 4) * 4);
            split2(gelu_as2(add2(pk2u(v[4 * j4], v[4 * j4 + 1]), pk2(b.x, b.y))), h[2 * j4], l[2 * j4]);
            split2(gelu_as2(add2(pk2u(v[4 * j4 + 2], v[4 * j4 + 3]), pk2(b.z, b.w))), h[2 * j4 + 1], l[2 * j4 + 1]);
          }
          tmem_st8(Dc + cc * 16, h);
          if (exact) tmem_st8(Dc + cc * 16 + 8, l);
        }
        announce();
      }

      // ---- stage 2: + bias, LayerNorm over the 128 columns (two warps per row exchange partial sums) ----
      {
        mbar_wait(&sm.d_ready[g], d_par);
        d_par ^= 1;
        tc_fence_after();
        Hb = tb + lane_base + next_block(nact) * 128u;
        const uint32_t Dc = Hb + col0;
        float s1 = 0.f;
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          uint32_t v[16];
          tmem_ld16(Dc + cc * 16, v);
          tmem_wait_ld();
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 b = lds128(bias_addr + (2 * 128 + cc * 16 + j4 * 4) * 4);
            const f32x2 S = add2(add2(pk2u(v[4 * j4], v[4 * j4 + 1]), pk2(b.x, b.y)),
                                 add2(pk2u(v[4 * j4 + 2], v[4 * j4 + 3]), pk2(b.z, b.w)));
            float sa, sb;
            unpk2(S, sa, sb);
            s1 += sa + sb;
          }
        }
        sm.xch[g][r][ch] = s1;
        named_bar_sync(bar_id, 64);
        const float mean = (sm.xch[g][r][0] + sm.xch[g][r][1]) * (1.f / 128.f);
        named_bar_sync(bar_id, 64);
        float s2 = 0.f;
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          uint32_t v[16];
          tmem_ld16(Dc + cc * 16, v);
          tmem_wait_ld();
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 b = lds128(bias_addr + (2 * 128 + cc * 16 + j4 * 4) * 4);
            const f32x2 NM = pk2(-mean, -mean);
            const f32x2 D0 = add2(add2(pk2u(v[4 * j4], v[4 * j4 + 1]), pk2(b.x, b.y)), NM);
            const f32x2 D1 = add2(add2(pk2u(v[4 * j4 + 2], v[4 * j4 + 3]), pk2(b.z, b.w)), NM);
            const f32x2 Q = fma2(D1, D1, mul2(D0, D0));
            float qa, qb;
            unpk2(Q, qa, qb);
            s2 += qa + qb;
          }
        }
        sm.xch[g][r][ch] = s2;
        named_bar_sync(bar_id, 64);
        const float rstd = 1.f / sqrtf((sm.xch[g][r][0] + sm.xch[g][r][1]) * (1.f / 128.f) + 1e-5f);
        uint8_t* blob = a.e_blob + (size_t)tile * 65536;
        const uint32_t lnw_addr = smem_u32(&sm.ln_w[col0]), lnb_addr = smem_u32(&sm.ln_b[col0]);
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          uint32_t v[16];
          tmem_ld16(Dc + cc * 16, v);
          tmem_wait_ld();
          uint32_t hh[8], ll[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 b = lds128(bias_addr + (2 * 128 + cc * 16 + j4 * 4) * 4);
            const float4 w = lds128(lnw_addr + (cc * 16 + j4 * 4) * 4);
            const float4 o = lds128(lnb_addr + (cc * 16 + j4 * 4) * 4);
            const f32x2 NM = pk2(-mean, -mean), RS = pk2(rstd, rstd);
            const f32x2 Y0 = fma2(mul2(add2(add2(pk2u(v[4 * j4], v[4 * j4 + 1]), pk2(b.x, b.y)), NM), RS), pk2(w.x, w.y),
                                  pk2(o.x, o.y));
            const f32x2 Y1 = fma2(mul2(add2(add2(pk2u(v[4 * j4 + 2], v[4 * j4 + 3]), pk2(b.z, b.w)), NM), RS),
                                  pk2(w.z, w.w), pk2(o.z, o.w));
            split2(Y0, hh[2 * j4], ll[2 * j4]);
            split2(Y1, hh[2 * j4 + 1], ll[2 * j4 + 1]);
          }
          if (valid) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
              const int kc = (col0 + cc * 16 + half * 8) >> 3;    // 8-wide k-chunk index
              *reinterpret_cast<uint4*>(blob + ((size_t)kc * 128 + r) * 16) =
                  make_uint4(hh[half * 4], hh[half * 4 + 1], hh[half * 4 + 2], hh[half * 4 + 3]);
              if (exact)
                *reinterpret_cast<uint4*>(blob + 32768 + ((size_t)kc * 128 + r) * 16) =
                    make_uint4(ll[half * 4], ll[half * 4 + 1], ll[half * 4 + 2], ll[half * 4 + 3]);
            }
          }
        }
        // the next tile's stage-0 operand overwrites columns [0, 64) of this block, part of which the OTHER column-half
        // warp of my rows may still be reading: both halves of a row group leave the LayerNorm together
        named_bar_sync(bar_id, 64);
      }
    }
  } else if (STATIC && warp == MMA_WARP3) {
    // MMA issue in a fixed order: (group, stage, slot) - GEMM n reads the slot's home block and writes block (3 + n) & 3
    const uint32_t leader = elect_leader();
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const int n_my_groups = blockIdx.x < ngroups ? (ngroups - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const uint32_t wbase = smem_u32(sm.w);
    uint32_t a_par_bits = 0, home[NSLOT3], n = 0;
#pragma unroll
    for (int g = 0; g < NSLOT3; g++) home[g] = g;
    if (n_my_groups > 0) mbar_wait(&sm.w_full, 0);
    for (int i = 0; i < n_my_groups; i++) {
      const int grp = blockIdx.x + i * gridDim.x;
      const int nact = min(NSLOT3, ntiles - grp * NSLOT3);
      // tail group: the blocks of the absent slots are written too - not before those slots have read their last rows
#pragma unroll
      for (int q = 1; q < NSLOT3; q++) {
        if (q < nact) continue;
        uint32_t spins = 0;
        while (!__all_sync(0xffffffffu, mbar_test_wait(&sm.a_ready[q], (a_par_bits >> q) & 1u)))
          if (++spins > (1u << 26)) __trap();
        a_par_bits ^= 1u << q;
        tc_fence_after();
      }
#pragma unroll 1
      for (int s = 0; s < 3; s++) {
        const uint32_t off = s == 0 ? OFF_ENC0 : (s == 1 ? OFF_ENC2 : OFF_ENC4);
        const uint32_t part = s == 0 ? W0 : W1;
        const int nks = s == 0 ? 4 : 8;
#pragma unroll
        for (int g = 0; g < NSLOT3; g++) {
          if (g >= nact) continue;
          uint32_t spins = 0;
          while (!__all_sync(0xffffffffu, mbar_test_wait(&sm.a_ready[g], (a_par_bits >> g) & 1u)))
            if (++spins > (1u << 26)) __trap();
          a_par_bits ^= 1u << g;
          tc_fence_after();
          const uint32_t blk = (3u + n) & 3u;
          const uint32_t d = tb + blk * 128u, ab = tb + home[g] * 128u;
          const int passes = exact ? 3 : 1;
          uint32_t accum = 0;
          for (int p = 0; p < passes; p++) {
            const uint32_t bb = wbase + off + (p == 2 ? part : 0);
            const uint32_t aa = ab + (p == 1 ? 8u : 0u);
            for (int ks = 0; ks < nks; ks++) {
              umma_ts_elect(d, aa + ks * 16, umma_desc_sw128(bb + (ks >> 2) * 16384 + (ks & 3) * 32), idesc, accum, leader);
              accum = 1;
            }
          }
          if (leader) umma_commit(&sm.d_ready[g]);
          __syncwarp();
          home[g] = blk;
          n++;
        }
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP3) {
    // MMA issue: event loop over the three slots (see mp_tc3.cu); weights resident, so only the operands gate a GEMM
    const uint32_t leader = elect_leader();
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const int n_my_groups = blockIdx.x < ngroups ? (ngroups - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int totalQ = 3 * n_my_groups;
    int Qg[NSLOT3];
    uint32_t a_par[NSLOT3], home[NSLOT3];
#pragma unroll
    for (int g = 0; g < NSLOT3; g++) {
      Qg[g] = 0;
      a_par[g] = 0;
      home[g] = g;
    }
    uint32_t floating = NSLOT3;
    int first = 0;
    uint32_t spins = 0;
    const uint32_t wbase = smem_u32(sm.w);
    bool w_ready = false;
    for (;;) {
      bool done = true;
#pragma unroll
      for (int g = 0; g < NSLOT3; g++) done = done && Qg[g] >= totalQ;
      if (done) break;
      if (!w_ready) {
        mbar_wait(&sm.w_full, 0);
        w_ready = true;
      }
      bool progressed = false;
      int pick = -1, best = NSLOT3;
#pragma unroll
      for (int g = 0; g < NSLOT3; g++) {
        const int Q = Qg[g];
        if (Q >= totalQ) continue;
        const int grp = blockIdx.x + (Q / 3) * gridDim.x;
        if (grp * NSLOT3 + g >= ntiles) {          // absent tile of the tail group
          Qg[g] = totalQ;
          progressed = true;
          continue;
        }
        if (!__all_sync(0xffffffffu, mbar_test_wait(&sm.a_ready[g], a_par[g]))) continue;
        int pr = g - first;
        if (pr < 0) pr += NSLOT3;
        if (pr < best) {
          best = pr;
          pick = g;
        }
      }
#pragma unroll
      for (int g = 0; g < NSLOT3; g++) {
        if (g != pick) continue;
        const int s = Qg[g] % 3;
        a_par[g] ^= 1;
        tc_fence_after();
        const uint32_t off = s == 0 ? OFF_ENC0 : (s == 1 ? OFF_ENC2 : OFF_ENC4);
        const uint32_t part = s == 0 ? W0 : W1;
        const int nks = s == 0 ? 4 : 8;
        const uint32_t d = tb + floating * 128u, ab = tb + home[g] * 128u;
        if (leader) sm.home[g] = floating;
        __threadfence_block();
        const int passes = exact ? 3 : 1;
        uint32_t accum = 0;
        for (int p = 0; p < passes; p++) {
          const uint32_t bb = wbase + off + (p == 2 ? part : 0);
          const uint32_t aa = ab + (p == 1 ? 8u : 0u);
          for (int ks = 0; ks < nks; ks++) {
            umma_ts_elect(d, aa + ks * 16, umma_desc_sw128(bb + (ks >> 2) * 16384 + (ks & 3) * 32), idesc, accum, leader);
            accum = 1;
          }
        }
        if (leader) umma_commit(&sm.d_ready[g]);
        __syncwarp();
        const uint32_t old_home = home[g];
        home[g] = floating;
        floating = old_home;
        first = g + 1 == NSLOT3 ? 0 : g + 1;
        Qg[g]++;
        progressed = true;
      }
      if (progressed) spins = 0;
      else if (++spins > (1u << 26)) __trap();
    }
    __syncwarp();
  } else {
    if (lane == 0 && blockIdx.x < ngroups) {
      mbar_arrive_expect_tx(&sm.w_full, W_TOTAL);
      for (int i = 0; i < W_TOTAL / 8192; i++) bulk_g2s(sm.w + i * 8192, a.w_img + i * 8192, 8192, &sm.w_full);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP3) tmem_dealloc(tb, 512);
}

}  // namespace

int edge_encode_tc_launch(gamd_ctx* ctx, const float4* pos_feat, const int* orig_id, int atoms_per_frame,
                          const float box[3], cudaStream_t st) {
  const size_t smem = sizeof(SmemEnc) + 1024;
  if (!(ctx->attr_mask & GAMD_ATTR_ENC_TC)) {
    GAMD_CUDA(cudaFuncSetAttribute(k_edge_encode_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->attr_mask |= GAMD_ATTR_ENC_TC;
  }
  const ModelW& mw = ctx->mw;
  EncTcArgs a;
  a.w_img = ctx->d_wimg_enc;
  a.bias = ctx->d_tc_bias_enc;
  a.ln_w = mw.eln_w;
  a.ln_b = mw.eln_b;
  a.centers = mw.centers;
  a.pos = pos_feat;
  a.col = ctx->col_idx;
  a.edst = ctx->edge_dst;
  a.n_edges = ctx->n_edges;
  a.orig_id = orig_id;
  a.bond = ctx->d_bond;
  a.e_blob = reinterpret_cast<uint8_t*>(ctx->e_emb);
  a.length_mean = mw.length_mean;
  a.length_std = mw.length_std;
  a.box[0] = box[0]; a.box[1] = box[1]; a.box[2] = box[2];
  a.n_edge_in = mw.n_edge_in;
  a.use_bond = mw.use_bond;
  a.expand_edge = mw.expand_edge;
  a.atoms_per_frame = atoms_per_frame;
  a.exact = ctx->desc.precision == GAMD_PREC_BF16X3 ? 1 : 0;
  a.dynbox = mw.kind == GAMD_MODEL_DYNBOX ? 1 : 0;
  if (ctx->enc_variant == 3 || ctx->enc_variant == 4) {
    if (!(ctx->attr_mask & GAMD_ATTR_ENC_TC3)) {
      GAMD_CUDA(cudaFuncSetAttribute(k_edge_encode_tc3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemEnc3)));
      GAMD_CUDA(cudaFuncSetAttribute(k_edge_encode_tc3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemEnc3)));
      ctx->attr_mask |= GAMD_ATTR_ENC_TC3;
    }
    if (ctx->enc_variant == 4) k_edge_encode_tc3<true><<<ctx->sm_count, THREADS3, sizeof(SmemEnc3), st>>>(a);
    else k_edge_encode_tc3<false><<<ctx->sm_count, THREADS3, sizeof(SmemEnc3), st>>>(a);
  } else {
    k_edge_encode_tc<<<ctx->sm_count, THREADS, smem, st>>>(a);
  }
  GAMD_LAUNCH_CHECK();
  return 0;
}
