// Node update on the tensor cores (K7 / K10 / K11): for every 128-node tile
//
//   [mid]   t = silu(phi_edge(agg) + phi_dst(hn_prev));  h += phi.1(t);   hn = LN_next(h);
//           srcA = src_affine(hn); dstA = dst_affine(hn); pd = phi_dst(hn)            (nn_module.py:147, :202, :136-137)
//   [first] h = node_emb | node_encoder(type);  hn = LN_0(h);  srcA, dstA, pd          (nn_module.py:681 / :554)
//   [last]  h += phi.1(...);  force = decoder.2(gelu(decoder.0(h)))                    (nn_module.py:684)
//
// Same machinery as mp_tc.cu: the A operand of every GEMM lives in TMEM (thread = node row), weights stream through
// a ring of SWIZZLE_128B images, accumulators in TMEM, bf16x3 split for fp32-grade results.  Row-major global
// rows (agg, pd, h in; h, hn, srcA, dstA, pd out) are moved with coalesced 64-byte row segments through a
// per-warp staging tile in shared memory, because the TMEM register layout is one row per thread.
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int TILE = 128;
constexpr int WCHUNK = 32768;
constexpr int RING = 4;
constexpr int GROW = 80;             // staging row stride (64 B data + 16 B pad)
constexpr int EPI_WARPS = 16;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int THREADS = (EPI_WARPS + 2) * 32;

enum { W_PEDGE = 0, W_PHI = 1, W_SRC = 2, W_DST = 3, W_PDST = 4, W_DEC0 = 5 };   // image order inside a layer slab
constexpr int MODE_FIRST = 0, MODE_MID = 1, MODE_LAST = 2;

struct __align__(1024) SmemNode {
  uint8_t w[RING][WCHUNK];
  uint8_t stage[EPI_WARPS][2][32 * GROW];   // [0] inbound rows, [1] outbound rows
  float bias[6][128];
  float ln_w[128], ln_b[128];
  float h0[128];                            // first mode, LJ: the shared node embedding
  float dec2[3][128];
  float xch[2][128][4];
  uint64_t full[RING], empty[RING], a_ready[2], d_ready[2];
  uint32_t tmem_base;
};

struct NodeTcArgs {
  const uint8_t* w_img;      // this layer's slab: [6 matrices][hi|lo][WCHUNK]   (cur layer: pedge, phi, -, -, -, dec0)
  const uint8_t* w_img_next; // next layer's slab (src, dst, pdst)
  const float *b_pedge, *b_phi, *b_src, *b_dst, *b_pdst, *b_dec0, *ln_w, *ln_b;
  const float *node_emb, *nenc_w, *nenc_b, *dec2_w, *dec2_b;
  const float4* pos_feat;    // .w = node type (water)
  const float* agg;
  float *h, *hn, *srcA, *dstA, *pd, *pred;
  int n_atoms, mode, exact;
  // domain decomposition, fused halo push: rows of owned atoms that a neighbouring rank needs are ALSO stored straight
  // into that rank's receive buffer over NVLink peer memory ([LN(h) | src_affine] = 256 floats per slot)
  const int *perm, *push_slot[2];   // sorted row -> local atom; local atom -> slot in the left / right buffer or -1
  float* push_rows[2];              // the neighbours' receive buffers (peer-mapped), nullptr = off
  int push_n;                       // entries of the slot maps (owned atoms)
};

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async16s(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ float silu_exact(float x) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + t));
  return x * r;
}
__device__ __forceinline__ float gelu_as(float x) {   // exact-erf GELU, A&S 7.1.26 (see enc_tc.cu)
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), fmaf(-p, e, 1.f), hx);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct RowIO {
  uint32_t in_buf, out_buf;   // shared addresses of my warp's staging tiles
  uint32_t my_row;            // lane * GROW
  uint32_t l_off;             // (lane>>2)*GROW + (lane&3)*16
  int l_row, l_col;           // lane>>2, (lane&3)*4
  int row0;                   // first node row of my warp's 32 rows
  int n_atoms, col0;
  int slot[2];                // my row's slot in the left / right neighbour's receive buffer (-1: not sent)
  float* push_rows[2];
};

// 16 columns [col0 + cc*16, +16) of my warp's 32 rows of a row-major [n,128] fp32 matrix -> x[16] of my row,
// in two halves, so that the copy of the next chunk runs underneath the arithmetic on the current one (and the
// first chunk of a phase underneath the GEMM that precedes it): issue_rows starts the copy into my warp's inbound tile,
// take_rows waits for it and reads my row.  One copy in flight per warp: issue the next one only after take_rows.
__device__ __forceinline__ void issue_rows(const RowIO& io, const float* __restrict__ M, int cc) {
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; it++) {
    int row = io.row0 + it * 8 + io.l_row;
    row = row < io.n_atoms ? row : io.n_atoms - 1;
    cp_async16s(io.in_buf + io.l_off + it * 8 * GROW, M + (size_t)row * 128 + io.col0 + cc * 16 + io.l_col);
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void take_rows(const RowIO& io, float (&x)[16]) {
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncwarp();
#pragma unroll
  for (int j4 = 0; j4 < 4; j4++) {
    const float4 v = lds128(io.in_buf + io.my_row + j4 * 16);
    x[4 * j4] = v.x; x[4 * j4 + 1] = v.y; x[4 * j4 + 2] = v.z; x[4 * j4 + 3] = v.w;
  }
}

// x[16] of my row -> 16 columns of my warp's 32 rows of a row-major matrix (coalesced 64-byte row segments)
__device__ __forceinline__ void store_rows(const RowIO& io, float* __restrict__ M, int cc, const float (&x)[16]) {
  __syncwarp();
#pragma unroll
  for (int j4 = 0; j4 < 4; j4++)
    sts128(io.out_buf + io.my_row + j4 * 16, make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]));
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int row = io.row0 + it * 8 + io.l_row;
    const float4 v = lds128(io.out_buf + io.l_off + it * 8 * GROW);
    if (row < io.n_atoms) *reinterpret_cast<float4*>(M + (size_t)row * 128 + io.col0 + cc * 16 + io.l_col) = v;
  }
}

// the same store, plus the peer-memory copy of the rows a neighbouring rank needs: 64-byte segments at
// remote[slot][part * 128 + column] (part 0 = LN(h), 1 = src_affine)
__device__ __forceinline__ void store_rows_push(const RowIO& io, float* __restrict__ M, int cc, const float (&x)[16], int part) {
  store_rows(io, M, cc, x);
#pragma unroll
  for (int side = 0; side < 2; side++) {
    if (!io.push_rows[side]) continue;     // warp-uniform
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int slot = __shfl_sync(0xffffffffu, io.slot[side], it * 8 + io.l_row);
      if (slot >= 0) {
        const float4 v = lds128(io.out_buf + io.l_off + it * 8 * GROW);
        *reinterpret_cast<float4*>(io.push_rows[side] + (size_t)slot * 256 + part * 128 + io.col0 + cc * 16 + io.l_col) = v;
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1) k_node_tc(NodeTcArgs a) {
  extern __shared__ __align__(1024) uint8_t raw[];
  SmemNode& sm = *reinterpret_cast<SmemNode*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (a.n_atoms + TILE - 1) / TILE;
  const int npairs = (ntiles + 1) / 2;
  const bool exact = a.exact != 0;
  const int mode = a.mode;
  // GEMM stages of this mode, as indices into the weight image table
  //   first: src, dst, pdst (next layer)        mid: pedge, phi (cur) + src, dst, pdst (next)       last: pedge, phi, dec0 (cur)
  const int n_stage = mode == MODE_MID ? 5 : 3;
  auto stage_img = [&](int s) -> const uint8_t* {
    if (mode == MODE_FIRST) return a.w_img_next + (size_t)(W_SRC + s) * 2 * WCHUNK;
    if (mode == MODE_LAST) return a.w_img + (size_t)(s == 2 ? W_DEC0 : s) * 2 * WCHUNK;
    return s < 2 ? a.w_img + (size_t)s * 2 * WCHUNK : a.w_img_next + (size_t)s * 2 * WCHUNK;
  };

  if (warp == MMA_WARP) tmem_alloc(&sm.tmem_base, 512);
  if (tid == 0) {
    for (int i = 0; i < RING; i++) {
      mbar_init(&sm.full[i], 1);
      mbar_init(&sm.empty[i], 1);
    }
    for (int g = 0; g < 2; g++) {
      mbar_init(&sm.a_ready[g], 8);        // one arrival per epilogue warp of the tile
      mbar_init(&sm.d_ready[g], 1);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 128; i += THREADS) {
    sm.bias[0][i] = a.b_pedge ? a.b_pedge[i] : 0.f;
    sm.bias[1][i] = a.b_phi ? a.b_phi[i] : 0.f;
    sm.bias[2][i] = a.b_src ? a.b_src[i] : 0.f;
    sm.bias[3][i] = a.b_dst ? a.b_dst[i] : 0.f;
    sm.bias[4][i] = a.b_pdst ? a.b_pdst[i] : 0.f;
    sm.bias[5][i] = a.b_dec0 ? a.b_dec0[i] : 0.f;
    sm.ln_w[i] = a.ln_w ? a.ln_w[i] : 1.f;
    sm.ln_b[i] = a.ln_b ? a.ln_b[i] : 0.f;
    sm.h0[i] = a.node_emb ? a.node_emb[i] : 0.f;
    for (int k = 0; k < 3; k++) sm.dec2[k][i] = a.dec2_w ? a.dec2_w[k * 128 + i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = sm.tmem_base;

  if (warp < EPI_WARPS) {
    const int g = warp >> 3, ch = (warp >> 2) & 1, wq = warp & 3;
    const int r = wq * 32 + lane;
    const int col0 = ch * 64;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t Dc = tb + lane_base + g * 256 + col0;
    const uint32_t AH = tb + lane_base + g * 256 + 128 + col0 / 2, AL = AH + 64;
    const int bar_id = 1 + g * 4 + wq;
    RowIO io;
    io.in_buf = smem_u32(sm.stage[warp][0]);
    io.out_buf = smem_u32(sm.stage[warp][1]);
    io.my_row = lane * GROW;
    io.l_row = lane >> 2;
    io.l_col = (lane & 3) * 4;
    io.l_off = (lane >> 2) * GROW + (lane & 3) * 16;
    io.n_atoms = a.n_atoms;
    io.col0 = col0;
    io.slot[0] = io.slot[1] = -1;
    io.push_rows[0] = a.push_rows[0];
    io.push_rows[1] = a.push_rows[1];
    uint32_t d_par = 0;
    uint64_t* const a_bar = &sm.a_ready[g];
    uint64_t* const d_bar = &sm.d_ready[g];
    auto bias_at = [&](int which, int cc, int j4) { return lds128(smem_u32(&sm.bias[which][col0 + cc * 16 + j4 * 4])); };
    auto write_A = [&](int cc, const float (&x)[16]) {
      uint32_t h[8], l[8];
#pragma unroll
      for (int j = 0; j < 8; j++) split_bf16(x[2 * j], x[2 * j + 1], h[j], l[j]);
      tmem_st8(AH + cc * 8, h);
      if (exact) tmem_st8(AL + cc * 8, l);
    };
    auto signal_A = [&]() {
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_bar);
    };
    auto wait_D = [&]() {
      mbar_wait(d_bar, d_par);
      d_par ^= 1;
      tc_fence_after();
    };
    auto load_D = [&](int cc, float (&x)[16]) {
      uint32_t v[16];
      tmem_ld16(Dc + cc * 16, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; j++) x[j] = __uint_as_float(v[j]);
    };
    // LayerNorm of the row whose fp32 values sit in my D columns (both column halves exchange partial sums),
    // then: hn -> global, hn (bf16 hi/lo) -> A operand
    auto layer_norm_to_A = [&]() {
      float s1 = 0.f;
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        float x[16];
        load_D(cc, x);
#pragma unroll
        for (int j = 0; j < 16; j++) s1 += x[j];
      }
      sm.xch[g][r][ch] = s1;
      named_bar_sync(bar_id, 64);
      const float mean = (sm.xch[g][r][0] + sm.xch[g][r][1]) * (1.f / 128.f);
      named_bar_sync(bar_id, 64);
      float s2 = 0.f;
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        float x[16];
        load_D(cc, x);
#pragma unroll
        for (int j = 0; j < 16; j++) s2 += (x[j] - mean) * (x[j] - mean);
      }
      sm.xch[g][r][ch] = s2;
      named_bar_sync(bar_id, 64);
      const float rstd = 1.f / sqrtf((sm.xch[g][r][0] + sm.xch[g][r][1]) * (1.f / 128.f) + 1e-5f);
      named_bar_sync(bar_id, 64);
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        float x[16];
        load_D(cc, x);
#pragma unroll
        for (int j4 = 0; j4 < 4; j4++) {
          const float4 w = lds128(smem_u32(&sm.ln_w[col0 + cc * 16 + j4 * 4]));
          const float4 o = lds128(smem_u32(&sm.ln_b[col0 + cc * 16 + j4 * 4]));
          x[4 * j4] = (x[4 * j4] - mean) * rstd * w.x + o.x;
          x[4 * j4 + 1] = (x[4 * j4 + 1] - mean) * rstd * w.y + o.y;
          x[4 * j4 + 2] = (x[4 * j4 + 2] - mean) * rstd * w.z + o.z;
          x[4 * j4 + 3] = (x[4 * j4 + 3] - mean) * rstd * w.w + o.w;
        }
        store_rows_push(io, a.hn, cc, x, 0);
        write_A(cc, x);
      }
    };

    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int tile = pair * 2 + g;
      if (tile >= ntiles) continue;
      const int n0 = tile * TILE;
      io.row0 = n0 + wq * 32;
      const int node = n0 + r;
      const bool valid = node < a.n_atoms;
      if (a.push_rows[0] || a.push_rows[1]) {
        const int lid = valid ? a.perm[node] : -1;
#pragma unroll
        for (int side = 0; side < 2; side++)
          io.slot[side] = (lid >= 0 && a.push_rows[side] && lid < a.push_n) ? a.push_slot[side][lid] : -1;
      }

      if (mode == MODE_FIRST) {
        // h0 = node_emb (LJ) or node_encoder(type) (water); stored to h and (fp32) into my D columns for the LN
        const float t = (valid && a.nenc_w) ? a.pos_feat[node].w : 0.f;
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          float x[16];
#pragma unroll
          for (int j = 0; j < 16; j++) {
            const int c = col0 + cc * 16 + j;
            x[j] = a.nenc_w ? fmaf(t, a.nenc_w[c], a.nenc_b[c]) : sm.h0[c];
          }
          store_rows(io, a.h, cc, x);
          uint32_t v[16];
#pragma unroll
          for (int j = 0; j < 16; j++) v[j] = __float_as_uint(x[j]);
          tmem_st16(Dc + cc * 16, v);
        }
        tmem_wait_st();
        layer_norm_to_A();
        signal_A();
      } else {
        // A <- agg rows (the copy of chunk cc + 1 runs underneath the split / TMEM store of chunk cc)
        issue_rows(io, a.agg, 0);
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          float x[16];
          take_rows(io, x);
          if (cc < 3) issue_rows(io, a.agg, cc + 1);
          write_A(cc, x);
        }
        signal_A();
        issue_rows(io, a.pd, 0);                   // first phi_dst chunk underneath the phi_edge GEMM
        // phi_edge(agg) + b + phi_dst(hn_prev) -> SiLU -> A
        wait_D();
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          float x[16], p[16];
          load_D(cc, x);
          take_rows(io, p);
          if (cc < 3) issue_rows(io, a.pd, cc + 1);
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 b = bias_at(0, cc, j4);
            x[4 * j4] = silu_exact(x[4 * j4] + b.x + p[4 * j4]);
            x[4 * j4 + 1] = silu_exact(x[4 * j4 + 1] + b.y + p[4 * j4 + 1]);
            x[4 * j4 + 2] = silu_exact(x[4 * j4 + 2] + b.z + p[4 * j4 + 2]);
            x[4 * j4 + 3] = silu_exact(x[4 * j4 + 3] + b.w + p[4 * j4 + 3]);
          }
          write_A(cc, x);
        }
        signal_A();
        issue_rows(io, a.h, 0);                    // first residual chunk underneath the phi GEMM
        // phi.1 + b + h (residual with the un-normalised h) -> h
        wait_D();
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          float x[16], hr[16];
          load_D(cc, x);
          take_rows(io, hr);
          if (cc < 3) issue_rows(io, a.h, cc + 1);
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 b = bias_at(1, cc, j4);
            x[4 * j4] += b.x + hr[4 * j4];
            x[4 * j4 + 1] += b.y + hr[4 * j4 + 1];
            x[4 * j4 + 2] += b.z + hr[4 * j4 + 2];
            x[4 * j4 + 3] += b.w + hr[4 * j4 + 3];
          }
          if (mode == MODE_MID) {
            store_rows(io, a.h, cc, x);
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = __float_as_uint(x[j]);
            tmem_st16(Dc + cc * 16, v);          // keep the fp32 row in TMEM for the LayerNorm passes
          } else {
            write_A(cc, x);                      // last layer: the decoder consumes h directly
          }
        }
        if (mode == MODE_MID) {
          tmem_wait_st();
          layer_norm_to_A();
        }
        signal_A();
      }

      if (mode != MODE_LAST) {
        // three affines of hn: + bias -> srcA, dstA, pd
#pragma unroll 1
        for (int k = 0; k < 3; k++) {
          wait_D();
          float* out = k == 0 ? a.srcA : (k == 1 ? a.dstA : a.pd);
#pragma unroll
          for (int cc = 0; cc < 4; cc++) {
            float x[16];
            load_D(cc, x);
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) {
              const float4 b = bias_at(2 + k, cc, j4);
              x[4 * j4] += b.x; x[4 * j4 + 1] += b.y; x[4 * j4 + 2] += b.z; x[4 * j4 + 3] += b.w;
            }
            if (k == 0) store_rows_push(io, out, cc, x, 1);
            else store_rows(io, out, cc, x);
          }
          // D of this slot is free again (A = hn stays): lets the MMA warp start the next affine / next tile
          if (k < 2) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_bar);
          }
        }
      } else {
        // decoder: gelu(dec0(h) + b) . dec2^T + b2 -> 3 force components
        wait_D();
        float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          float x[16];
          load_D(cc, x);
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 b = bias_at(5, cc, j4);
            const float y0 = gelu_as(x[4 * j4] + b.x), y1 = gelu_as(x[4 * j4 + 1] + b.y);
            const float y2 = gelu_as(x[4 * j4 + 2] + b.z), y3 = gelu_as(x[4 * j4 + 3] + b.w);
#pragma unroll
            for (int k = 0; k < 3; k++) {
              const float4 w = lds128(smem_u32(&sm.dec2[k][col0 + cc * 16 + j4 * 4]));
              o[k] = fmaf(y0, w.x, fmaf(y1, w.y, fmaf(y2, w.z, fmaf(y3, w.w, o[k]))));
            }
          }
        }
        if (ch == 1) {
          sm.xch[g][r][0] = o[0]; sm.xch[g][r][1] = o[1]; sm.xch[g][r][2] = o[2];
        }
        named_bar_sync(bar_id, 64);
        if (ch == 0 && valid) {
          a.pred[(size_t)node * 3 + 0] = o[0] + sm.xch[g][r][0] + a.dec2_b[0];
          a.pred[(size_t)node * 3 + 1] = o[1] + sm.xch[g][r][1] + a.dec2_b[1];
          a.pred[(size_t)node * 3 + 2] = o[2] + sm.xch[g][r][2] + a.dec2_b[2];
        }
        named_bar_sync(bar_id, 64);
      }
    }
  } else if (warp == MMA_WARP) {
    // the whole warp runs the issue loop on warp-uniform values; only the MMAs / commits are predicated on one
    // elected lane (tc_common.cuh: issuing from inside `if (lane == 0)` halves the MMA issue rate)
    {
      const uint32_t leader = elect_leader();
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      uint32_t a_par[2] = {0, 0};
      uint32_t q = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        for (int s = 0; s < n_stage; s++) {
          const uint32_t slot_hi = q % RING, slot_lo = (q + 1) % RING;
          mbar_wait(&sm.full[slot_hi], (q / RING) & 1);
          if (exact) mbar_wait(&sm.full[slot_lo], ((q + 1) / RING) & 1);
          const uint32_t bhi = smem_u32(sm.w[slot_hi]), blo = smem_u32(sm.w[slot_lo]);
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
              const uint32_t bb = (p == 2) ? blo : bhi;
              const uint32_t aa = (p == 1) ? al : ah;
#pragma unroll
              for (int ks = 0; ks < 8; ks++) {
                umma_ts_elect(d, aa + ks * 8, umma_desc_sw128(bb + (ks >> 2) * 16384 + (ks & 3) * 32), idesc, accum, leader);
                accum = 1;
              }
            }
            if (leader) umma_commit(&sm.d_ready[g]);
          }
          if (leader) umma_commit(&sm.empty[slot_hi]);
          if (exact && leader) umma_commit(&sm.empty[slot_lo]);
          q += exact ? 2 : 1;
        }
      }
    }
    __syncwarp();
  } else {
    if (lane == 0) {
      uint32_t q = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x)
        for (int s = 0; s < n_stage; s++)
          for (int part = 0; part < (exact ? 2 : 1); part++) {
            const int slot = q % RING;
            if (q >= RING) mbar_wait(&sm.empty[slot], ((q / RING) - 1) & 1);
            mbar_arrive_expect_tx(&sm.full[slot], WCHUNK);
            const uint8_t* src = stage_img(s) + (size_t)part * WCHUNK;
#pragma unroll
            for (int i = 0; i < 4; i++) bulk_g2s(sm.w[slot] + i * 8192, src + i * 8192, 8192, &sm.full[slot]);
            q++;
          }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tb, 512);
}

// rows whose edge run straddles 32-edge blocks: agg = tail part of the first block + head parts of the others
__global__ void k_agg_fixup(const int* __restrict__ row_ptr, const float* __restrict__ part, float* __restrict__ agg,
                            int n_atoms, const int* __restrict__ n_edges) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n_atoms) return;
  int rs = row_ptr[i], re = row_ptr[i + 1];
  if (*n_edges == 0) re = rs;      // empty / overflowed edge list (see k_nbr_guard): no receiver has edges
  float4* dst = reinterpret_cast<float4*>(agg + (size_t)i * 128) + lane;
  if (re <= rs) {
    *dst = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int t0 = rs >> 5, t1 = (re - 1) >> 5;
  if (t0 == t1) return;
  float4 v = reinterpret_cast<const float4*>(part + ((size_t)t0 * 2 + 1) * 128)[lane];
  for (int t = t0 + 1; t <= t1; t++) {
    const float4 u = reinterpret_cast<const float4*>(part + ((size_t)t * 2 + 0) * 128)[lane];
    v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
  }
  *dst = v;
}

}  // namespace

// mode: 0 first (layer-0 prologue), 1 mid (after layer `layer`, prepares layer+1), 2 last (after the last layer)
int node_update_tc_launch(gamd_ctx* ctx, int mode, int layer, const float4* pos_feat, int64_t n_atoms, cudaStream_t st) {
  const size_t smem = sizeof(SmemNode) + 1024;
  if (!(ctx->attr_mask & GAMD_ATTR_NODE_TC)) {
    GAMD_CUDA(cudaFuncSetAttribute(k_node_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->attr_mask |= GAMD_ATTR_NODE_TC;
  }
  const ModelW& mw = ctx->mw;
  if (mode != MODE_FIRST) {
    k_agg_fixup<<<ceil_div(n_atoms * 32, 256), 256, 0, st>>>(ctx->row_ptr, ctx->part, ctx->agg, (int)n_atoms, ctx->n_edges);
    GAMD_LAUNCH_CHECK();
  }
  NodeTcArgs a{};
  const size_t slab = (size_t)6 * 2 * WCHUNK;
  const int cur = mode == MODE_FIRST ? 0 : layer, nxt = mode == MODE_FIRST ? 0 : layer + 1;
  a.w_img = ctx->d_wimg_node + (size_t)cur * slab;
  a.w_img_next = ctx->d_wimg_node + (size_t)(nxt < mw.n_layers ? nxt : cur) * slab;
  if (mode != MODE_FIRST) {
    a.b_pedge = mw.layer[cur].pedge_b;
    a.b_phi = mw.layer[cur].phi_b;
  }
  if (mode != MODE_LAST) {
    const LayerW& L = mw.layer[nxt];
    a.b_src = L.src_b; a.b_dst = L.dst_b; a.b_pdst = L.pdst_b; a.ln_w = L.ln_w; a.ln_b = L.ln_b;
  } else {
    a.b_dec0 = mw.dec0_b; a.dec2_w = mw.dec2_w; a.dec2_b = mw.dec2_b;
  }
  a.node_emb = mw.node_emb; a.nenc_w = mw.nenc_w; a.nenc_b = mw.nenc_b;
  a.pos_feat = pos_feat;
  a.agg = ctx->agg;
  a.h = ctx->h; a.hn = ctx->hn; a.srcA = ctx->srcA; a.dstA = ctx->dstA; a.pd = ctx->pd; a.pred = ctx->pred;
  a.n_atoms = (int)n_atoms;
  a.mode = mode;
  if (mode != MODE_LAST && ctx->dd_push_armed) {
    a.perm = ctx->perm;
    a.push_n = (int)ctx->dd_push_n;
    for (int side = 0; side < 2; side++) {
      a.push_slot[side] = ctx->dd_push_slot[side];
      a.push_rows[side] = ctx->dd_push_slot[side] ? ctx->dd_push_rows[side] : nullptr;
    }
    ctx->dd_push_armed = false;     // one layer per arming
  }
  // the node-sized GEMMs always run the 3-pass split: they cost ~4 % of the edge-sized ones, and single-pass bf16
  // there alone pushes the force error from 2e-3 to 1.3e-2
  a.exact = 1;
  const int ntiles = ceil_div(n_atoms, TILE);
  const int npairs = (ntiles + 1) / 2;
  k_node_tc<<<npairs < ctx->sm_count ? npairs : ctx->sm_count, THREADS, smem, st>>>(a);
  GAMD_LAUNCH_CHECK();
  return 0;
}
