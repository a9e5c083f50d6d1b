// Stage 2 hot kernel on the 5th-generation tensor cores: one message-passing layer's edge chain
//
//   e_emb = theta_edge( edge_affine(e) + src_affine(hn)[src] + dst_affine(hn)[dst] )
//   agg[dst] = sum over the receiver-sorted edge run of hn[src] * e_emb
//
// (code/nn_module.py:135-142) as FOUR chained 128x128x128 GEMMs per 128-edge tile, issued with
// tcgen05.mma (kind::f16, bf16 operands, fp32 accumulation in TMEM).  The A operand of every GEMM
// lives in TMEM (written by the epilogue of the previous GEMM with tcgen05.st), the weights in shared
// memory (SWIZZLE_128B K-major images fetched with 1-D bulk async copies), the accumulator in TMEM.
//
//   precision "bf16x3": x = hi + lo (two bf16), D = Ahi*Bhi + Alo*Bhi + Ahi*Blo  -> fp32-grade result
//   precision "bf16"  : single pass on the hi parts
//
// CTA = 18 warps: 16 epilogue warps (thread = edge row; two 128-edge tiles in flight, 8 warps each = 4 TMEM
// lane quadrants x 2 column halves; the two tiles ping-pong on the tensor core and share every weight
// load), 1 MMA-issue warp, 1 weight-producer warp.  Neighbour features are gathered with coalesced cp.async into per-warp staging rows; the final
// segmented sum walks the receiver-sorted rows in order (messages transposed through the staging tile, lane =
// feature column) - no atomics, deterministic; rows that straddle a 32-edge block go to the `part` side buffer
// and are summed (in order) by the node kernel.
#include "common.cuh"
#include "tc_common.cuh"
#include <cstdlib>

namespace {
using namespace tc;

constexpr int TILE = 128;
constexpr int WCHUNK = 32768;        // one weight part image (128 x 128 bf16)
constexpr int RING = 4;
constexpr int GROW = 80;             // staging row stride in bytes (64 data + 16 pad: conflict-free LDS.128)
constexpr int EPI_WARPS = 16;        // 2 tiles in flight x 4 lane quadrants x 2 column halves
constexpr int MMA_WARP = EPI_WARPS;   // warp EPI_WARPS + 1 is the weight producer
constexpr int THREADS = (EPI_WARPS + 2) * 32;

struct __align__(1024) SmemTC {
  uint8_t w[RING][WCHUNK];
  uint8_t gather[EPI_WARPS][2][32 * GROW];
  float bias[4][128];
  uint64_t full[RING], empty[RING], a_ready[2], d_ready[2];
  uint32_t tmem_base;
};

struct MpTcArgs {
  const uint8_t* w_img;   // [4 stages][2 parts][WCHUNK]
  const float* bias;      // [4][128]
  const uint8_t* e_blob;  // [ntiles][2][32 KB]: chunk-major rows (see edge encoder)
  const int *row_ptr, *col, *edst, *n_edges;
  const float *hn, *srcA, *dstA;
  float *agg, *part;
  const int *tile_list, *n_list;   // optional: process only these tiles (domain decomposition: interior / boundary)
  int exact;
  long long* dbg;   // development: clock64 timeline of CTA 0 (nullptr = off)
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float silu_fast(float x) {
  // x * sigmoid(x) with ex2.approx / rcp.approx (both ~1 ulp): |err| ~ 2e-7 relative
  float t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-1.4426950408889634f * x));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + t));
  return x * r;
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void cp_async16s(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gmem) : "memory");
}

// per-thread state of an epilogue thread for the tile it is working on
struct EpiCtx {
  uint32_t Dc, AH, AL;          // TMEM addresses (my lane quadrant, my 64-column half of my tile slot)
  uint32_t gbuf[2];             // shared addresses of my warp's two gather staging buffers
  uint32_t grow_off;            // my row inside a staging buffer (lane * GROW)
  uint32_t gl_dst;              // cp.async destination offset of this lane: (lane>>2)*GROW + (lane&3)*16
  int gl_row, gl_col;           // cp.async source row (lane>>2) and float offset ((lane&3)*4)
  uint32_t bias_addr;           // shared address of bias[0][col0]
  int col0, src;
  bool valid;
  uint32_t end_mask;            // bit j: row j of my warp's 32 rows is the last edge of a receiver run (in this block)
  float* out_row;
  const float4* dst_row;
  const float *srcA, *hn;
};

// gather #gi of the current tile: gi = 0..7 -> (array = gi<4 ? srcA : hn, 16-column chunk = gi&3);
// 64 B of each of my warp's 32 neighbour rows, coalesced: 4 lanes per row, 8 rows per instruction
__device__ __forceinline__ void issue_gather(const EpiCtx& c, int gi) {
  const float* base = (gi < 4 ? c.srcA : c.hn) + c.col0 + (gi & 3) * 16 + c.gl_col;
  const uint32_t dst = c.gbuf[gi & 1] + c.gl_dst;
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int sj = __shfl_sync(0xffffffffu, c.src, it * 8 + c.gl_row);
    cp_async16s(dst + it * 8 * GROW, base + (size_t)sj * 128);
  }
  cp_async_commit();
}

__device__ __forceinline__ float silu_tanh(float x) {
  // x * sigmoid(x) = 0.5 x (1 + tanh(x/2)) with one MUFU.TANH (rel. error 2^-11: bf16 "fast" mode only)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// x -> silu(x) for two elements, split into bf16 hi / lo pairs; packed fp32x2 arithmetic halves the FMA-pipe
// instruction count of the (issue-bound) activation stages
__device__ __forceinline__ void silu_split_pair(f32x2 X, uint32_t& hi, uint32_t& lo) {
  const f32x2 T = mul2(X, pk2(-1.4426950408889634f, -1.4426950408889634f));
  float t0, t1;
  unpk2(T, t0, t1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
  const f32x2 U = add2(pk2(e0, e1), pk2(1.f, 1.f));
  float u0, u1;
  unpk2(U, u0, u1);
  float r0, r1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(u0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(u1));
  const f32x2 Y = mul2(X, pk2(r0, r1));
  float y0, y1;
  unpk2(Y, y0, y1);
  hi = pack_bf16(y0, y1);
  const f32x2 R = fma2(pk2u(hi << 16, hi & 0xffff0000u), pk2(-1.f, -1.f), Y);
  float q0, q1;
  unpk2(R, q0, q1);
  lo = pack_bf16(q0, q1);
}

// epilogue of GEMM stage S for my 64 columns (4 chunks of 16)
template <int S, bool EXACT>
__device__ __forceinline__ void stage_epilogue(const EpiCtx& c) {
  float4 dn[4];
  if (S == 1) {
#pragma unroll
    for (int i = 0; i < 4; i++) dn[i] = __ldg(c.dst_row + i);
  }
  uint32_t vbuf[2][16];
  tmem_ld16(c.Dc, vbuf[0]);
#pragma unroll
  for (int cc = 0; cc < 4; cc++) {
    // TMEM loads are pipelined one chunk ahead: wait for chunk cc, then put chunk cc+1 in flight
    tmem_wait_ld();
    if (cc < 3) tmem_ld16(c.Dc + (cc + 1) * 16, vbuf[(cc + 1) & 1]);
    const uint32_t(&v)[16] = vbuf[cc & 1];
    uint32_t grow = 0;
    float4 dc[4];
    if (S == 1 || S == 3) {
      // software pipeline over the 8 gathers of this tile: srcA chunks 0-3 (stage 1), hn chunks 0-3 (stage 3)
      const int gi = (S == 1 ? 0 : 4) + cc;
      if (gi + 1 < 8) {
        issue_gather(c, gi + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      grow = c.gbuf[gi & 1] + c.grow_off;
      if (S == 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) dc[i] = dn[i];
        if (cc < 3) {
#pragma unroll
          for (int i = 0; i < 4; i++) dn[i] = __ldg(c.dst_row + (cc + 1) * 4 + i);
        }
      }
    }
    if (S < 3 && EXACT) {
      // packed path: bias (+ per-edge terms) + SiLU + bf16 hi/lo split on fp32x2 pairs
      uint32_t h[8], l[8];
#pragma unroll
      for (int j4 = 0; j4 < 4; j4++) {
        const float4 b = lds128(c.bias_addr + (S * 128 + cc * 16 + j4 * 4) * 4);
        f32x2 X0 = add2(pk2u(v[4 * j4], v[4 * j4 + 1]), pk2(b.x, b.y));
        f32x2 X1 = add2(pk2u(v[4 * j4 + 2], v[4 * j4 + 3]), pk2(b.z, b.w));
        if (S == 1) {
          const float4 sv = lds128(grow + j4 * 16);
          X0 = add2(X0, add2(pk2(sv.x, sv.y), pk2(dc[j4].x, dc[j4].y)));
          X1 = add2(X1, add2(pk2(sv.z, sv.w), pk2(dc[j4].z, dc[j4].w)));
        }
        silu_split_pair(X0, h[2 * j4], l[2 * j4]);
        silu_split_pair(X1, h[2 * j4 + 1], l[2 * j4 + 1]);
      }
      tmem_st8(c.AH + cc * 8, h);
      tmem_st8(c.AL + cc * 8, l);
      continue;
    }
    float x[16];
#pragma unroll
    for (int j4 = 0; j4 < 4; j4++) {
      const float4 b = lds128(c.bias_addr + (S * 128 + cc * 16 + j4 * 4) * 4);
      x[4 * j4] = __uint_as_float(v[4 * j4]) + b.x;
      x[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + b.y;
      x[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + b.z;
      x[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + b.w;
    }
    if (S == 1) {
#pragma unroll
      for (int j4 = 0; j4 < 4; j4++) {
        const float4 sv = lds128(grow + j4 * 16);
        x[4 * j4] += sv.x + dc[j4].x;
        x[4 * j4 + 1] += sv.y + dc[j4].y;
        x[4 * j4 + 2] += sv.z + dc[j4].z;
        x[4 * j4 + 3] += sv.w + dc[j4].w;
      }
    }
    if (S < 3) {
      uint32_t h[8], l[8];
      if (EXACT) {
#pragma unroll
        for (int j = 0; j < 8; j++) split_bf16(silu_fast(x[2 * j]), silu_fast(x[2 * j + 1]), h[j], l[j]);
        tmem_st8(c.AH + cc * 8, h);
        tmem_st8(c.AL + cc * 8, l);
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) h[j] = pack_bf16(silu_tanh(x[2 * j]), silu_tanh(x[2 * j + 1]));
        tmem_st8(c.AH + cc * 8, h);
      }
    } else {
      // message = hn[src] * e_emb; parked (fp32) in my own accumulator columns until all four chunks are done
      uint32_t pr[16];
#pragma unroll
      for (int j4 = 0; j4 < 4; j4++) {
        const float4 hv = lds128(grow + j4 * 16);
        pr[4 * j4] = __float_as_uint(c.valid ? x[4 * j4] * hv.x : 0.f);
        pr[4 * j4 + 1] = __float_as_uint(c.valid ? x[4 * j4 + 1] * hv.y : 0.f);
        pr[4 * j4 + 2] = __float_as_uint(c.valid ? x[4 * j4 + 2] * hv.z : 0.f);
        pr[4 * j4 + 3] = __float_as_uint(c.valid ? x[4 * j4 + 3] * hv.w : 0.f);
      }
      tmem_st16(c.Dc + cc * 16, pr);
    }
  }
  if (S == 3) {
    // segmented sum over the receiver-sorted rows, without atomics and without shuffles: the 32 x 32 block of
    // messages is transposed through my warp's (now idle) staging buffers, lane = feature column walks down the
    // rows in order and stores a finished receiver's 32 sums as one coalesced 128-byte row segment
    tmem_wait_st();
    const uint32_t T = c.gbuf[0];              // 32 rows x 36 words (both staging buffers, 4608 B)
    const uint32_t lane = c.grow_off / GROW;
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
      uint32_t v0[16], v1[16];
      tmem_ld16(c.Dc + p * 32, v0);
      tmem_ld16(c.Dc + p * 32 + 16, v1);
      tmem_wait_ld();
      __syncwarp();
#pragma unroll
      for (int j4 = 0; j4 < 4; j4++) {
        sts128(T + lane * 144 + j4 * 16, v0[4 * j4], v0[4 * j4 + 1], v0[4 * j4 + 2], v0[4 * j4 + 3]);
        sts128(T + lane * 144 + 64 + j4 * 16, v1[4 * j4], v1[4 * j4 + 1], v1[4 * j4 + 2], v1[4 * j4 + 3]);
      }
      __syncwarp();
      float acc = 0.f;
#pragma unroll
      for (int jb = 0; jb < 32; jb += 16) {
        float t[16];   // loads batched ahead of the (serial) add chain
#pragma unroll
        for (int j = 0; j < 16; j++) t[j] = lds32(T + (jb + j) * 144 + lane * 4);
#pragma unroll
        for (int j = 0; j < 16; j++) {
          acc += t[j];
          if ((c.end_mask >> (jb + j)) & 1u) {
            const unsigned long long ptr = __shfl_sync(0xffffffffu, (unsigned long long)c.out_row, jb + j);
            reinterpret_cast<float*>(ptr)[p * 32 + lane] = acc;
            acc = 0.f;
          }
        }
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(THREADS, 1) k_mp_edge_tc(MpTcArgs a) {
  extern __shared__ __align__(1024) uint8_t raw[];
  SmemTC& sm = *reinterpret_cast<SmemTC*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int E = *a.n_edges;
  // tiles are addressed by SLOT: slot -> tile is the identity, or a look-up in the caller's tile list
  const int ntiles = a.tile_list ? *a.n_list : (E + TILE - 1) / TILE;
  const int npairs = (ntiles + 1) / 2;

  if (warp == MMA_WARP) tmem_alloc(&sm.tmem_base, 512);
  if (tid == 0) {
    for (int i = 0; i < RING; i++) {
      mbar_init(&sm.full[i], 1);
      mbar_init(&sm.empty[i], 2);   // both tile slots must have consumed a weight stage before it is replaced
    }
    for (int g = 0; g < 2; g++) {
      mbar_init(&sm.a_ready[g], 8);        // one arrival per epilogue warp of the tile
      mbar_init(&sm.d_ready[g], 1);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 4 * 128; i += THREADS) (&sm.bias[0][0])[i] = a.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = sm.tmem_base;

  if (warp < EPI_WARPS) {
    // ===== epilogue warps: thread = edge row; warp = (tile slot g, column half ch, lane quadrant wq) =====
    const int g = warp >> 3, ch = (warp >> 2) & 1, wq = warp & 3;
    const int r = wq * 32 + lane;
    const bool exact = a.exact != 0;
    EpiCtx c;
    c.col0 = ch * 64;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    c.Dc = tb + lane_base + g * 256 + c.col0;
    c.AH = tb + lane_base + g * 256 + 128 + c.col0 / 2;
    c.AL = c.AH + 64;
    c.gbuf[0] = smem_u32(sm.gather[warp][0]);
    c.gbuf[1] = smem_u32(sm.gather[warp][1]);
    c.grow_off = lane * GROW;
    c.gl_row = lane >> 2;
    c.gl_col = (lane & 3) * 4;
    c.gl_dst = (lane >> 2) * GROW + (lane & 3) * 16;
    c.bias_addr = smem_u32(&sm.bias[0][c.col0]);
    c.srcA = a.srcA;
    c.hn = a.hn;
    uint64_t* const a_bar = &sm.a_ready[g];
    uint64_t* const d_bar = &sm.d_ready[g];
    uint32_t d_par = 0;
    long long* dbg_rec = a.dbg ? a.dbg + warp * 256 : nullptr;
    int dbg_n = 0;
    // endpoints of my edge in the first tile; the next tile's are fetched while the current one is processed
    int nsrc = 0, ndst = -1;
    int next_tile = -1;
    {
      const int s0 = blockIdx.x * 2 + g;
      if (s0 < ntiles) next_tile = a.tile_list ? __ldg(a.tile_list + s0) : s0;
      const int e_first = next_tile * TILE + r;
      if (next_tile >= 0 && e_first < E) {
        nsrc = a.col[e_first];
        ndst = a.edst[e_first];
      }
    }
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int slot = pair * 2 + g;
      if (slot >= ntiles) continue;
      const int tile = next_tile;
      const bool dbg_on = dbg_rec && blockIdx.x == 0 && lane == 0 && dbg_n + 14 <= 256;
      if (dbg_on) dbg_rec[dbg_n++] = clock64();
      const int e0 = tile * TILE;
      const int e = e0 + r;
      c.valid = e < E;
      c.src = nsrc;
      const int dst = ndst;
      // e tile loads first: they depend on nothing but the tile index
      const uint4* bh = reinterpret_cast<const uint4*>(a.e_blob + (size_t)tile * 65536) + (ch * 8) * 128 + r;
      uint4 q[8];
#pragma unroll
      for (int i = 0; i < 8; i++) q[i] = __ldg(bh + i * 128);
      {  // the next tile of this slot and the endpoints of my edge in it
        const int nslot = slot + 2 * (int)gridDim.x;
        next_tile = -1;
        if (nslot < ntiles) next_tile = a.tile_list ? __ldg(a.tile_list + nslot) : nslot;
        const int en = next_tile * TILE + r;
        nsrc = 0;
        ndst = -1;
        if (next_tile >= 0 && en < E) {
          nsrc = a.col[en];
          ndst = a.edst[en];
        }
      }
      c.dst_row = reinterpret_cast<const float4*>(a.dstA + (size_t)(dst < 0 ? 0 : dst) * 128 + c.col0);
      issue_gather(c, 0);
      {  // pull my share of the NEXT tile's e blob into L2 while this tile is being processed
        const int ntile = next_tile;
        if (ntile >= 0 && (lane & 7) == 0) {
          const uint8_t* nb = a.e_blob + (size_t)ntile * 65536 + ((size_t)(ch * 8) * 128 + r) * 16;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + i * 2048));
            if (exact) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + 32768 + i * 2048));
          }
        }
      }

      // ---- stage 0 operand: my half of the e tile (bf16 hi / lo) -> TMEM A ------------------
      {
#pragma unroll
        for (int part = 0; part < 2; part++) {
          if (part == 1) {
            if (!exact) break;
#pragma unroll
            for (int i = 0; i < 8; i++) q[i] = __ldg(bh + 32768 / 16 + i * 128);
          }
#pragma unroll
          for (int c2 = 0; c2 < 2; c2++) {
            uint32_t h[16];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              h[4 * i] = q[c2 * 4 + i].x; h[4 * i + 1] = q[c2 * 4 + i].y;
              h[4 * i + 2] = q[c2 * 4 + i].z; h[4 * i + 3] = q[c2 * 4 + i].w;
            }
            tmem_st16((part ? c.AL : c.AH) + c2 * 16, h);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_bar);
      }
      if (dbg_on) dbg_rec[dbg_n++] = clock64();

      // receiver runs inside my warp's 32 rows (receiver-sorted): where they end and where their sums go
      const int nd = __shfl_down_sync(0xffffffffu, dst, 1);
      const bool seg_end = c.valid && (lane == 31 || nd != dst);
      c.end_mask = __ballot_sync(0xffffffffu, seg_end);
      c.out_row = nullptr;
      if (seg_end) {
        const int bstart = e0 + wq * 32;
        const int bend = min(bstart + 32, E);
        const int blk = bstart >> 5;
        if (a.row_ptr[dst] < bstart) c.out_row = a.part + ((size_t)blk * 2 + 0) * 128;
        else if (a.row_ptr[dst + 1] > bend) c.out_row = a.part + ((size_t)blk * 2 + 1) * 128;
        else c.out_row = a.agg + (size_t)dst * 128;
        c.out_row += c.col0;
      }

#define GAMD_STAGE(S)                              \
  if (dbg_on) dbg_rec[dbg_n++] = clock64();        \
  mbar_wait(d_bar, d_par);                         \
  d_par ^= 1;                                      \
  tc_fence_after();                                \
  if (dbg_on) dbg_rec[dbg_n++] = clock64();        \
  if (exact) stage_epilogue<S, true>(c);           \
  else stage_epilogue<S, false>(c);                \
  if (S < 3) {                                     \
    tmem_wait_st();                                \
    tc_fence_before();                             \
    __syncwarp();                                  \
    if (lane == 0) mbar_arrive(a_bar);             \
  }                                                \
  if (dbg_on) dbg_rec[dbg_n++] = clock64();
      GAMD_STAGE(0)
      GAMD_STAGE(1)
      GAMD_STAGE(2)
      GAMD_STAGE(3)
#undef GAMD_STAGE
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issue (one thread): an event loop over the two tiles in flight =====================
    // (A single thread issues a tcgen05.mma only every ~120 cycles - profiles/probes/tc_ubench.cu - so a 24-MMA stage costs ~2.9k
    // cycles of issue.  Variants with one issuer per tile slot, two issuers per stage, or K steps issued as the
    // previous epilogue publishes 16-column chunks were all measured SLOWER: they let the two tiles fall into
    // lockstep on the MUFU pipe; two issuers per accumulator also lose run-to-run determinism.)
    // Each tile slot walks its own sequence of (pair, stage) steps Q = 4 * pair_iteration + stage and is served as
    // soon as its A operand is in TMEM and the weights of that stage are in shared memory, independently of the
    // other slot.  bf16x3: weights live in two slot pairs (hi, lo) indexed by Q & 1; a pair is reloaded with stage
    // Q + 2 once BOTH tiles have consumed stage Q (wempty counts 2), so the slots may drift apart by one stage.
    // The whole warp runs the loop on warp-uniform values; only the MMAs, commits and arrives are predicated on one
    // elected lane (see tc_common.cuh: issuing from inside `if (lane == 0)` halves the MMA issue rate).
    {
      const uint32_t leader = elect_leader();
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const int n_my_pairs = blockIdx.x < npairs ? (npairs - 1 - blockIdx.x) / gridDim.x + 1 : 0;
      const int totalQ = 4 * n_my_pairs;
      int Qg[2] = {0, 0};
      uint32_t a_par[2] = {0, 0};
      bool w_res = false;   // bf16 mode: all four hi images resident after the first load
      uint32_t spins = 0;
      while (Qg[0] < totalQ || Qg[1] < totalQ) {
        bool progressed = false;
#pragma unroll
        for (int g = 0; g < 2; g++) {
          const int Q = Qg[g];
          if (Q >= totalQ) continue;
          const int s = Q & 3;
          const int pair = blockIdx.x + (Q >> 2) * gridDim.x;
          const bool tile_valid = pair * 2 + g < ntiles;
          const int sp = Q & 1;
          if (a.exact) {
            if (!__all_sync(0xffffffffu, mbar_test_wait(&sm.full[sp], (Q >> 1) & 1))) continue;
          } else if (!w_res) {
            if (!__all_sync(0xffffffffu, mbar_test_wait(&sm.full[0], 0))) continue;
            w_res = true;
          }
          if (!tile_valid) {          // absent second tile of the tail pair: release the weights on its behalf
            if (a.exact && leader) mbar_arrive(&sm.empty[sp]);
            Qg[g]++;
            progressed = true;
            continue;
          }
          if (!__all_sync(0xffffffffu, mbar_test_wait(&sm.a_ready[g], a_par[g]))) continue;
          a_par[g] ^= 1;
          tc_fence_after();
          const uint32_t bhi = smem_u32(a.exact ? sm.w[2 * sp] : sm.w[s]);
          const uint32_t blo = smem_u32(sm.w[2 * sp + 1]);
          const uint32_t d = tb + g * 256, ah = d + 128, al = d + 192;
          const int passes = a.exact ? 3 : 1;
          uint32_t accum = 0;
          for (int p = 0; p < passes; p++) {
            const uint32_t bb = (p == 2) ? blo : bhi;
            const uint32_t aa = (p == 1) ? al : ah;
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
              umma_ts_elect(d, aa + ks * 8, umma_desc_sw128(bb + (ks >> 2) * 16384 + (ks & 3) * 32), idesc, accum,
                            leader);
              accum = 1;
            }
          }
          if (leader) {
            umma_commit(&sm.d_ready[g]);
            if (a.exact) umma_commit(&sm.empty[sp]);
          }
          __syncwarp();
          Qg[g]++;
          progressed = true;
        }
        if (progressed) spins = 0;
        else if (++spins > (1u << 26)) __trap();
      }
    }
    __syncwarp();
  } else {
    // ===================== weight producer (one thread) =====================
    if (lane == 0) {
      const int n_my_pairs = blockIdx.x < npairs ? (npairs - 1 - blockIdx.x) / gridDim.x + 1 : 0;
      if (!a.exact) {
        if (n_my_pairs) {
          mbar_arrive_expect_tx(&sm.full[0], 4 * WCHUNK);
          for (int s = 0; s < 4; s++)
#pragma unroll
            for (int i = 0; i < 4; i++)
              bulk_g2s(sm.w[s] + i * 8192, a.w_img + (size_t)s * 2 * WCHUNK + i * 8192, 8192, &sm.full[0]);
        }
      } else {
        const int totalQ = 4 * n_my_pairs;
        for (int Q = 0; Q < totalQ; Q++) {
          const int sp = Q & 1, n = Q >> 1;
          if (n >= 1) mbar_wait(&sm.empty[sp], (n - 1) & 1);
          mbar_arrive_expect_tx(&sm.full[sp], 2 * WCHUNK);
          const uint8_t* src = a.w_img + (size_t)(Q & 3) * 2 * WCHUNK;     // [hi | lo] of this stage, contiguous
#pragma unroll
          for (int i = 0; i < 8; i++) bulk_g2s(sm.w[2 * sp] + i * 8192, src + i * 8192, 8192, &sm.full[sp]);
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tb, 512);
}

}  // namespace

int mp_edge_tc_launch(gamd_ctx* ctx, int layer, cudaStream_t st, int which) {
  // GAMD_MP_VARIANT (read at gamd_create): 3 = three tiles in flight (mp_tc3.cu), 4 = the same without the commit wait
  if (ctx->mp_variant == 3 || ctx->mp_variant == 4) return mp_edge_tc3_launch(ctx, layer, st, which, ctx->mp_variant == 3);
  // 5 = CTA pairs (cta_group::2) with resident weights, three tiles in flight (mp_tc2cta.cu); 6 = without the commit wait
  // launch-bound systems (at most a tile or two per SM: LJ-258, TIP3P-774) run the two-tile single-CTA kernel below: its
  // tile chain is shorter (33 k cycles against 46 k), and with one tile per CTA only the chain length counts
  const bool tiny = ctx->model_atoms > 0 && ctx->model_atoms <= ctx->mp_small_atoms;
  // 7 = 6 with every GEMM issued as two N = 64 halves (separate commits: the epilogue starts on the first half)
  // 8 = 6 with the one-round-trip tile set-up (hi part through the staging rows, lo part through registers)
  if (ctx->mp_variant >= 5 && ctx->mp_variant <= 12 && !tiny)
    return mp_edge_tc2_launch(ctx, layer, st, which, ctx->mp_variant == 5, ctx->mp_variant == 7);
  const size_t smem = sizeof(SmemTC) + 1024;
  if (!(ctx->attr_mask & GAMD_ATTR_MP_TC)) {
    GAMD_CUDA(cudaFuncSetAttribute(k_mp_edge_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->attr_mask |= GAMD_ATTR_MP_TC;
  }
  MpTcArgs a;
  a.w_img = ctx->d_wimg + (size_t)layer * 8 * WCHUNK;
  a.bias = ctx->d_tc_bias + (size_t)layer * 4 * 128;
  a.e_blob = reinterpret_cast<const uint8_t*>(ctx->e_emb);
  a.row_ptr = ctx->row_ptr;
  a.col = ctx->col_idx;
  a.edst = ctx->edge_dst;
  a.n_edges = ctx->n_edges;
  a.hn = ctx->hn;
  a.srcA = ctx->srcA;
  a.dstA = ctx->dstA;
  a.agg = ctx->agg;
  a.part = ctx->part;
  a.tile_list = which >= 0 ? ctx->tile_list[which] : nullptr;
  a.n_list = which >= 0 ? ctx->tile_count + which : nullptr;
  a.exact = ctx->desc.precision == GAMD_PREC_BF16X3 ? 1 : 0;
  a.dbg = (ctx->dbg_timeline && layer == 1) ? reinterpret_cast<long long*>(ctx->e_emb + (size_t)ctx->cap_edges * 128) : nullptr;
  // interior launch of a tile-split layer: optionally leave a few SMs to the halo exchange running beside it
  const int reserve = ctx->dd_reserve_sms;
  const int grid = which == 0 && reserve > 0 && reserve < ctx->sm_count ? ctx->sm_count - reserve : ctx->sm_count;
  k_mp_edge_tc<<<grid, THREADS, smem, st>>>(a);
  GAMD_LAUNCH_CHECK();
  return 0;
}
