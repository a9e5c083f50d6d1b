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
// CTA PAIRS (this file): two CTAs of a cluster (the two SMs of a TPC) run tcgen05.mma.cta_group::2 with M = 256 -
// each CTA owns a 128-edge tile (its 128 TMEM lanes) and HALF of every weight matrix (64 of the 128 output columns'
// rows of B), so all four stages' hi and lo images (4 x 2 x 16 KB = 128 KB per CTA) stay RESIDENT in shared memory
// for the whole kernel: no weight ring, no producer traffic (the single-CTA kernels stream 128-256 KB of weights per
// tile from L2), and - the point - nothing couples the progress of the tiles in flight any more: with a 2-unit ring
// three tiles have to stay within two consecutive stages of each other, which cost the 3-tile single-CTA kernel
// (mp_tc3.cu) 40 % of its chain time in waits.
// Per CTA three tiles are in flight.  TMEM holds 4 blocks of 128 columns: three tile "homes" and one floating block.
// A tile's activations are written IN PLACE over its accumulator (16 fp32 columns -> 8 columns of bf16 hi pairs + 8 of
// lo pairs), its next GEMM reads them from the home and writes the floating block, which becomes the new home (the
// leader's MMA warp publishes it in both CTAs' shared memory before the commit); the old home is the next floating
// block.  The leader CTA (cluster rank 0) issues every MMA for the pair; the epilogue threads of both CTAs arrive on
// the leader's a_ready barrier (cluster-scope release), commits are multicast to both CTAs' d_ready barriers.
// CTA = 25 warps (80 registers): 24 epilogue warps (thread = edge row; 8 warps per tile = 4 TMEM lane quadrants x 2
// column halves) and 1 MMA-issue warp, which also fetches the weights once.  Neighbour features are gathered with coalesced cp.async
// into per-warp XOR-swizzled staging rows (no padding: 96 KB); the final
// segmented sum walks the receiver-sorted rows in order (messages transposed through the staging tile, lane =
// feature column) - no atomics, deterministic; rows that straddle a 32-edge block go to the `part` side buffer
// and are summed (in order) by the node kernel.
#include "common.cuh"
#include "tc_common.cuh"
#include <cstdlib>

namespace {
using namespace tc;

constexpr int TILE = 128;
constexpr int WPART = 16384;         // one weight part image of ONE CTA: 64 of the 128 rows of B (64 x 128 bf16)
constexpr int WTOTAL = 4 * 2 * WPART; // [stage][hi | lo] resident per CTA
constexpr int NSLOT = 3;             // tiles in flight
constexpr int GROW = 64;             // staging row stride in bytes (un-padded; 16-byte chunks XOR-swizzled)
constexpr int GBUF = 32 * GROW;      // one staging buffer: 32 rows x 64 B
constexpr int EPI_WARPS = 8 * NSLOT; // tiles in flight x 4 lane quadrants x 2 column halves
constexpr int MMA_WARP = EPI_WARPS;   // also loads the weights once at the start
constexpr int THREADS = (EPI_WARPS + 1) * 32;   // 800 threads: an 80-register budget (26 warps would cap it at 72)

struct __align__(1024) SmemTC {
  uint8_t w[4][2][WPART];            // my half (64 rows) of every stage's B image: [stage][hi | lo]
  uint8_t gather[EPI_WARPS][2][GBUF];
  float bias[4][128];
  uint64_t w_full, a_ready[NSLOT], d_ready[NSLOT][2];   // a_ready: leader CTA only (16 warp arrivals); d_ready[g][h]: N half h
                                                         // of the slot's accumulator (N-split kernels; else [g][0] only)
  volatile uint32_t home[NSLOT];     // TMEM column block (0..3) holding the accumulator of the slot's GEMM in flight
  uint32_t tmem_base;
};
static_assert(sizeof(SmemTC) <= 232448, "shared memory budget (227 KB per CTA)");

struct MpTcArgs {
  const uint8_t* w_img;   // [2 CTA halves][4 stages][2 parts][WPART]
  const float* bias;      // [4][128]
  const uint8_t* e_blob;  // [ntiles][2][32 KB]: chunk-major rows (see edge encoder)
  const int *row_ptr, *col, *edst, *n_edges;
  const float *hn, *srcA, *dstA;
  float *agg, *part;
  const int *tile_list, *n_list;   // optional: process only these tiles (domain decomposition: interior / boundary)
  int exact;
  int row_prefetch;        // L2 prefetch of the tile's sender rows at tile start (GAMD_MP_ROW_PREFETCH)
  uint32_t wait_hint_ns;   // suspend-time hint of the epilogue warps' accumulator waits (0 = plain poll loop)
  long long* dbg;   // development: clock64 timeline of CTA 0 (nullptr = off)
};

__device__ __forceinline__ long long gtime() {   // nanoseconds, comparable across SMs (clock64 is per SM)
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
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
__device__ __forceinline__ uint4 ldg_na(const uint4* p) {   // streaming read: do not displace the L1-resident dst_affine rows
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void cp_async16s(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gmem) : "memory");
}

// column offset (inside my warp's share of the 128 accumulator columns) of my chunk cc = 0..3.  Plain kernels: a warp
// owns 64 consecutive columns.  N-split kernels issue every GEMM as two N = 64 halves with separate commits; a warp
// owns 32 columns of EACH half (chunks 0, 1 in the first, 2, 3 in the second), so that its epilogue starts when the
// first half has landed and runs beside the second half's MMAs.
template <bool NSPLIT>
__device__ __forceinline__ constexpr int coff(int cc) {
  return NSPLIT ? (cc >> 1) * 64 + (cc & 1) * 16 : cc * 16;
}

// per-thread state of an epilogue thread for the tile it is working on
struct EpiCtx {
  uint64_t* d_bar1;             // N-split: barrier of the second accumulator half
  uint32_t d_par1;              // its parity for the current stage
  uint32_t wait_hint_ns;
  uint32_t Dc;                  // TMEM address: my lane quadrant, my 64-column half of my tile's current home block
  uint32_t gbuf[2];             // shared addresses of my warp's two gather staging buffers
  uint32_t grow_off;            // my row inside a staging buffer (lane * GROW)
  uint32_t gswz;                // XOR swizzle of my row's 16-byte chunks: (lane >> 1) & 3
  uint32_t lane;
  uint32_t gl_dst[4];           // cp.async destination offsets of this lane for rows (lane>>2) + 8*it (swizzled)
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
template <bool NSPLIT>
__device__ __forceinline__ void issue_gather(const EpiCtx& c, int gi) {
  const float* base = (gi < 4 ? c.srcA : c.hn) + c.col0 + coff<NSPLIT>(gi & 3) + c.gl_col;
  const uint32_t dst = c.gbuf[gi & 1];
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int sj = __shfl_sync(0xffffffffu, c.src, it * 8 + c.gl_row);
    cp_async16s(dst + c.gl_dst[it], base + (size_t)sj * 128);
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

// 2^t for two elements on the FMA / ALU pipes instead of MUFU.EX2 (the SiLU stages keep the 16-lane XU pipe 75-80 % busy
// while three tiles run their epilogues side by side): t clamped to [-126, 126], n = rint(t) by the magic-number add,
// 2^f on [-0.5, 0.5] as the degree-6 Taylor polynomial (relative error < 1.6e-7, ex2.approx: 2 ulp), the exponent
// spliced in with an integer shift-add.
__device__ __forceinline__ void ex2_poly_pair(float t0, float t1, float& e0, float& e1) {
  t0 = fminf(fmaxf(t0, -126.f), 126.f);
  t1 = fminf(fmaxf(t1, -126.f), 126.f);
  const f32x2 T = pk2(t0, t1);
  const f32x2 TM = add2(T, pk2(12582912.f, 12582912.f));                 // 1.5 * 2^23 + n
  const f32x2 N = add2(TM, pk2(-12582912.f, -12582912.f));
  const f32x2 F = fma2(N, pk2(-1.f, -1.f), T);
  f32x2 P = fma2(pk2(1.5403530393381606e-4f, 1.5403530393381606e-4f), F, pk2(1.3333558146428443e-3f, 1.3333558146428443e-3f));
  P = fma2(P, F, pk2(9.618129107628477e-3f, 9.618129107628477e-3f));
  P = fma2(P, F, pk2(5.550410866482158e-2f, 5.550410866482158e-2f));
  P = fma2(P, F, pk2(2.402265069591007e-1f, 2.402265069591007e-1f));
  P = fma2(P, F, pk2(6.931471805599453e-1f, 6.931471805599453e-1f));
  P = fma2(P, F, pk2(1.f, 1.f));
  float p0, p1, m0, m1;
  unpk2(P, p0, p1);
  unpk2(TM, m0, m1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(m0) << 23));
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(m1) << 23));
}

// x -> silu(x) for two elements, split into bf16 hi / lo pairs; packed fp32x2 arithmetic halves the FMA-pipe
// instruction count of the (issue-bound) activation stages.  POLY: the exponential on the FMA pipe (ex2_poly_pair)
template <bool POLY = false>
__device__ __forceinline__ void silu_split_pair(f32x2 X, uint32_t& hi, uint32_t& lo) {
  const f32x2 T = mul2(X, pk2(-1.4426950408889634f, -1.4426950408889634f));
  float t0, t1;
  unpk2(T, t0, t1);
  float e0, e1;
  if (POLY) {
    ex2_poly_pair(t0, t1, e0, e1);
  } else {
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
  }
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
template <int S, bool EXACT, bool NSPLIT, int POLY = 0>
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
    if (NSPLIT && cc == 1) {       // chunks 2, 3 are in the second N half of the accumulator
      mbar_wait_hint(c.d_bar1, c.d_par1, c.wait_hint_ns);
      tc_fence_after();
    }
    if (cc < 3) tmem_ld16(c.Dc + coff<NSPLIT>(cc + 1), vbuf[(cc + 1) & 1]);
    const uint32_t(&v)[16] = vbuf[cc & 1];
    uint32_t grow = 0;
    float4 dc[4];
    if (S == 1 || S == 3) {
      // software pipeline over the 8 gathers of this tile: srcA chunks 0-3 (stage 1), hn chunks 0-3 (stage 3).
      // Gathers 0 and 1 are issued at the start of the tile, gather gi + 2 once chunk gi has been consumed (it reuses
      // that staging buffer): two gathers are in flight at any time and the hn chunks 0, 1 land during stage 2.
      const int gi = (S == 1 ? 0 : 4) + cc;
      if (gi + 1 < 8) cp_async_wait<1>();
      else cp_async_wait<0>();
      __syncwarp();
      grow = c.gbuf[gi & 1] + c.grow_off;
      if (S == 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) dc[i] = dn[i];
        if (cc < 3) {
#pragma unroll
          for (int i = 0; i < 4; i++) dn[i] = __ldg(c.dst_row + coff<NSPLIT>(cc + 1) / 4 + i);
        }
      }
    }
    if (S < 3 && EXACT) {
      // packed path: bias (+ per-edge terms) + SiLU + bf16 hi/lo split on fp32x2 pairs
      uint32_t h[8], l[8];
#pragma unroll
      for (int j4 = 0; j4 < 4; j4++) {
        const float4 b = lds128(c.bias_addr + (S * 128 + coff<NSPLIT>(cc) + j4 * 4) * 4);
        f32x2 X0 = add2(pk2u(v[4 * j4], v[4 * j4 + 1]), pk2(b.x, b.y));
        f32x2 X1 = add2(pk2u(v[4 * j4 + 2], v[4 * j4 + 3]), pk2(b.z, b.w));
        if (S == 1) {
          const float4 sv = lds128(grow + ((j4 ^ c.gswz) << 4));
          X0 = add2(X0, add2(pk2(sv.x, sv.y), pk2(dc[j4].x, dc[j4].y)));
          X1 = add2(X1, add2(pk2(sv.z, sv.w), pk2(dc[j4].z, dc[j4].w)));
        }
        // POLY = 1: every second pair's exponential leaves the XU pipe (MUFU work -25 %), 2: every pair's (-50 %)
        silu_split_pair<POLY == 2>(X0, h[2 * j4], l[2 * j4]);
        silu_split_pair<POLY != 0>(X1, h[2 * j4 + 1], l[2 * j4 + 1]);
      }
      tmem_st8(c.Dc + coff<NSPLIT>(cc), h);        // in place over the accumulator chunk just read: [hi pairs | lo pairs]
      tmem_st8(c.Dc + coff<NSPLIT>(cc) + 8, l);
      if (S == 1) issue_gather<NSPLIT>(c, cc + 2);
      continue;
    }
    if (S == 3) {
      // message = hn[src] * e_emb on packed fp32x2, parked (fp32) in my own accumulator columns until all four chunks
      // are done.  Rows past the end of the edge list (tail of the last tile, phantom tiles) may hold anything: they
      // follow the last valid row of their warp, so the in-order segmented sum never stores a sum that includes them.
      uint32_t pr[16];
#pragma unroll
      for (int j4 = 0; j4 < 4; j4++) {
        const float4 b = lds128(c.bias_addr + (3 * 128 + coff<NSPLIT>(cc) + j4 * 4) * 4);
        const float4 hv = lds128(grow + ((j4 ^ c.gswz) << 4));
        const f32x2 M0 = mul2(add2(pk2u(v[4 * j4], v[4 * j4 + 1]), pk2(b.x, b.y)), pk2(hv.x, hv.y));
        const f32x2 M1 = mul2(add2(pk2u(v[4 * j4 + 2], v[4 * j4 + 3]), pk2(b.z, b.w)), pk2(hv.z, hv.w));
        float m0, m1, m2, m3;
        unpk2(M0, m0, m1);
        unpk2(M1, m2, m3);
        pr[4 * j4] = __float_as_uint(m0);
        pr[4 * j4 + 1] = __float_as_uint(m1);
        pr[4 * j4 + 2] = __float_as_uint(m2);
        pr[4 * j4 + 3] = __float_as_uint(m3);
      }
      tmem_st16(c.Dc + coff<NSPLIT>(cc), pr);
      if (cc < 2) issue_gather<NSPLIT>(c, 4 + cc + 2);
      continue;
    }
    float x[16];
#pragma unroll
    for (int j4 = 0; j4 < 4; j4++) {
      const float4 b = lds128(c.bias_addr + (S * 128 + coff<NSPLIT>(cc) + j4 * 4) * 4);
      x[4 * j4] = __uint_as_float(v[4 * j4]) + b.x;
      x[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + b.y;
      x[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + b.z;
      x[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + b.w;
    }
    if (S == 1) {
#pragma unroll
      for (int j4 = 0; j4 < 4; j4++) {
        const float4 sv = lds128(grow + ((j4 ^ c.gswz) << 4));
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
        tmem_st8(c.Dc + coff<NSPLIT>(cc), h);
        tmem_st8(c.Dc + coff<NSPLIT>(cc) + 8, l);
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) h[j] = pack_bf16(silu_tanh(x[2 * j]), silu_tanh(x[2 * j + 1]));
        tmem_st8(c.Dc + coff<NSPLIT>(cc), h);
      }
      if (S == 1) issue_gather<NSPLIT>(c, cc + 2);
    } else {
      // message = hn[src] * e_emb; parked (fp32) in my own accumulator columns until all four chunks are done
      uint32_t pr[16];
#pragma unroll
      for (int j4 = 0; j4 < 4; j4++) {
        const float4 hv = lds128(grow + ((j4 ^ c.gswz) << 4));
        pr[4 * j4] = __float_as_uint(c.valid ? x[4 * j4] * hv.x : 0.f);
        pr[4 * j4 + 1] = __float_as_uint(c.valid ? x[4 * j4 + 1] * hv.y : 0.f);
        pr[4 * j4 + 2] = __float_as_uint(c.valid ? x[4 * j4 + 2] * hv.z : 0.f);
        pr[4 * j4 + 3] = __float_as_uint(c.valid ? x[4 * j4 + 3] * hv.w : 0.f);
      }
      tmem_st16(c.Dc + coff<NSPLIT>(cc), pr);
      if (cc < 2) issue_gather<NSPLIT>(c, 4 + cc + 2);
    }
  }
  if (S == 3) {
    // segmented sum over the receiver-sorted rows, without atomics and without shuffles: the 32 x 32 block of
    // messages is transposed through my warp's (now idle) staging buffers, lane = feature column walks down the
    // rows in order and stores a finished receiver's 32 sums as one coalesced 128-byte row segment
    tmem_wait_st();
    const uint32_t T = c.gbuf[0];              // 32 rows x 128 B (both staging buffers, 4096 B), chunks XOR-swizzled
    const uint32_t lane = c.lane;
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
      uint32_t v0[16], v1[16];
      tmem_ld16(c.Dc + coff<NSPLIT>(2 * p), v0);
      tmem_ld16(c.Dc + coff<NSPLIT>(2 * p + 1), v1);
      tmem_wait_ld();
      __syncwarp();
#pragma unroll
      for (int j4 = 0; j4 < 4; j4++) {
        sts128(T + lane * 128 + ((j4 ^ (lane & 7)) << 4), v0[4 * j4], v0[4 * j4 + 1], v0[4 * j4 + 2], v0[4 * j4 + 3]);
        sts128(T + lane * 128 + (((4 + j4) ^ (lane & 7)) << 4), v1[4 * j4], v1[4 * j4 + 1], v1[4 * j4 + 2], v1[4 * j4 + 3]);
      }
      __syncwarp();
      float acc = 0.f;
#pragma unroll
      for (int jb = 0; jb < 32; jb += 16) {
        float t[16];   // loads batched ahead of the (serial) add chain
#pragma unroll
        for (int j = 0; j < 16; j++)
          t[j] = lds32(T + (jb + j) * 128 + (((lane >> 2) ^ ((jb + j) & 7)) << 4) + (lane & 3) * 4);
#pragma unroll
        for (int j = 0; j < 16; j++) {
          acc += t[j];
          if ((c.end_mask >> (jb + j)) & 1u) {
            const unsigned long long ptr = __shfl_sync(0xffffffffu, (unsigned long long)c.out_row, jb + j);
            reinterpret_cast<float*>(ptr)[(NSPLIT ? p * 64 : p * 32) + lane] = acc;
            acc = 0.f;
          }
        }
      }
    }
    __syncwarp();
  }
}

// STATIC: the leader serves the slots in a fixed round-robin order (slot 0, 1, 2 of stage s, then stage s + 1 ...) instead
// of "whichever is ready".  GEMM number n of a CTA pair then always writes TMEM block (3 + n) & 3, so every epilogue
// thread knows its accumulator block from (stage, slot, slots active in its group) alone: the leader no longer publishes
// the block through shared memory of both CTAs (remote store + cluster fence per GEMM) and polls ONE barrier at CTA
// scope instead of three at cluster scope - 0.55 us of leader time per GEMM, which paced the whole pair (the slots
// run in lock step, twelve GEMMs back to back per tile round: profiles/experiments/README.md).
// STAMP: dynamic service order (whichever slot is ready), but the block hand-over carries a sequence stamp instead of
// being ordered by a cluster fence: the leader writes (GEMM count of the slot << 8 | block) to both CTAs' shared memory
// with plain stores, an epilogue thread re-reads it until the stamp is the one it expects (normally at once: the store
// is more than a microsecond older than the commit that releases the thread); the leader polls at CTA scope.
template <bool SAFE_WAR, bool NSPLIT, bool SETUP1, int POLY = 0, bool STATIC = false, bool STAMP = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) k_mp_edge_tc2(MpTcArgs a) {
  extern __shared__ __align__(1024) uint8_t raw[];
  SmemTC& sm = *reinterpret_cast<SmemTC*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0 && (smem_u32(raw) & 1023u)) __trap();   // SWIZZLE_128B weight images need 1024-byte alignment
  const uint32_t rank = cluster_ctarank();             // 0 = leader (issues the MMAs of the pair)
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int E = *a.n_edges;
  // tiles are addressed by SLOT: slot -> tile is the identity, or a look-up in the caller's tile list.  A SUPER slot
  // is a pair of slots (2 q, 2 q + 1) worked on by the CTA pair as one M = 256 tile; the odd one may not exist (tail):
  // the peer CTA then runs the phantom tile just past the edge list (every row invalid, no store).
  const int ntiles = a.tile_list ? *a.n_list : (E + TILE - 1) / TILE;
  const int phantom = (E + TILE - 1) / TILE;
  const int nsuper = (ntiles + 1) / 2;
  const int ngroups = (nsuper + NSLOT - 1) / NSLOT;      // a CTA pair works on NSLOT consecutive super slots at a time

  if (warp == MMA_WARP) tmem_alloc2(&sm.tmem_base, 512);
  if (tid == 0) {
    mbar_init(&sm.w_full, 1);
    for (int g = 0; g < NSLOT; g++) {
      mbar_init(&sm.a_ready[g], 16);      // one arrival per epilogue warp of the slot, BOTH CTAs
      mbar_init(&sm.d_ready[g][0], 1);
      mbar_init(&sm.d_ready[g][1], 1);
      sm.home[g] = g;
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 4 * 128; i += THREADS) (&sm.bias[0][0])[i] = a.bias[i];
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP && lane == 0) {
    // my half of the layer's weights, resident for the whole kernel
    mbar_arrive_expect_tx(&sm.w_full, WTOTAL);
    const uint8_t* src = a.w_img + (size_t)rank * WTOTAL;
    for (int i = 0; i < WTOTAL / 8192; i++) bulk_g2s(&sm.w[0][0][0] + i * 8192, src + i * 8192, 8192, &sm.w_full);
  }
  mbar_wait(&sm.w_full, 0);
  // both CTAs: barriers initialised, TMEM allocated, weights resident - before any remote arrive or pair MMA
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = sm.tmem_base;

  if (warp < EPI_WARPS) {
    // ===== epilogue warps: thread = edge row; warp = (tile slot g, column half ch, lane quadrant wq) =====
    const int g = warp >> 3, ch = (warp >> 2) & 1, wq = warp & 3;
    const int r = wq * 32 + lane;
    const bool exact = a.exact != 0;
    EpiCtx c;
    c.col0 = NSPLIT ? ch * 32 : ch * 64;
    const uint32_t lane_base = ((uint32_t)(wq * 32) << 16) + c.col0;
    c.Dc = tb + lane_base + g * 128;          // the slot's first home block is block g
    c.gbuf[0] = smem_u32(sm.gather[warp][0]);
    c.gbuf[1] = smem_u32(sm.gather[warp][1]);
    c.lane = lane;
    c.grow_off = lane * GROW;
    c.gswz = (lane >> 1) & 3;
    c.gl_row = lane >> 2;
    c.gl_col = (lane & 3) * 4;
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int row = it * 8 + (lane >> 2);
      c.gl_dst[it] = row * GROW + ((((uint32_t)lane & 3u) ^ (((uint32_t)row >> 1) & 3u)) << 4);
    }
    c.bias_addr = smem_u32(&sm.bias[0][c.col0]);
    c.srcA = a.srcA;
    c.hn = a.hn;
    const uint32_t a_bar = mapa_u32(smem_u32(&sm.a_ready[g]), 0);   // the LEADER's barrier (shared::cluster address)
    uint64_t* const d_bar = &sm.d_ready[g][0];
    c.d_bar1 = &sm.d_ready[g][1];
    c.wait_hint_ns = a.wait_hint_ns;
    uint32_t d_par = 0;
    long long* dbg_rec = a.dbg ? a.dbg + (rank * 32 + warp) * 256 : nullptr;   // development timeline of cluster 0
    int dbg_n = 0;
    // my tile of super slot q: slot 2 q + rank, or the phantom tile
    auto tile_of = [&](int sslot) -> int {
      if (sslot >= nsuper) return -1;
      const int slot = 2 * sslot + (int)rank;
      if (slot >= ntiles) return phantom;
      return a.tile_list ? __ldg(a.tile_list + slot) : slot;
    };
    int next_tile = tile_of(cid * NSLOT + g);
    uint32_t nstage = 0;     // GEMMs of my slot so far (the stamp of the block hand-over, STAMP kernels)
    for (int grp = cid; grp < ngroups; grp += ncl) {
      const int sslot = grp * NSLOT + g;
      if (sslot >= nsuper) {
        // absent slot of the tail group.  Under the fixed service order the tail group's GEMMs rotate through all four
        // blocks, this slot's last home included: tell the leader that the previous tile's sums have left it
        if (STATIC) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(a_bar);
        }
        continue;
      }
      const uint32_t nact = (uint32_t)min(NSLOT, nsuper - grp * NSLOT);   // slots at work in this group (< NSLOT: tail)
      const int tile = next_tile;
      const bool dbg_on = dbg_rec && cid == 0 && lane == 0 && dbg_n + 14 <= 256;
      if (dbg_on) dbg_rec[dbg_n++] = gtime();
      const int e0 = tile * TILE;
      const int e = e0 + r;
      c.valid = e < E;
      // 16-byte chunk i = 0..7 of my row's share of the e tile: K step (i >> 1) of my four, see coff()
      const uint4* bh = reinterpret_cast<const uint4*>(a.e_blob + (size_t)tile * 65536) + (c.col0 / 8) * 128 + r;
      auto chunk_of = [](int i) { return NSPLIT ? (i >> 2) * 8 + (i & 3) : i; };
      int dst = -1;
      c.src = 0;
      if (SETUP1) {
        // ---- stage 0 operand in ONE L2 round trip: the hi part of my share travels by cp.async through my warp's
        // staging rows (idle until the first gather), the lo part through registers, both in flight together; the
        // A-ready arrive comes before anything else the tile needs (endpoints, gathers, prefetches) ----
        const uint32_t hrow = c.gbuf[0] + lane * 16;           // chunk-major: chunk i of lane l at 512 i + 16 l (4 KB)
        __syncwarp();                                          // the previous tile's segmented sum read these rows
#pragma unroll
        for (int i = 0; i < 8; i++) cp_async16s(hrow + i * 512, bh + chunk_of(i) * 128);
        cp_async_commit();
        auto hi_from_staging = [&](int j) {
          const float4 u0 = lds128(hrow + (2 * j) * 512);
          const float4 u1 = lds128(hrow + (2 * j + 1) * 512);
          const uint32_t h[8] = {__float_as_uint(u0.x), __float_as_uint(u0.y), __float_as_uint(u0.z), __float_as_uint(u0.w),
                                 __float_as_uint(u1.x), __float_as_uint(u1.y), __float_as_uint(u1.z), __float_as_uint(u1.w)};
          tmem_st8(c.Dc + coff<NSPLIT>(j), h);
        };
        if (c.valid) {
          c.src = __ldg(a.col + e);
          dst = __ldg(a.edst + e);
        }
        if (exact) {
          uint4 ql[8];
#pragma unroll
          for (int i = 0; i < 8; i++)
            ql[i] = ldg_na(bh + 32768 / 16 + chunk_of(i) * 128);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t l[8] = {ql[2 * j].x, ql[2 * j].y, ql[2 * j].z, ql[2 * j].w,
                                   ql[2 * j + 1].x, ql[2 * j + 1].y, ql[2 * j + 1].z, ql[2 * j + 1].w};
            tmem_st8(c.Dc + coff<NSPLIT>(j) + 8, l);
          }
        }
        cp_async_wait<0>();               // my own row: no other lane's copy is read
#pragma unroll
        for (int j = 0; j < 4; j++) hi_from_staging(j);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(a_bar);
      }
      uint4 q[8];
      if (!SETUP1) {
        // everything this tile needs from global memory first (independent loads, one L2 latency for all of them):
        // my half of the e tile and the endpoints of my edge
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = __ldg(bh + chunk_of(i) * 128);
        if (c.valid) {
          c.src = __ldg(a.col + e);
          dst = __ldg(a.edst + e);
        }
      }
      // the next tile of this slot: pull its e blob and edge endpoints into L2 while this tile is being processed
      next_tile = tile_of(sslot + NSLOT * ncl);
      if (next_tile >= 0 && next_tile != phantom && (lane & 7) == 0) {
        const uint8_t* nb = a.e_blob + (size_t)next_tile * 65536 + ((size_t)(c.col0 / 8) * 128 + r) * 16;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + chunk_of(i) * 2048));
          if (exact) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + 32768 + chunk_of(i) * 2048));
        }
        if (ch == 0) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.col + (size_t)next_tile * TILE + r));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.edst + (size_t)next_tile * TILE + r));
        }
      }
      c.dst_row = reinterpret_cast<const float4*>(a.dstA + (size_t)(dst < 0 ? 0 : dst) * 128 + c.col0);
      issue_gather<NSPLIT>(c, 0);        // (starts with a warp barrier: the staging rows are free again)
      issue_gather<NSPLIT>(c, 1);
      if (a.row_prefetch && !NSPLIT) {
        // gathers 2, 3 (src_affine) and 6, 7 (hn) are issued only one chunk before they are consumed - two staging
        // buffers per warp - and a sender row's first touch comes from DRAM: pull my sender's 256 B of both arrays
        // into L2 now, stages ahead (no registers, no staging)
        const char* ps = reinterpret_cast<const char*>(a.srcA + (size_t)c.src * 128 + c.col0);
        const char* ph = reinterpret_cast<const char*>(a.hn + (size_t)c.src * 128 + c.col0);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ps));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + 128));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ph));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ph + 128));
        if (a.row_prefetch > 1) {      // ... and my receiver's dst_affine half row (consumed in stage 1)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(c.dst_row)));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(c.dst_row) + 128));
        }
      }

      if (!SETUP1) {
        // ---- stage 0 operand: my half of the e tile -> my 64 columns of the home block, K step j at columns 16 j:
        //      [8 columns of bf16 hi pairs | 8 columns of lo pairs] (the layout every epilogue writes in place) ----
#pragma unroll
        for (int part = 0; part < 2; part++) {
          if (part == 1) {
            if (!exact) break;
#pragma unroll
            for (int i = 0; i < 8; i++) q[i] = __ldg(bh + 32768 / 16 + chunk_of(i) * 128);
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t h[8] = {q[2 * j].x, q[2 * j].y, q[2 * j].z, q[2 * j].w,
                                   q[2 * j + 1].x, q[2 * j + 1].y, q[2 * j + 1].z, q[2 * j + 1].w};
            tmem_st8(c.Dc + coff<NSPLIT>(j) + part * 8, h);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(a_bar);
      }
      if (dbg_on) dbg_rec[dbg_n++] = gtime();

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

      // the GEMM of a stage writes the floating block; the leader's MMA warp publishes which one (in both CTAs'
      // shared memory, followed by a cluster-scope fence) BEFORE it issues the GEMM whose commit releases this wait,
      // so a CTA-scope wait suffices here (a cluster-scope acquire would invalidate L1 - the dst_affine rows - four
      // times per tile)
#define GAMD_STAGE(S)                              \
  if (dbg_on) dbg_rec[dbg_n++] = gtime();        \
  if (S == 1) {   /* my receiver's dst_affine half row (two lines) into L1 underneath the wait */ \
    asm volatile("prefetch.global.L1 [%0];" ::"l"(c.dst_row)); \
    asm volatile("prefetch.global.L1 [%0];" ::"l"(c.dst_row + 8)); \
  }                                                \
  mbar_wait_hint(d_bar, d_par, a.wait_hint_ns);    \
  d_par ^= 1;                                      \
  tc_fence_after();                                \
  nstage++;                                        \
  if (STAMP) {                                     \
    uint32_t hv = sm.home[g];                      \
    while ((hv >> 8) != nstage) hv = sm.home[g];   \
    c.Dc = tb + lane_base + (hv & 0xffu) * 128u;   \
  } else {                                         \
    c.Dc = tb + lane_base + (STATIC ? ((3u + S * nact + (uint32_t)g) & 3u) : sm.home[g]) * 128u; \
  }                                                \
  c.d_par1 = d_par ^ 1u;                           \
  if (dbg_on) dbg_rec[dbg_n++] = gtime();        \
  if (exact) stage_epilogue<S, true, NSPLIT, POLY>(c);   \
  else stage_epilogue<S, false, NSPLIT>(c);        \
  if (S < 3) {                                     \
    tmem_wait_st();                                \
    tc_fence_before();                             \
    __syncwarp();                                  \
    if (lane == 0) mbar_arrive_cluster_relaxed(a_bar); \
  }                                                \
  if (dbg_on) dbg_rec[dbg_n++] = gtime();
      { GAMD_STAGE(0) }
      { GAMD_STAGE(1) }
      { GAMD_STAGE(2) }
      { GAMD_STAGE(3) }
#undef GAMD_STAGE
    }
  } else if (STATIC && warp == MMA_WARP && rank == 0) {
    // ===================== MMA issue for the CTA pair, fixed service order =====================
    const uint32_t leader = elect_leader();
    const uint32_t idesc = umma_idesc_bf16(256, 128);
    const int n_my_groups = cid < ngroups ? (ngroups - 1 - cid) / ncl + 1 : 0;
    const uint32_t wbase = smem_u32(&sm.w[0][0][0]);
    uint32_t a_par_bits = 0;                  // bit g: parity of slot g's next A-ready phase
    uint32_t home[NSLOT];
#pragma unroll
    for (int g = 0; g < NSLOT; g++) home[g] = g;
    uint32_t n = 0;                           // GEMMs issued so far: GEMM n writes block (3 + n) & 3
    long long* dbg_rec = (a.dbg && cid == 0) ? a.dbg + MMA_WARP * 256 : nullptr;
    int dbg_n = 0;
    for (int i = 0; i < n_my_groups; i++) {
      const int grp = cid + i * ncl;
      const int nact = min(NSLOT, nsuper - grp * NSLOT);
      // tail group: the blocks of the absent slots are written too - not before those slots have read their last sums
#pragma unroll
      for (int q = 1; q < NSLOT; q++) {
        if (q < nact) continue;
        uint32_t spins = 0;
        while (!__all_sync(0xffffffffu, mbar_test_wait(&sm.a_ready[q], (a_par_bits >> q) & 1u)))
          if (++spins > (1u << 26)) __trap();
        a_par_bits ^= 1u << q;
        tc_fence_after();
      }
#pragma unroll 1
      for (int s = 0; s < 4; s++) {
        const uint32_t bhi = wbase + (uint32_t)(s * 2) * WPART;
        const uint64_t dsc = umma_desc_sw128(bhi);
        const uint32_t dhi = (uint32_t)(dsc >> 32);
        const uint32_t dlo_hi = (uint32_t)dsc, dlo_lo = dlo_hi + (WPART >> 4);
#pragma unroll
        for (int g = 0; g < NSLOT; g++) {
          if (g >= nact) continue;
          uint32_t spins = 0;
          while (!__all_sync(0xffffffffu, mbar_test_wait(&sm.a_ready[g], (a_par_bits >> g) & 1u)))
            if (++spins > (1u << 26)) __trap();
          a_par_bits ^= 1u << g;
          tc_fence_after();
          const bool dbg_on = dbg_rec && lane == 0 && dbg_n + 3 <= 256;
          if (dbg_on) {
            dbg_rec[dbg_n++] = g * 4 + s;
            dbg_rec[dbg_n++] = gtime();
          }
          const uint32_t blk = (3u + n) & 3u;
          const uint32_t d = tb + blk * 128u, ab = tb + home[g] * 128u;
#pragma unroll
          for (int ks = 0; ks < 8; ks++)     // A_hi * B_hi
            umma_ts2_elect_lh(d, ab + ks * 16, dlo_hi + (ks >> 2) * (8192 >> 4) + (ks & 3) * 2, dhi, idesc, ks ? 1u : 0u, leader);
          if (a.exact) {
#pragma unroll
            for (int ks = 0; ks < 8; ks++)   // A_lo * B_hi
              umma_ts2_elect_lh(d, ab + 8 + ks * 16, dlo_hi + (ks >> 2) * (8192 >> 4) + (ks & 3) * 2, dhi, idesc, 1u, leader);
#pragma unroll
            for (int ks = 0; ks < 8; ks++)   // A_hi * B_lo
              umma_ts2_elect_lh(d, ab + ks * 16, dlo_lo + (ks >> 2) * (8192 >> 4) + (ks & 3) * 2, dhi, idesc, 1u, leader);
          }
          if (leader) umma_commit2_mc(&sm.d_ready[g][0], (uint16_t)3);
          __syncwarp();
          if (dbg_on) dbg_rec[dbg_n++] = gtime();
          home[g] = blk;
          n++;
        }
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP && rank == 0) {
    // ===================== MMA issue for the CTA pair: an event loop over the tiles in flight =====================
    // Each slot walks its own sequence of (group, stage) steps Q = 4 * group_iteration + stage and is served as soon
    // as both CTAs' A operands are in TMEM, independently of the other slots (round-robin from the slot after the one
    // served last) - the weights of every stage are resident, so nothing else couples the slots.  A GEMM reads its A
    // operand from the slot's home block and writes the floating block; afterwards the roles swap.  The next GEMM
    // therefore overwrites the block the previous one reads: SAFE_WAR waits for the previous commit first
    // (tcgen05.mma of one thread execute in issue order, so this is belt and braces).
    // The whole warp runs the loop on warp-uniform values; only the MMAs and commits are predicated on one elected
    // lane (see tc_common.cuh: issuing from inside `if (lane == 0)` halves the MMA issue rate).
    {
      const uint32_t leader = elect_leader();
      const uint32_t idesc = umma_idesc_bf16(256, NSPLIT ? 64 : 128);
      const int n_my_groups = cid < ngroups ? (ngroups - 1 - cid) / ncl + 1 : 0;
      const int totalQ = 4 * n_my_groups;
      // per-slot state lives in registers: every loop over the slots is fully unrolled with static indices
      int Qg[NSLOT];
      uint32_t a_par[NSLOT], home[NSLOT], home_peer[NSLOT];
#pragma unroll
      for (int g = 0; g < NSLOT; g++) {
        Qg[g] = 0;
        a_par[g] = 0;
        home[g] = g;
        home_peer[g] = mapa_u32(smem_u32(const_cast<uint32_t*>(&sm.home[g])), 1);
      }
      uint32_t floating = NSLOT;
      uint32_t last_bar = 0, last_par = 0;   // d_ready barrier / parity of the GEMM issued last (guards the block it read)
      int first = 0;                         // round-robin start
      uint32_t spins = 0;
      uint32_t c_par_bits = 0;               // bit g: parity of slot g's next commit on d_ready
      const uint32_t wbase = smem_u32(&sm.w[0][0][0]);
      long long* dbg_rec = (a.dbg && cid == 0) ? a.dbg + MMA_WARP * 256 : nullptr;   // (slot, pick time, commit time) per GEMM
      int dbg_n = 0;
      for (;;) {
        bool done = true;
#pragma unroll
        for (int g = 0; g < NSLOT; g++) done = done && Qg[g] >= totalQ;
        if (done) break;
        bool progressed = false;
        // the floating block is the A operand of the GEMM issued last: it may be overwritten once that GEMM completed
        bool war_ok = true;
        if (SAFE_WAR && last_bar) {
          uint32_t ok;
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
              "selp.u32 %0, 1, 0, p;\n\t}"
              : "=r"(ok)
              : "r"(last_bar), "r"(last_par)
              : "memory");
          war_ok = __all_sync(0xffffffffu, ok != 0);
        }
        int pick = -1, best = NSLOT;
#pragma unroll
        for (int g = 0; g < NSLOT; g++) {
          const int Q = Qg[g];
          if (Q >= totalQ) continue;
          const int grp = cid + (Q >> 2) * ncl;
          if (grp * NSLOT + g >= nsuper) {          // absent super slot of the tail group
            Qg[g] = totalQ;
            progressed = true;
            continue;
          }
          if (!war_ok) continue;
          if (!__all_sync(0xffffffffu, STAMP ? mbar_test_wait(&sm.a_ready[g], a_par[g])
                                                 : mbar_test_wait_cluster(&sm.a_ready[g], a_par[g]))) continue;
          int pr = g - first;
          if (pr < 0) pr += NSLOT;
          if (pr < best) {
            best = pr;
            pick = g;
          }
        }
#pragma unroll
        for (int g = 0; g < NSLOT; g++) {
          if (g != pick) continue;
          const int s = Qg[g] & 3;
          a_par[g] ^= 1;
          tc_fence_after();
          const bool dbg_on = dbg_rec && lane == 0 && dbg_n + 3 <= 256;
          if (dbg_on) {
            dbg_rec[dbg_n++] = g * 4 + s;
            dbg_rec[dbg_n++] = gtime();
          }
          const uint32_t bhi = wbase + (uint32_t)(s * 2) * WPART;      // [hi | lo] images of stage s, WPART apart
          const uint32_t d = tb + floating * 128u, ab = tb + home[g] * 128u;
          if (leader) {       // read by the slot's epilogue threads (both CTAs) after the commit arrives
            if (STAMP) {
              const uint32_t hv = ((uint32_t)(Qg[g] + 1) << 8) | floating;
              sm.home[g] = hv;
              st_cluster_u32(home_peer[g], hv);
            } else {
              sm.home[g] = floating;
              st_cluster_u32(home_peer[g], floating);
              fence_acq_rel_cluster();
            }
          }
          __syncwarp();
          // 24 (bf16x3) or 8 (bf16) MMAs per N part, fully unrolled: per MMA only the TMEM column of A and the
          // start-address field of the B descriptor change, both by compile-time constants.  N-split: two N = 64 GEMMs
          // (rows 0-31 / 32-63 of each CTA's B image, accumulator columns 0-63 / 64-127), each with its own commit
          {
            const uint64_t dsc = umma_desc_sw128(bhi);
            const uint32_t dhi = (uint32_t)(dsc >> 32);
#pragma unroll
            for (int h = 0; h < (NSPLIT ? 2 : 1); h++) {
              const uint32_t dlo_hi = (uint32_t)dsc + h * (4096 >> 4), dlo_lo = dlo_hi + (WPART >> 4);
              const uint32_t dh = d + h * 64;
#pragma unroll
              for (int ks = 0; ks < 8; ks++)     // A_hi * B_hi
                umma_ts2_elect_lh(dh, ab + ks * 16, dlo_hi + (ks >> 2) * (8192 >> 4) + (ks & 3) * 2, dhi, idesc, ks ? 1u : 0u, leader);
              if (a.exact) {
#pragma unroll
                for (int ks = 0; ks < 8; ks++)   // A_lo * B_hi
                  umma_ts2_elect_lh(dh, ab + 8 + ks * 16, dlo_hi + (ks >> 2) * (8192 >> 4) + (ks & 3) * 2, dhi, idesc, 1u, leader);
#pragma unroll
                for (int ks = 0; ks < 8; ks++)   // A_hi * B_lo
                  umma_ts2_elect_lh(dh, ab + ks * 16, dlo_lo + (ks >> 2) * (8192 >> 4) + (ks & 3) * 2, dhi, idesc, 1u, leader);
              }
              if (leader) umma_commit2_mc(&sm.d_ready[g][h], (uint16_t)3);
            }
          }
          __syncwarp();
          if (dbg_on) dbg_rec[dbg_n++] = gtime();
          last_bar = smem_u32(&sm.d_ready[g][NSPLIT ? 1 : 0]);
          last_par = (c_par_bits >> g) & 1u;
          c_par_bits ^= 1u << g;
          const uint32_t old_home = home[g];
          home[g] = floating;
          floating = old_home;
          first = g + 1 == NSLOT ? 0 : g + 1;
          Qg[g]++;
          progressed = true;
        }
        if (progressed) spins = 0;
        else if (++spins > (1u << 26)) __trap();
      }
    }
    __syncwarp();
  }
  // nobody leaves (or frees TMEM) while the peer may still arrive on my barriers / run MMAs into my TMEM
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == MMA_WARP) tmem_dealloc2(tb, 512);
}

}  // namespace

// CTA pairs, resident weights, three tiles in flight per SM (see the header of this file)
int mp_edge_tc2_launch(gamd_ctx* ctx, int layer, cudaStream_t st, int which, bool safe_war, bool nsplit) {
  const size_t smem = sizeof(SmemTC);
  if (!(ctx->attr_mask & GAMD_ATTR_MP_TC2)) {
    GAMD_CUDA(cudaFuncSetAttribute(k_mp_edge_tc2<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GAMD_CUDA(cudaFuncSetAttribute(k_mp_edge_tc2<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GAMD_CUDA(cudaFuncSetAttribute(k_mp_edge_tc2<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GAMD_CUDA(cudaFuncSetAttribute(k_mp_edge_tc2<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GAMD_CUDA(cudaFuncSetAttribute(k_mp_edge_tc2<false, false, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GAMD_CUDA(cudaFuncSetAttribute(k_mp_edge_tc2<false, false, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GAMD_CUDA(cudaFuncSetAttribute(k_mp_edge_tc2<false, false, true, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GAMD_CUDA(cudaFuncSetAttribute(k_mp_edge_tc2<false, false, true, 0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->attr_mask |= GAMD_ATTR_MP_TC2;
  }
  MpTcArgs a;
  a.w_img = ctx->d_wimg2 + (size_t)layer * 2 * WTOTAL;
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
  a.wait_hint_ns = (uint32_t)ctx->wait_hint_ns;
  a.row_prefetch = ctx->mp_row_prefetch;
  a.dbg = (ctx->dbg_timeline && layer == 1) ? reinterpret_cast<long long*>(ctx->e_emb + (size_t)ctx->cap_edges * 128) : nullptr;
  // interior launch of a tile-split layer: optionally leave a few SMs to the halo exchange running beside it
  const int reserve = ctx->dd_reserve_sms;
  int grid = which == 0 && reserve > 0 && reserve < ctx->sm_count ? ctx->sm_count - reserve : ctx->sm_count;
  grid &= ~1;       // whole CTA pairs
  // nsplit needs the N-split row order of the pair weight images (capi.cu builds them by ctx->mp_variant)
  if (ctx->mp_variant == 11) k_mp_edge_tc2<false, false, true, 0, true><<<grid, THREADS, smem, st>>>(a);
  else if (ctx->mp_variant == 12) k_mp_edge_tc2<false, false, true, 0, false, true><<<grid, THREADS, smem, st>>>(a);
  else if (ctx->mp_variant == 9) k_mp_edge_tc2<false, false, true, 1><<<grid, THREADS, smem, st>>>(a);
  else if (ctx->mp_variant == 10) k_mp_edge_tc2<false, false, true, 2><<<grid, THREADS, smem, st>>>(a);
  else if (ctx->mp_variant == 8) k_mp_edge_tc2<false, false, true><<<grid, THREADS, smem, st>>>(a);
  else if (nsplit) k_mp_edge_tc2<false, true, false><<<grid, THREADS, smem, st>>>(a);
  else if (safe_war) k_mp_edge_tc2<true, false, false><<<grid, THREADS, smem, st>>>(a);
  else k_mp_edge_tc2<false, false, false><<<grid, THREADS, smem, st>>>(a);
  GAMD_LAUNCH_CHECK();
  return 0;
}
