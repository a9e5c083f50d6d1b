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
// CTA = 10 warps: 2 epilogue warpgroups (thread = edge row; each owns one 128-edge tile in flight, the two
// tiles ping-pong on the tensor core and share every weight load), 1 MMA-issue warp, 1 weight-producer
// warp.  Neighbour features are gathered with coalesced cp.async into per-warp staging rows; the final
// segmented sum is a warp-shuffle segmented scan over the receiver-sorted rows - no atomics; rows
// that straddle a 32-edge block go to the `part` side buffer and are summed (in order) by the node kernel.
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int TILE = 128;
constexpr int WCHUNK = 32768;        // one weight part image (128 x 128 bf16)
constexpr int RING = 4;
constexpr int GROW = 144;            // staging row stride in bytes (128 data + 16 pad: conflict-free LDS.128)
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 320;

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
  int exact;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
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

__global__ void __launch_bounds__(THREADS, 1) k_mp_edge_tc(MpTcArgs a) {
  extern __shared__ __align__(1024) uint8_t raw[];
  SmemTC& sm = *reinterpret_cast<SmemTC*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int E = *a.n_edges;
  const int ntiles = (E + TILE - 1) / TILE;
  const int npairs = (ntiles + 1) / 2;

  if (warp == 8) tmem_alloc(&sm.tmem_base, 512);
  if (tid == 0) {
    for (int i = 0; i < RING; i++) {
      mbar_init(&sm.full[i], 1);
      mbar_init(&sm.empty[i], 1);
    }
    for (int g = 0; g < 2; g++) {
      mbar_init(&sm.a_ready[g], 128);
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
    // ===================== epilogue warpgroups: thread = edge row =====================
    const int g = warp >> 2, wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t Dc = tb + lane_base + g * 256, AH = Dc + 128, AL = Dc + 192;
    uint8_t* gbuf0 = sm.gather[warp][0];
    uint8_t* gbuf1 = sm.gather[warp][1];
    uint32_t d_par = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int tile = pair * 2 + g;
      if (tile >= ntiles) continue;
      const int e0 = tile * TILE;
      const int e = e0 + r;
      const bool valid = e < E;
      const int src = valid ? a.col[e] : 0;
      const int dst = valid ? a.edst[e] : -1;
      const int dstc = dst < 0 ? 0 : dst;

      auto issue_gather = [&](const float* base, int c, uint8_t* buf) {
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; it++) {
          int j = it * 4 + (lane >> 3);
          int sj = __shfl_sync(0xffffffffu, src, j);
          cp_async16(buf + j * GROW + (lane & 7) * 16, base + (size_t)sj * 128 + c * 32 + (lane & 7) * 4);
        }
        cp_async_commit();
      };
      issue_gather(a.srcA, 0, gbuf0);   // gather #0 (stage-1 chunk 0)

      // ---- stage 0 operand: e tile (bf16 hi / lo) -> TMEM A -----------------------------
      {
        const uint4* bh = reinterpret_cast<const uint4*>(a.e_blob + (size_t)tile * 65536);
        const uint4* bl = reinterpret_cast<const uint4*>(a.e_blob + (size_t)tile * 65536 + 32768);
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
          uint32_t h[16];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            uint4 q = __ldg(bh + (c4 * 4 + i) * 128 + r);
            h[4 * i] = q.x; h[4 * i + 1] = q.y; h[4 * i + 2] = q.z; h[4 * i + 3] = q.w;
          }
          tmem_st16(AH + c4 * 16, h);
          if (a.exact) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
              uint4 q = __ldg(bl + (c4 * 4 + i) * 128 + r);
              h[4 * i] = q.x; h[4 * i + 1] = q.y; h[4 * i + 2] = q.z; h[4 * i + 3] = q.w;
            }
            tmem_st16(AL + c4 * 16, h);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&sm.a_ready[g]);
      }

      // segment structure of my warp's 32 rows (receiver-sorted)
      uint32_t same = 0;   // bit k: row (lane - 2^k) belongs to my segment
#pragma unroll
      for (int k = 0; k < 5; k++) {
        int od = __shfl_up_sync(0xffffffffu, dst, 1 << k);
        if (lane >= (1 << k) && od == dst) same |= 1u << k;
      }
      const int nd = __shfl_down_sync(0xffffffffu, dst, 1);
      const bool seg_end = valid && (lane == 31 || nd != dst);
      float* out_row = nullptr;
      if (seg_end) {
        const int bstart = e0 + wq * 32;
        const int bend = min(bstart + 32, E);
        const int blk = bstart >> 5;
        if (a.row_ptr[dst] < bstart) out_row = a.part + ((size_t)blk * 2 + 0) * 128;
        else if (a.row_ptr[dst + 1] > bend) out_row = a.part + ((size_t)blk * 2 + 1) * 128;
        else out_row = a.agg + (size_t)dst * 128;
      }

#pragma unroll 1
      for (int s = 0; s < 4; s++) {
        mbar_wait(&sm.d_ready[g], d_par);
        d_par ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; c++) {
          uint32_t v0[16], v1[16];
          tmem_ld16(Dc + c * 32, v0);
          tmem_ld16(Dc + c * 32 + 16, v1);
          const uint8_t* grow = nullptr;
          if (s == 1 || s == 3) {
            // software pipeline over the 8 gathers of this tile: srcA chunks 0-3, then hn chunks 0-3
            const int gi = (s == 1 ? 0 : 4) + c;
            uint8_t* cur = (gi & 1) ? gbuf1 : gbuf0;
            uint8_t* nxt = (gi & 1) ? gbuf0 : gbuf1;
            if (gi + 1 < 8) {
              issue_gather(gi + 1 < 4 ? a.srcA : a.hn, (gi + 1) & 3, nxt);
              cp_async_wait<1>();
            } else {
              cp_async_wait<0>();
            }
            __syncwarp();
            grow = cur + lane * GROW;
          }
          tmem_wait_ld();
          float x[32];
#pragma unroll
          for (int j = 0; j < 16; j++) {
            x[j] = __uint_as_float(v0[j]);
            x[16 + j] = __uint_as_float(v1[j]);
          }
#pragma unroll
          for (int j4 = 0; j4 < 8; j4++) {
            float4 b = *reinterpret_cast<const float4*>(&sm.bias[s][c * 32 + j4 * 4]);
            x[4 * j4] += b.x; x[4 * j4 + 1] += b.y; x[4 * j4 + 2] += b.z; x[4 * j4 + 3] += b.w;
          }
          if (s == 1) {
            const float4* dr = reinterpret_cast<const float4*>(a.dstA + (size_t)dstc * 128 + c * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; j4++) {
              float4 sv = *reinterpret_cast<const float4*>(grow + j4 * 16);
              float4 dv = __ldg(dr + j4);
              x[4 * j4] += sv.x + dv.x; x[4 * j4 + 1] += sv.y + dv.y;
              x[4 * j4 + 2] += sv.z + dv.z; x[4 * j4 + 3] += sv.w + dv.w;
            }
          }
          if (s < 3) {
            uint32_t h[16], l[16];
#pragma unroll
            for (int j = 0; j < 16; j++) split_bf16(silu_fast(x[2 * j]), silu_fast(x[2 * j + 1]), h[j], l[j]);
            tmem_st16(AH + c * 16, h);
            if (a.exact) tmem_st16(AL + c * 16, l);
          } else {
            // message = hn[src] * e_emb, then segmented inclusive scan down the receiver-sorted rows
#pragma unroll
            for (int j4 = 0; j4 < 8; j4++) {
              float4 hv = *reinterpret_cast<const float4*>(grow + j4 * 16);
              x[4 * j4] = valid ? x[4 * j4] * hv.x : 0.f;
              x[4 * j4 + 1] = valid ? x[4 * j4 + 1] * hv.y : 0.f;
              x[4 * j4 + 2] = valid ? x[4 * j4 + 2] * hv.z : 0.f;
              x[4 * j4 + 3] = valid ? x[4 * j4 + 3] * hv.w : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 5; k++) {
              const bool take = (same >> k) & 1u;
#pragma unroll
              for (int j = 0; j < 32; j++) {
                float t = __shfl_up_sync(0xffffffffu, x[j], 1 << k);
                if (take) x[j] += t;
              }
            }
            if (out_row) {
#pragma unroll
              for (int j4 = 0; j4 < 8; j4++)
                *reinterpret_cast<float4*>(out_row + c * 32 + j4 * 4) =
                    make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]);
            }
          }
        }
        if (s < 3) {
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&sm.a_ready[g]);
        }
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issue (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      uint32_t a_par[2] = {0, 0};
      uint32_t q = 0;
      bool first = true;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        for (int s = 0; s < 4; s++) {
          uint32_t slot_hi, slot_lo = 0;
          if (a.exact) {
            slot_hi = q % RING;
            slot_lo = (q + 1) % RING;
            mbar_wait(&sm.full[slot_hi], (q / RING) & 1);
            mbar_wait(&sm.full[slot_lo], ((q + 1) / RING) & 1);
          } else {
            slot_hi = s;
            if (first) mbar_wait(&sm.full[s], 0);
          }
          const uint32_t bhi = smem_u32(sm.w[slot_hi]), blo = smem_u32(sm.w[slot_lo]);
          for (int g = 0; g < 2; g++) {
            if (pair * 2 + g >= ntiles) continue;
            mbar_wait(&sm.a_ready[g], a_par[g]);
            a_par[g] ^= 1;
            tc_fence_after();
            const uint32_t d = tb + g * 256, ah = d + 128, al = d + 192;
            const int passes = a.exact ? 3 : 1;
            uint32_t accum = 0;
            for (int p = 0; p < passes; p++) {
              const uint32_t bb = (p == 2) ? blo : bhi;
              const uint32_t aa = (p == 1) ? al : ah;
#pragma unroll
              for (int ks = 0; ks < 8; ks++) {
                umma_ts(d, aa + ks * 8, umma_desc_sw128(bb + (ks >> 2) * 16384 + (ks & 3) * 32), idesc, accum);
                accum = 1;
              }
            }
            umma_commit(&sm.d_ready[g]);
          }
          if (a.exact) {
            umma_commit(&sm.empty[slot_hi]);
            umma_commit(&sm.empty[slot_lo]);
            q += 2;
          }
        }
        first = false;
      }
    }
    __syncwarp();
  } else {
    // ===================== weight producer (one thread) =====================
    if (lane == 0) {
      auto load = [&](int slot, int stage, int part) {
        mbar_arrive_expect_tx(&sm.full[slot], WCHUNK);
        const uint8_t* src = a.w_img + ((size_t)stage * 2 + part) * WCHUNK;
#pragma unroll
        for (int i = 0; i < 4; i++) bulk_g2s(sm.w[slot] + i * 8192, src + i * 8192, 8192, &sm.full[slot]);
      };
      if (!a.exact) {
        if (blockIdx.x < npairs)
          for (int s = 0; s < 4; s++) load(s, s, 0);
      } else {
        uint32_t q = 0;
        for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x)
          for (int s = 0; s < 4; s++)
            for (int part = 0; part < 2; part++) {
              const int slot = q % RING;
              if (q >= RING) mbar_wait(&sm.empty[slot], ((q / RING) - 1) & 1);
              load(slot, s, part);
              q++;
            }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, 512);
}

}  // namespace

int mp_edge_tc_launch(gamd_ctx* ctx, int layer, cudaStream_t st) {
  static bool attr_done = false;
  const size_t smem = sizeof(SmemTC) + 1024;
  if (!attr_done) {
    GAMD_CUDA(cudaFuncSetAttribute(k_mp_edge_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
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
  a.exact = ctx->desc.precision == GAMD_PREC_BF16X3 ? 1 : 0;
  k_mp_edge_tc<<<ctx->sm_count, THREADS, smem, st>>>(a);
  GAMD_LAUNCH_CHECK();
  return 0;
}
