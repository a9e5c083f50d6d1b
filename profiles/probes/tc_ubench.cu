// Stand-alone micro-benchmark of how tcgen05.mma, tcgen05.ld / st and the MUFU pipe share one SM (run on the B200
// box).  One CTA: warps 0-7 are "epilogue" warps (two per TMEM lane quadrant) looping over TMEM loads / stores /
// MUFU work, warp 8 issues a train of 128x128x16 bf16 MMAs (A from TMEM or from shared memory).  Every role
// reports its own clock64 span, so the cost of running them together can be compared with running them alone.
// Development tool behind the numbers in DESIGN.md ("what shares what"); not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_ubench profiles/probes/tc_ubench.cu   (from the repository root)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../gamd_b200/csrc/tc_common.cuh"

using namespace tc;

struct __align__(1024) USmem {
  uint8_t b[2][32768];   // two 128x128 bf16 SW128 K-major images (contents irrelevant)
  uint8_t a[32768];
  uint64_t bar;
  uint64_t bar2[2];
  uint32_t tmem_base;
};

struct UArgs {
  int n_mma;      // MMAs issued back to back by warp 8 (0 = none)
  int ts;         // 1: A operand from TMEM, 0: from shared memory
  int n;          // MMA N (64 / 128 / 256)
  int dalt;       // accumulators used round-robin (1 = always the same)
  int two_issuers;
  int const_tb;   // 1: TMEM addresses are compile-time constants (base 0), 0: derived from the allocated base
  int ld_warps;   // how many of warps 0-7 run the loop
  int n_ld;       // tcgen05.ld x16 per loop warp
  int n_st;       // tcgen05.st x16 per loop warp
  int n_mufu;     // ex2+rcp pairs x16 per loop warp
  long long* out; // [9][2] start / end clocks
};

template <bool CONST_TB, int STYLE>
__global__ void __launch_bounds__(320, 1) k_ubench(UArgs a) {
  extern __shared__ __align__(1024) uint8_t raw[];
  USmem& sm = *reinterpret_cast<USmem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (int)sizeof(sm.b) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm.b)[i] = 0;
  for (int i = tid; i < (int)sizeof(sm.a) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm.a)[i] = 0;
  if (warp == 8) tmem_alloc(&sm.tmem_base, 512);
  if (tid == 0) {
    mbar_init(&sm.bar, 1);
    mbar_init(&sm.bar2[0], 1);
    mbar_init(&sm.bar2[1], 1);
    fence_barrier_init();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (sm.tmem_base != 0) __trap();   // the only CTA on the SM owns all 512 columns: its base is 0
  const uint32_t tb = CONST_TB ? 0u : sm.tmem_base;
  long long t0 = 0, t1 = 0;
  if (warp < 8) {
    const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;   // columns 0..127
    uint32_t v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = lane + j;
    tmem_st16(base, v);
    tmem_wait_st();
    __syncwarp();
    t0 = clock64();
    if (warp < a.ld_warps) {
      uint32_t accum = 0;
      for (int i = 0; i < a.n_ld; i += 4) {   // four loads in flight per wait (throughput, not latency)
        uint32_t w0[16], w1[16], w2[16];
        tmem_ld16(base, v);
        tmem_ld16(base + 16, w0);
        tmem_ld16(base + 32, w1);
        tmem_ld16(base + 48, w2);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; j++) accum += v[j] + w0[j] + w1[j] + w2[j];
      }
      for (int i = 0; i < a.n_st; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = accum + j;
        tmem_st16(base + (i & 3) * 16, v);
      }
      if (a.n_st) tmem_wait_st();
      float f = __uint_as_float((accum & 0xffff) | 0x3f800000u);
      float fs[16];
#pragma unroll
      for (int j = 0; j < 16; j++) fs[j] = f + j;
      for (int i = 0; i < a.n_mufu; i++) {   // 16 independent ex2 + rcp chains
#pragma unroll
        for (int j = 0; j < 16; j++) {
          float e, r;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fs[j]));
          asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
          fs[j] = fs[j] * r;
        }
      }
#pragma unroll
      for (int j = 0; j < 16; j++) f += fs[j];
      if (f == 123.f || accum == 0xdeadbeef) a.out[30] = 1;
    }
    t1 = clock64();
  } else if (STYLE == 1 && warp == 8) {
    const uint32_t leader = elect_leader();
    const uint32_t idesc = umma_idesc_bf16(128, a.n);
    t0 = clock64();
    const uint64_t bd0 = umma_desc_sw128(smem_u32(sm.b[0]));
    const uint32_t d0 = tb + 256;
    for (int i = 0; i < a.n_mma; i += 8) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const uint64_t koff = (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4);
        umma_ts_elect(d0, tb + 128 + k * 8, bd0 + koff, idesc, (i > 0 || k > 0) ? 1u : 0u, leader);
      }
    }
    if (leader) {
      umma_commit(&sm.bar2[0]);
      mbar_wait(&sm.bar2[0], 0);
    }
    __syncwarp();
    t1 = clock64();
  } else if ((warp == 8 || (warp == 9 && a.two_issuers)) && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, a.n);
    t0 = clock64();
    if (a.n_mma) {
      // descriptors are loop invariant up to small immediates: the issue loop must not be the limiter
      const uint64_t bd0 = umma_desc_sw128(smem_u32(sm.b[0]));
      const uint64_t ad0 = umma_desc_sw128(smem_u32(sm.a));
      const uint32_t d0 = tb + 256 + (warp - 8) * 128, d1 = a.dalt > 1 ? tb + 384 : d0;
      const uint32_t acc1 = a.dalt > 1 ? 0u : 1u;
      for (int i = 0; i < a.n_mma; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const uint64_t koff = (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4);
          const uint32_t d = (k & 1) ? d1 : d0;
          const uint32_t acc = (i > 0) ? 1u : (k == 0 ? 0u : (k == 1 ? acc1 : 1u));
          if (a.ts) umma_ts(d, tb + 128 + k * 8, bd0 + koff, idesc, acc);
          else umma_ss(d, ad0 + koff, bd0 + koff, idesc, acc);
        }
      }
      umma_commit(&sm.bar2[warp - 8]);
      mbar_wait(&sm.bar2[warp - 8], 0);
    }
    t1 = clock64();
  }
  if (lane == 0) {
    a.out[warp * 2] = t0;
    a.out[warp * 2 + 1] = t1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tb, 512);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64 * sizeof(long long));
  const size_t smem = sizeof(USmem) + 1024;
  cudaFuncSetAttribute(k_ubench<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_ubench<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_ubench<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_ubench<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Case { const char* name; int n_mma, ts, ld_warps, n_ld, n_st, n_mufu, n = 128, dalt = 1, const_tb = 0, two = 0, style = 0; };
  const Case cases[] = {
      {"mma TS alone (96)", 96, 1, 0, 0, 0, 0},
      {"mma SS alone (96)", 96, 0, 0, 0, 0, 0},
      {"ld x16 alone, 8 warps x 64", 0, 0, 8, 64, 0, 0},
      {"ld x16 alone, 4 warps x 64", 0, 0, 4, 64, 0, 0},
      {"ld x16 alone, 1 warp x 64", 0, 0, 1, 64, 0, 0},
      {"st x16 alone, 8 warps x 64", 0, 0, 8, 0, 64, 0},
      {"mufu alone, 8 warps x 64x16 pairs", 0, 0, 8, 0, 0, 64},
      {"mufu alone, 4 warps x 64x16 pairs", 0, 0, 4, 0, 0, 64},
      {"mma TS (96) + ld 8 warps x 64", 96, 1, 8, 64, 0, 0},
      {"mma SS (96) + ld 8 warps x 64", 96, 0, 8, 64, 0, 0},
      {"mma TS (96) + st 8 warps x 64", 96, 1, 8, 0, 64, 0},
      {"mma TS (96) + mufu 8 warps x 64", 96, 1, 8, 0, 0, 64},
      {"mma TS (96) + ld+st+mufu 8 warps x 32", 96, 1, 8, 32, 32, 32},
      {"ld+st+mufu alone 8 warps x 32", 0, 0, 8, 32, 32, 32},
      {"mma TS N=256 (96)", 96, 1, 0, 0, 0, 0, 256, 1},
      {"mma SS N=256 (96)", 96, 0, 0, 0, 0, 0, 256, 1},
      {"mma TS N=64 (96)", 96, 1, 0, 0, 0, 0, 64, 1},
      {"mma TS N=128, 2 accumulators (96)", 96, 1, 0, 0, 0, 0, 128, 2},
      {"mma SS N=128, 2 accumulators (96)", 96, 0, 0, 0, 0, 0, 128, 2},
      {"mma TS N=32 (96)", 96, 1, 0, 0, 0, 0, 32, 1},
      {"mma TS N=128 const addresses (96)", 96, 1, 0, 0, 0, 0, 128, 1, 1},
      {"mma SS N=128 const addresses (96)", 96, 0, 0, 0, 0, 0, 128, 1, 1},
      {"mma TS N=256 const addresses (96)", 96, 1, 0, 0, 0, 0, 256, 1, 1},
      {"mma TS N=64 const addresses (96)", 96, 1, 0, 0, 0, 0, 64, 1, 1},
      {"mma TS const + ld+st+mufu 8 warps x 32", 96, 1, 8, 32, 32, 32, 128, 1, 1},
      {"mma TS N=128, two issuer warps (96 each)", 96, 1, 0, 0, 0, 0, 128, 1, 0, 1},
      {"mma TS N=128 const, two issuer warps (96 each)", 96, 1, 0, 0, 0, 0, 128, 1, 1, 1},
      {"mma TS N=64, two issuer warps (96 each)", 96, 1, 0, 0, 0, 0, 64, 1, 0, 1},
      {"mma TS N=128 warp-uniform elect issue", 96, 1, 0, 0, 0, 0, 128, 1, 0, 0, 1},
      {"FAST mma (192) alone", 192, 1, 0, 0, 0, 0, 128, 1, 0, 0, 1},
      {"FAST mma (192) + ld 8 warps x 128", 192, 1, 8, 128, 0, 0, 128, 1, 0, 0, 1},
      {"FAST mma (192) + st 8 warps x 128", 192, 1, 8, 0, 128, 0, 128, 1, 0, 0, 1},
      {"FAST mma (192) + mufu 8 warps x 32", 192, 1, 8, 0, 0, 32, 128, 1, 0, 0, 1},
      {"FAST mma (192) + ld+st+mufu 8 warps 64/64/16", 192, 1, 8, 64, 64, 16, 128, 1, 0, 0, 1},
      {"ld 8 warps x 128 alone", 0, 1, 8, 128, 0, 0, 128, 1, 0, 0, 1},
      {"st 8 warps x 128 alone", 0, 1, 8, 0, 128, 0, 128, 1, 0, 0, 1},
      {"mufu 8 warps x 32 alone", 0, 1, 8, 0, 0, 32, 128, 1, 0, 0, 1},
      {"ld+st+mufu 8 warps 64/64/16 alone", 0, 1, 8, 64, 64, 16, 128, 1, 0, 0, 1},
      {"mma TS N=128 warp-uniform elect issue, const", 96, 1, 0, 0, 0, 0, 128, 1, 1, 0, 1},
      {"mma TS N=64 warp-uniform elect issue, const", 96, 1, 0, 0, 0, 0, 64, 1, 1, 0, 1},
  };
  for (const Case& c : cases) {
    UArgs a{c.n_mma, c.ts, c.n, c.dalt, c.two, c.const_tb, c.ld_warps, c.n_ld, c.n_st, c.n_mufu, d_out};
    long long h[64];
    for (int rep = 0; rep < 2; rep++) {   // second run is the warm one
      cudaMemset(d_out, 0, 64 * sizeof(long long));
      if (c.style == 1 && c.const_tb) k_ubench<true, 1><<<1, 320, smem>>>(a);
      else if (c.style == 1) k_ubench<false, 1><<<1, 320, smem>>>(a);
      else if (c.const_tb) k_ubench<true, 0><<<1, 320, smem>>>(a);
      else k_ubench<false, 0><<<1, 320, smem>>>(a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    }
    long long ld_max = 0;
    for (int w = 0; w < c.ld_warps; w++) ld_max = std::max(ld_max, h[2 * w + 1] - h[2 * w]);
    const long long mma = std::max(h[17] - h[16], h[19] - h[18]);
    printf("%-42s  mma span %7lld (%.1f cyc/MMA)   loop-warp span (max) %7lld\n", c.name, mma,
           c.n_mma ? (double)mma / c.n_mma : 0.0, ld_max);
  }
  return 0;
}
