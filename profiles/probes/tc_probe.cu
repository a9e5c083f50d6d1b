// Stand-alone probe of the tcgen05 building blocks in tc_common.cuh (run on the B200 box):
//   T1  D = A * B^T, A and B from shared memory (SWIZZLE_128B K-major), 128x128x128 bf16
//   T2  same with A written to TMEM by tcgen05.st (packed bf16 pairs) - the "TS" form
//   T3  3-pass split-bf16 (hi*hi + lo*hi + hi*lo) of fp32 operands vs an fp64 reference
// Prints max abs errors; exit code 0 iff all pass.  Development tool, not part of the library.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../gamd_b200/csrc/tc_common.cuh"

using namespace tc;

constexpr int M = 128, N = 128, K = 128;

struct __align__(1024) ProbeSmem {
  uint8_t a_hi[M * K * 2];
  uint8_t a_lo[M * K * 2];
  uint8_t b_hi[N * K * 2];
  uint8_t b_lo[N * K * 2];
  uint64_t bar;
  uint32_t tmem_base;
};

// mode 0: SS bf16; mode 1: TS bf16; mode 2: TS split x3 (fp32 inputs); d_col: accumulator column offset
__global__ void __launch_bounds__(128) k_probe(const float* __restrict__ A, const float* __restrict__ B,
                                               float* __restrict__ D, int mode, int d_col) {
  extern __shared__ __align__(1024) uint8_t raw[];
  ProbeSmem& sm = *reinterpret_cast<ProbeSmem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
  if (tid == 0) {
    mbar_init(&sm.bar, 1);
    fence_barrier_init();
  }
  // operands -> shared memory in the UMMA canonical layout (generic-proxy stores)
  for (int idx = tid; idx < M * K / 2; idx += 128) {
    int r = idx / (K / 2), k = (idx % (K / 2)) * 2;
    uint32_t hi, lo;
    split_bf16(A[r * K + k], A[r * K + k + 1], hi, lo);
    *reinterpret_cast<uint32_t*>(sm.a_hi + sw128_offset(r, k, M)) = hi;
    *reinterpret_cast<uint32_t*>(sm.a_lo + sw128_offset(r, k, M)) = lo;
    split_bf16(B[r * K + k], B[r * K + k + 1], hi, lo);
    *reinterpret_cast<uint32_t*>(sm.b_hi + sw128_offset(r, k, N)) = hi;
    *reinterpret_cast<uint32_t*>(sm.b_lo + sw128_offset(r, k, N)) = lo;
  }
  fence_proxy_async();        // make the generic-proxy smem writes visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sm.tmem_base;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t a_hi_col = 128 + ((d_col + 128) & 256 ? 0 : 0);   // A operand columns [384,448) hi, [448,512) lo
  const uint32_t AH = 384, AL = 448;
  (void)a_hi_col;
  if (mode >= 1) {
    // thread = row: write my row of A into TMEM as packed bf16 pairs (column j holds k = 2j, 2j+1)
    const int r = tid;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      uint32_t h[16], l[16];
#pragma unroll
      for (int j = 0; j < 16; j++) {
        int k = (c * 16 + j) * 2;
        split_bf16(A[r * K + k], A[r * K + k + 1], h[j], l[j]);
      }
      tmem_st16(tbase + lane_base + AH + c * 16, h);
      tmem_st16(tbase + lane_base + AL + c * 16, l);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(M, N);
    const uint32_t d = tbase + d_col;
    const int passes = mode == 2 ? 3 : 1;
    uint32_t accum = 0;
    for (int p = 0; p < passes; p++) {
      const uint8_t* bsrc = (p == 2) ? sm.b_lo : sm.b_hi;       // hi*hi, lo*hi, hi*lo
      const uint8_t* asrc = (p == 1) ? sm.a_lo : sm.a_hi;
      const uint32_t acol = (p == 1) ? AL : AH;
      for (int ks = 0; ks < K / 16; ks++) {
        uint32_t boff = (ks >> 2) * (N * 128) + (ks & 3) * 32;
        uint64_t bdesc = umma_desc_sw128(smem_u32(bsrc) + boff);
        if (mode == 0) {
          uint64_t adesc = umma_desc_sw128(smem_u32(asrc) + (ks >> 2) * (M * 128) + (ks & 3) * 32);
          umma_ss(d, adesc, bdesc, idesc, accum);
        } else {
          umma_ts(d, tbase + acol + ks * 8, bdesc, idesc, accum);
        }
        accum = 1;
      }
    }
    umma_commit(&sm.bar);
  }
  mbar_wait(&sm.bar, 0);
  tc_fence_after();
  {
    const int r = tid;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      uint32_t v[16];
      tmem_ld16(tbase + lane_base + d_col + c * 16, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; j++) D[r * N + c * 16 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

static float bf16_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
  r &= 0xFFFF0000u;
  float y;
  memcpy(&y, &r, 4);
  return y;
}

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : B) x = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.1f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  size_t smem = sizeof(ProbeSmem) + 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int fails = 0;
  struct Case { int mode, d_col; const char* name; double tol; };
  Case cases[] = {{0, 0, "T1 SS bf16      ", 1e-5}, {1, 0, "T2 TS bf16      ", 1e-5}, {1, 256, "T2b TS bf16 d@256", 1e-5},
                  {2, 0, "T3 TS bf16x3    ", 3e-5}};
  for (auto& c : cases) {
    cudaMemset(dD, 0, D.size() * 4);
    k_probe<<<1, 128, smem>>>(dA, dB, dD, c.mode, c.d_col);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e));
      return 2;
    }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; m++)
      for (int n = 0; n < N; n++) {
        double ref = 0;
        for (int k = 0; k < K; k++) {
          double a = c.mode == 2 ? A[m * K + k] : bf16_round(A[m * K + k]);
          double b = c.mode == 2 ? B[n * K + k] : bf16_round(B[n * K + k]);
          ref += a * b;
        }
        maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
        maxref = fmax(maxref, fabs(ref));
      }
    bool ok = maxerr / maxref < c.tol;
    printf("%s max|err| %.3e  max|ref| %.3e  rel %.3e  %s\n", c.name, maxerr, maxref, maxerr / maxref, ok ? "PASS" : "FAIL");
    fails += !ok;
  }
  return fails ? 1 : 0;
}
