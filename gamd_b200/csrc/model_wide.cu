// Generic-width fp32 path of the MDNet forward: every model shape the 128-wide kernels are not built for.
//
//   * WaterMDDynamicBoxNet at encoding / hidden / edge-embedding widths other than 128: the 256 / 128 / 256 x 5 model of
//     code/water/test_script/test_nosehoover_hb.py:69-81 (train_network_real_large.py:355-359) and wider ones
//     (any multiple of 128 up to 1024),
//   * update_edge: a layer replaces the edge embedding by edge_layer_norm(e_emb) for the layers after it
//     (code/nn_module.py:89-90, :139-146),
//   * expand_edge = False: 4 (+ bond flag) edge inputs, no RBF expansion (nn_module.py:312-313, :330-335),
//   * BatchNorm1d (eval mode, running statistics) instead of LayerNorm on the node features (nn_module.py:193-196).
//
// Same stage structure as model_fp32.cu (edge encoder / per-layer edge chain + segmented sum / node update), CUDA-core
// FFMA in fp32, one CTA per tile of TM = 16 R rows.  A tile's activations ping-pong between two shared-memory buffers of
// row stride XS = max width + 4; every Linear runs as N / 128 column blocks, each a [TM x K] x [K x 128] register-tiled
// product with the transposed weights streamed through a double-buffered cp.async ring.  R (4, 2 or 1 rows per thread)
// is chosen from the widest layer so that both buffers fit the 227 KB of shared memory.
#include "common.cuh"

namespace {

constexpr int NT = 256;   // 16 column groups x 16 row groups
constexpr int KC = 32;    // weight rows per ring slot
constexpr int NB = 128;   // output columns per block

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float gelu_f(float x) { return x * 0.5f * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

struct Sm {
  float *A, *B, *W;
  int *src, *dst;
  float* dh;
};

template <int R>
__device__ __forceinline__ Sm carve_smem(unsigned char* raw, int XS) {
  constexpr int TM = 16 * R;
  Sm s;
  s.A = reinterpret_cast<float*>(raw);
  s.B = s.A + TM * XS;
  s.W = s.B + TM * XS;
  s.src = reinterpret_cast<int*>(s.W + 2 * KC * NB);
  s.dst = s.src + TM;
  s.dh = reinterpret_cast<float*>(s.dst + TM);
  return s;
}

// acc[r][c]: rows ty*R + r; columns col0 + tx*4 + c (c < 4) and col0 + 64 + tx*4 + (c - 4) of the block nb (col0 = 128 nb)
template <int R>
__device__ __forceinline__ void gemm_block(float (&acc)[R][8], const float* Xs, int XS, int K, const float* __restrict__ Wt,
                                           int N, int nb, float* Wsm, int tid) {
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int r = 0; r < R; r++)
#pragma unroll
    for (int c = 0; c < 8; c++) acc[r][c] = 0.f;
  const int nch = K / KC;
  const float* wsrc = Wt + (size_t)nb * NB;
  auto load = [&](int buf, int ch) {
#pragma unroll
    for (int i = 0; i < (KC * NB / 4) / NT; i++) {
      const int idx = tid + i * NT, row = idx >> 5, c4 = idx & 31;
      cp_async16(Wsm + buf * KC * NB + row * NB + c4 * 4, wsrc + (size_t)(ch * KC + row) * N + c4 * 4);
    }
  };
  load(0, 0);
  cp_async_commit();
#pragma unroll 1
  for (int ch = 0; ch < nch; ch++) {
    if (ch + 1 < nch) {
      load((ch + 1) & 1, ch + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* Wb = Wsm + (ch & 1) * KC * NB;
    const float* Xr = Xs + (ty * R) * XS + ch * KC;
#pragma unroll
    for (int k4 = 0; k4 < KC / 4; k4++) {
      float4 a[R];
#pragma unroll
      for (int r = 0; r < R; r++) a[r] = *reinterpret_cast<const float4*>(Xr + r * XS + k4 * 4);
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const float4 w0 = *reinterpret_cast<const float4*>(Wb + (k4 * 4 + kk) * NB + tx * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(Wb + (k4 * 4 + kk) * NB + 64 + tx * 4);
#pragma unroll
        for (int r = 0; r < R; r++) {
          const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
          acc[r][0] = fmaf(av, w0.x, acc[r][0]);
          acc[r][1] = fmaf(av, w0.y, acc[r][1]);
          acc[r][2] = fmaf(av, w0.z, acc[r][2]);
          acc[r][3] = fmaf(av, w0.w, acc[r][3]);
          acc[r][4] = fmaf(av, w1.x, acc[r][4]);
          acc[r][5] = fmaf(av, w1.y, acc[r][5]);
          acc[r][6] = fmaf(av, w1.z, acc[r][6]);
          acc[r][7] = fmaf(av, w1.w, acc[r][7]);
        }
      }
    }
    __syncthreads();
  }
}

template <int R>
__device__ __forceinline__ void add_vec(float (&acc)[R][8], const float* __restrict__ b, int tx) {
  const float4 b0 = *reinterpret_cast<const float4*>(b + tx * 4);
  const float4 b1 = *reinterpret_cast<const float4*>(b + 64 + tx * 4);
#pragma unroll
  for (int r = 0; r < R; r++) {
    acc[r][0] += b0.x; acc[r][1] += b0.y; acc[r][2] += b0.z; acc[r][3] += b0.w;
    acc[r][4] += b1.x; acc[r][5] += b1.y; acc[r][6] += b1.z; acc[r][7] += b1.w;
  }
}

// acc[r][:] += M[rows[r]][col0 ...] of a row-major matrix with leading dimension ld; a negative row is skipped
template <int R>
__device__ __forceinline__ void add_rows(float (&acc)[R][8], const float* __restrict__ M, int ld, const int (&rows)[R],
                                         int col0, int tx) {
#pragma unroll
  for (int r = 0; r < R; r++) {
    if (rows[r] < 0) continue;
    const float* p = M + (size_t)rows[r] * ld + col0;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(p + tx * 4));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + 64 + tx * 4));
    acc[r][0] += v0.x; acc[r][1] += v0.y; acc[r][2] += v0.z; acc[r][3] += v0.w;
    acc[r][4] += v1.x; acc[r][5] += v1.y; acc[r][6] += v1.z; acc[r][7] += v1.w;
  }
}

template <int R>
__device__ __forceinline__ void store_smem(const float (&acc)[R][8], float* X, int XS, int col0, int tx, int ty) {
#pragma unroll
  for (int r = 0; r < R; r++) {
    float* p = X + (ty * R + r) * XS + col0;
    *reinterpret_cast<float4*>(p + tx * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    *reinterpret_cast<float4*>(p + 64 + tx * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
  }
}

template <int R>
__device__ __forceinline__ void store_global(const float (&acc)[R][8], float* __restrict__ M, int ld, int64_t row0,
                                             int64_t n_rows, int col0, int tx, int ty) {
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int64_t row = row0 + ty * R + r;
    if (row >= n_rows) continue;
    float* p = M + row * ld + col0;
    *reinterpret_cast<float4*>(p + tx * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    *reinterpret_cast<float4*>(p + 64 + tx * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
  }
}

template <int R, typename F>
__device__ __forceinline__ void apply(float (&acc)[R][8], F f) {
#pragma unroll
  for (int r = 0; r < R; r++)
#pragma unroll
    for (int c = 0; c < 8; c++) acc[r][c] = f(acc[r][c]);
}

// One Linear over the tile: for every 128-column block, acc = Xin[TM x K] Wt[K x 128 block] + bias, then epi(acc, col0).
template <int R, typename Epi>
__device__ __forceinline__ void dense(const float* Xin, int XS, int K, const float* __restrict__ Wt,
                                      const float* __restrict__ bias, int N, float* Wsm, int tid, Epi epi) {
  for (int nb = 0; nb < N / NB; nb++) {
    float acc[R][8];
    gemm_block<R>(acc, Xin, XS, K, Wt, N, nb, Wsm, tid);
    add_vec<R>(acc, bias + nb * NB, tid & 15);
    epi(acc, nb * NB);
  }
}

// LayerNorm (eps 1e-5, biased variance, two-pass) of the N columns of every tile row, one warp per row.  The result goes
// back into X (in_place) and / or to the global matrix out[row0 + r][0..N) (rows past n_rows are not stored).
template <int R>
__device__ __forceinline__ void ln_rows(float* X, int XS, int N, const float* __restrict__ w, const float* __restrict__ b,
                                        bool in_place, float* __restrict__ out, int64_t row0, int64_t n_rows, int tid) {
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int r = warp; r < 16 * R; r += NT / 32) {
    float* x = X + r * XS;
    float s = 0.f;
    for (int c = lane; c < N; c += 32) s += x[c];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)N;
    float q = 0.f;
    for (int c = lane; c < N; c += 32) {
      const float d = x[c] - mean;
      q += d * d;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.f / sqrtf(q / (float)N + 1e-5f);
    const bool st = out != nullptr && row0 + r < n_rows;
    for (int c = lane; c < N; c += 32) {
      const float y = (x[c] - mean) * rstd * w[c] + b[c];
      if (in_place) x[c] = y;
      if (st) out[(row0 + r) * N + c] = y;
    }
  }
  __syncthreads();
}

// eval-mode BatchNorm1d: y = (x - running_mean) / sqrt(running_var + eps) * w + b, column-wise
template <int R>
__device__ __forceinline__ void bn_rows(float* X, int XS, int N, const float* __restrict__ w, const float* __restrict__ b,
                                        const float* __restrict__ rm, const float* __restrict__ rv,
                                        float* __restrict__ out, int64_t row0, int64_t n_rows, int tid) {
  __syncthreads();
  for (int idx = tid; idx < 16 * R * N; idx += NT) {
    const int r = idx / N, c = idx - r * N;
    const float y = (X[r * XS + c] - rm[c]) / sqrtf(rv[c] + 1e-5f) * w[c] + b[c];
    X[r * XS + c] = y;
    if (row0 + r < n_rows) out[(row0 + r) * N + c] = y;
  }
  __syncthreads();
}

struct WDims {
  int D, H, De, Kin, XS;
};

// ------------------------------------------------------------------------------------------------------------------
// edge features + encoder + LayerNorm -> e[E, De]          (nn_module.py:322-336 / :603-634, :598-600, :646)
// ------------------------------------------------------------------------------------------------------------------
struct WEncArgs {
  const float *enc0_t, *enc0_b, *enc2_t, *enc2_b, *enc4_t, *enc4_b, *eln_w, *eln_b, *centers;
  float length_mean, length_std;
  int n_edge_in, use_bond, expand_edge, dynbox;
  float box[3];
};

template <int R>
__global__ void __launch_bounds__(NT) k_wide_encode(WDims d, WEncArgs a, const float4* __restrict__ pos,
                                                    const int* __restrict__ col, const int* __restrict__ edst,
                                                    const int* __restrict__ n_edges_dev, const int* __restrict__ orig_id,
                                                    const int* __restrict__ bond, int atoms_per_frame,
                                                    float* __restrict__ e_out) {
  constexpr int TM = 16 * R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm sm = carve_smem<R>(smem_raw, d.XS);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int XS = d.XS;
  const int E = *n_edges_dev;
  const int ntiles = (E + TM - 1) / TM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int e0 = tile * TM;
    if (tid < TM) {
      const int e = e0 + tid;
      float ux = 0.f, uy = 0.f, uz = 0.f, dh = 0.f, flag = 0.f;
      if (e < E) {
        const int c = edst[e], n = col[e];
        const float4 pc = pos[c], pn = pos[n];
        // static box: rel = pos[neigh] - pos[center], remainder(rel + L/2, L) - L/2 (nn_module.py:615-621);
        // dynamic box: get_neighbor hands over the min-image of pos[center] - pos[neigh] and calc_edge_feat flips its
        // sign (md_module.py:65-66, nn_module.py:327)
        float r[3] = {pn.x - pc.x, pn.y - pc.y, pn.z - pc.z};
        if (a.dynbox) { r[0] = pc.x - pn.x; r[1] = pc.y - pn.y; r[2] = pc.z - pn.z; }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const float half = 0.5f * a.box[k];
          const float t = __fadd_rn(r[k], half);
          float m = fmodf(t, a.box[k]);
          if (m < 0.f) m = __fadd_rn(m, a.box[k]);
          r[k] = __fsub_rn(m, half);
          if (a.dynbox) r[k] = -r[k];
        }
        const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r[0], r[0]), __fmul_rn(r[1], r[1])), __fmul_rn(r[2], r[2])));
        const float den = dist + 1e-8f;
        ux = r[0] / den; uy = r[1] / den; uz = r[2] / den;
        dh = (dist - a.length_mean) / a.length_std;
        if (a.use_bond) {
          const int ic = orig_id ? orig_id[c] : c, in = orig_id ? orig_id[n] : n;
          if (ic / atoms_per_frame == in / atoms_per_frame) {
            const int lc = ic % atoms_per_frame, ln = in % atoms_per_frame;
#pragma unroll
            for (int k = 0; k < GAMD_MAX_BOND; k++) flag = (bond[lc * GAMD_MAX_BOND + k] == ln) ? 1.f : flag;
          }
        }
      }
      float* x = sm.A + tid * XS;
      x[0] = ux; x[1] = uy; x[2] = uz; x[3] = dh;
      sm.dh[tid] = dh;
      sm.dst[tid] = (e < E) ? 1 : 0;
      if (a.use_bond) x[4 + (a.expand_edge ? GAMD_NRBF : 0)] = flag;
    }
    __syncthreads();
    if (a.expand_edge) {
      for (int idx = tid; idx < TM * GAMD_NRBF; idx += NT) {
        const int m = idx / GAMD_NRBF, c = idx - m * GAMD_NRBF;
        const float rr = sm.dh[m] - a.centers[c];
        sm.A[m * XS + 4 + c] = sm.dst[m] ? expf(-40.f * (rr * rr)) : 0.f;   // exp(-gamma r^2), gamma = 1 / 0.025
      }
    }
    const int npad = d.Kin - a.n_edge_in;
    for (int idx = tid; idx < TM * npad; idx += NT) {
      const int m = idx / npad, c = idx - m * npad;
      sm.A[m * XS + a.n_edge_in + c] = 0.f;
    }
    __syncthreads();
    dense<R>(sm.A, XS, d.Kin, a.enc0_t, a.enc0_b, d.H, sm.W, tid, [&](float (&acc)[R][8], int col0) {
      apply<R>(acc, gelu_f);
      store_smem<R>(acc, sm.B, XS, col0, tx, ty);
    });
    dense<R>(sm.B, XS, d.H, a.enc2_t, a.enc2_b, d.H, sm.W, tid, [&](float (&acc)[R][8], int col0) {
      apply<R>(acc, gelu_f);
      store_smem<R>(acc, sm.A, XS, col0, tx, ty);
    });
    dense<R>(sm.A, XS, d.H, a.enc4_t, a.enc4_b, d.De, sm.W, tid, [&](float (&acc)[R][8], int col0) {
      store_smem<R>(acc, sm.B, XS, col0, tx, ty);
    });
    ln_rows<R>(sm.B, XS, d.De, a.eln_w, a.eln_b, false, e_out, e0, E, tid);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// one message-passing layer, edge part: edge_affine(e) + src_affine(hn)[src] + dst_affine(hn)[dst] -> theta_edge ->
// (update_edge: e <- edge_layer_norm(e_emb)) -> * hn[src] -> segmented sum over the receiver-sorted run
// (nn_module.py:130-143)
// ------------------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(NT) k_wide_edge(WDims d, LayerW w, float* __restrict__ e_emb,
                                                  const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                  const int* __restrict__ edst, const int* __restrict__ n_edges_dev,
                                                  const float* __restrict__ hn, const float* __restrict__ srcA,
                                                  const float* __restrict__ dstA, float* __restrict__ agg,
                                                  float* __restrict__ part) {
  constexpr int TM = 16 * R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm sm = carve_smem<R>(smem_raw, d.XS);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int XS = d.XS, D = d.D, H = d.H, De = d.De;
  const int E = *n_edges_dev;
  const int ntiles = (E + TM - 1) / TM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int e0 = tile * TM;
    const int q = De / 4;
    for (int idx = tid; idx < TM * q; idx += NT) {
      const int m = idx / q, c4 = idx - m * q;
      float* dstp = sm.A + m * XS + c4 * 4;
      if (e0 + m < E) cp_async16(dstp, e_emb + (size_t)(e0 + m) * De + c4 * 4);
      else *reinterpret_cast<float4*>(dstp) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
    if (tid < TM) {
      const int e = e0 + tid;
      sm.src[tid] = e < E ? col[e] : -1;
      sm.dst[tid] = e < E ? edst[e] : -1;
    }
    cp_async_wait<0>();
    __syncthreads();
    int rs[R], rd[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      rs[r] = sm.src[ty * R + r];
      rd[r] = sm.dst[ty * R + r];
    }
    dense<R>(sm.A, XS, De, w.ea0_t, w.ea0_b, 128, sm.W, tid, [&](float (&acc)[R][8], int col0) {      // edge_affine.0
      apply<R>(acc, silu_f);
      store_smem<R>(acc, sm.B, XS, col0, tx, ty);
    });
    dense<R>(sm.B, XS, 128, w.ea2_t, w.ea2_b, H, sm.W, tid, [&](float (&acc)[R][8], int col0) {       // edge_affine.2
      add_rows<R>(acc, srcA, H, rs, col0, tx);
      add_rows<R>(acc, dstA, H, rd, col0, tx);
      apply<R>(acc, silu_f);                                                                        // theta_edge: act first
      store_smem<R>(acc, sm.A, XS, col0, tx, ty);
    });
    dense<R>(sm.A, XS, H, w.te1_t, w.te1_b, H, sm.W, tid, [&](float (&acc)[R][8], int col0) {
      apply<R>(acc, silu_f);
      store_smem<R>(acc, sm.B, XS, col0, tx, ty);
    });
    dense<R>(sm.B, XS, H, w.te3_t, w.te3_b, D, sm.W, tid, [&](float (&acc)[R][8], int col0) {
      store_smem<R>(acc, sm.A, XS, col0, tx, ty);
    });
    // the next layers see edge_layer_norm(e_emb) as their edge embedding (only this tile reads or writes these rows)
    if (w.uln_w) ln_rows<R>(sm.A, XS, D, w.uln_w, w.uln_b, false, e_emb, e0, E, tid);
    else __syncthreads();
    // message = hn[src] * e_emb
    const int qd = D / 4;
    for (int idx = tid; idx < TM * qd; idx += NT) {
      const int m = idx / qd, c4 = idx - m * qd;
      float4* p = reinterpret_cast<float4*>(sm.A + m * XS + c4 * 4);
      const int s = sm.src[m];
      if (s < 0) {
        *p = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        const float4 hv = __ldg(reinterpret_cast<const float4*>(hn + (size_t)s * D + c4 * 4));
        float4 v = *p;
        v.x *= hv.x; v.y *= hv.y; v.z *= hv.z; v.w *= hv.w;
        *p = v;
      }
    }
    __syncthreads();
    // segmented sum over the receiver-sorted run, one thread per feature column; runs cut by the tile go to `part`
    const int tile_end = min(e0 + TM, E);
    for (int c = tid; c < D; c += NT) {
      int cur = sm.dst[0];
      int seg_start = 0;
      float s = 0.f;
      for (int m = 0; m <= TM; m++) {
        const int dd = (m < TM) ? sm.dst[m] : -2;
        if (dd != cur) {
          if (cur >= 0) {
            const bool head = (seg_start == 0) && (row_ptr[cur] < e0);
            const bool tail = (e0 + m == tile_end) && (row_ptr[cur + 1] > tile_end);
            if (head) part[((size_t)tile * 2 + 0) * D + c] = s;
            else if (tail) part[((size_t)tile * 2 + 1) * D + c] = s;
            else agg[(size_t)cur * D + c] = s;
          }
          cur = dd;
          seg_start = m;
          s = 0.f;
          if (dd < 0) break;
        }
        s += sm.A[m * XS + c];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// node update: (first) h0 -> norm_0 -> affines; (middle) agg -> phi(...) + h -> norm_{l+1} -> affines; (last) decoder
// (nn_module.py:147, :198-206, :401-406)
// ------------------------------------------------------------------------------------------------------------------
struct WNodeArgs {
  LayerW cur, next;
  const float *dec0_t, *dec0_b, *dec2_w, *dec2_b;
  const float *node_emb, *nenc_w, *nenc_b;
};

template <int R, int MODE>   // 0 first, 1 middle, 2 last
__global__ void __launch_bounds__(NT) k_wide_node(WDims d, WNodeArgs a, int n_atoms, const int* __restrict__ n_edges_dev,
                                                  const int* __restrict__ row_ptr, const float4* __restrict__ pos_feat,
                                                  const float* __restrict__ agg, const float* __restrict__ part,
                                                  float* __restrict__ h, float* __restrict__ hn, float* __restrict__ srcA,
                                                  float* __restrict__ dstA, float* __restrict__ pd,
                                                  float* __restrict__ pred) {
  constexpr int TM = 16 * R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm sm = carve_smem<R>(smem_raw, d.XS);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int XS = d.XS, D = d.D, H = d.H;
  const int ntiles = (n_atoms + TM - 1) / TM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n0 = tile * TM;
    int rows[R];
#pragma unroll
    for (int r = 0; r < R; r++) rows[r] = (n0 + ty * R + r < n_atoms) ? n0 + ty * R + r : -1;
    if (MODE == 0) {
      // h0 = node_emb.repeat(N, 1) (nn_module.py:681) or node_encoder(type) (:401, :554)
      for (int idx = tid; idx < TM * D; idx += NT) {
        const int m = idx / D, c = idx - m * D;
        const int i = n0 + m;
        float v = 0.f;
        if (i < n_atoms) {
          v = a.nenc_w ? fmaf(pos_feat[i].w, a.nenc_w[c], a.nenc_b[c]) : a.node_emb[c];
          h[(size_t)i * D + c] = v;
        }
        sm.A[m * XS + c] = v;
      }
    } else {
      // agg rows: whole runs come from agg, runs cut by edge tiles from part
      const int qd = D / 4;
      for (int idx = tid; idx < TM * qd; idx += NT) {
        const int m = idx / qd, c4 = idx - m * qd;
        const int i = n0 + m;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n_atoms) {
          int rs = row_ptr[i], re = row_ptr[i + 1];
          if (*n_edges_dev == 0) re = rs;   // empty / overflowed edge list (k_nbr_guard)
          if (re > rs) {
            const int t0 = rs / TM, t1 = (re - 1) / TM;
            if (t0 == t1) {
              v = *reinterpret_cast<const float4*>(agg + (size_t)i * D + c4 * 4);
            } else {
              v = *reinterpret_cast<const float4*>(part + ((size_t)t0 * 2 + 1) * D + c4 * 4);
              for (int t = t0 + 1; t <= t1; t++) {
                const float4 u = *reinterpret_cast<const float4*>(part + ((size_t)t * 2 + 0) * D + c4 * 4);
                v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
              }
            }
          }
        }
        *reinterpret_cast<float4*>(sm.A + m * XS + c4 * 4) = v;
      }
      __syncthreads();
      dense<R>(sm.A, XS, D, a.cur.pedge_t, a.cur.pedge_b, H, sm.W, tid, [&](float (&acc)[R][8], int col0) {   // phi_edge(agg)
        add_rows<R>(acc, pd, H, rows, col0, tx);                                                           // + phi_dst(hn)
        apply<R>(acc, silu_f);
        store_smem<R>(acc, sm.B, XS, col0, tx, ty);
      });
      dense<R>(sm.B, XS, H, a.cur.phi_t, a.cur.phi_b, D, sm.W, tid, [&](float (&acc)[R][8], int col0) {      // phi.1
        add_rows<R>(acc, h, D, rows, col0, tx);                                                            // residual: raw h
        if (MODE == 1) store_global<R>(acc, h, D, n0, n_atoms, col0, tx, ty);
        store_smem<R>(acc, sm.A, XS, col0, tx, ty);
      });
    }
    if (MODE != 2) {
      if (a.next.bn_mean) bn_rows<R>(sm.A, XS, D, a.next.ln_w, a.next.ln_b, a.next.bn_mean, a.next.bn_var, hn, n0, n_atoms, tid);
      else ln_rows<R>(sm.A, XS, D, a.next.ln_w, a.next.ln_b, true, hn, n0, n_atoms, tid);
      dense<R>(sm.A, XS, D, a.next.src_t, a.next.src_b, H, sm.W, tid, [&](float (&acc)[R][8], int col0) {
        store_global<R>(acc, srcA, H, n0, n_atoms, col0, tx, ty);
      });
      dense<R>(sm.A, XS, D, a.next.dst_t, a.next.dst_b, H, sm.W, tid, [&](float (&acc)[R][8], int col0) {
        store_global<R>(acc, dstA, H, n0, n_atoms, col0, tx, ty);
      });
      dense<R>(sm.A, XS, D, a.next.pdst_t, a.next.pdst_b, H, sm.W, tid, [&](float (&acc)[R][8], int col0) {
        store_global<R>(acc, pd, H, n0, n_atoms, col0, tx, ty);
      });
    } else {
      // graph_decoder: Linear -> GELU -> Linear(3)   (nn_module.py:320, :406)
      dense<R>(sm.A, XS, D, a.dec0_t, a.dec0_b, H, sm.W, tid, [&](float (&acc)[R][8], int col0) {
        apply<R>(acc, gelu_f);
        store_smem<R>(acc, sm.B, XS, col0, tx, ty);
      });
      __syncthreads();
      const int warp = tid >> 5, lane = tid & 31;
      for (int r = warp; r < TM; r += NT / 32) {
        float o[3] = {0.f, 0.f, 0.f};
        for (int c = lane; c < H; c += 32) {
          const float x = sm.B[r * XS + c];
#pragma unroll
          for (int k = 0; k < 3; k++) o[k] = fmaf(x, a.dec2_w[k * H + c], o[k]);
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) o[k] += __shfl_xor_sync(0xffffffffu, o[k], off);
        if (lane == 0 && n0 + r < n_atoms) {
          pred[(size_t)(n0 + r) * 3 + 0] = o[0] + a.dec2_b[0];
          pred[(size_t)(n0 + r) * 3 + 1] = o[1] + a.dec2_b[1];
          pred[(size_t)(n0 + r) * 3 + 2] = o[2] + a.dec2_b[2];
        }
      }
    }
    __syncthreads();
  }
}

size_t smem_bytes(int R, int XS) { return (size_t)(2 * 16 * R * XS + 2 * KC * NB) * 4 + (size_t)3 * 16 * R * 4; }

WDims dims_of(const gamd_ctx* ctx) {
  WDims d;
  d.D = ctx->desc.encoding_size;
  d.H = ctx->desc.hidden_dim;
  d.De = ctx->desc.edge_dim;
  d.Kin = ctx->desc.expand_edge ? 64 : 32;
  d.XS = ctx->wide_xs;
  return d;
}

template <int R>
int set_attrs(gamd_ctx* ctx, int bytes) {
  GAMD_CUDA(cudaFuncSetAttribute(k_wide_encode<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  GAMD_CUDA(cudaFuncSetAttribute(k_wide_edge<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  GAMD_CUDA(cudaFuncSetAttribute(k_wide_node<R, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  GAMD_CUDA(cudaFuncSetAttribute(k_wide_node<R, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  GAMD_CUDA(cudaFuncSetAttribute(k_wide_node<R, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

int wide_attrs(gamd_ctx* ctx) {
  if (ctx->attr_mask & GAMD_ATTR_WIDE) return 0;
  const int bytes = (int)smem_bytes(ctx->wide_r, ctx->wide_xs);
  int rc = ctx->wide_r == 4 ? set_attrs<4>(ctx, bytes) : ctx->wide_r == 2 ? set_attrs<2>(ctx, bytes) : set_attrs<1>(ctx, bytes);
  if (rc) return rc;
  ctx->attr_mask |= GAMD_ATTR_WIDE;
  return 0;
}

WNodeArgs node_args(const ModelW& mw) {
  WNodeArgs na{};
  na.dec0_t = mw.dec0_t; na.dec0_b = mw.dec0_b; na.dec2_w = mw.dec2_w; na.dec2_b = mw.dec2_b;
  na.node_emb = mw.node_emb; na.nenc_w = mw.nenc_w; na.nenc_b = mw.nenc_b;
  return na;
}

template <int R>
int launch_begin(gamd_ctx* ctx, const float4* pos_feat, const int* orig_id, int64_t n_atoms, int atoms_per_frame,
                 const float box[3], cudaStream_t st) {
  const ModelW& mw = ctx->mw;
  const WDims d = dims_of(ctx);
  const size_t smem = smem_bytes(R, d.XS);
  WEncArgs ea{mw.enc0_t, mw.enc0_b, mw.enc2_t, mw.enc2_b, mw.enc4_t, mw.enc4_b, mw.eln_w, mw.eln_b, mw.centers,
              mw.length_mean, mw.length_std, mw.n_edge_in, mw.use_bond, mw.expand_edge,
              mw.kind == GAMD_MODEL_DYNBOX ? 1 : 0, {box[0], box[1], box[2]}};
  prof_mark(ctx, "edge_encode", st);
  k_wide_encode<R><<<ctx->sm_count * 2, NT, smem, st>>>(d, ea, pos_feat, ctx->col_idx, ctx->edge_dst, ctx->n_edges, orig_id,
                                                        ctx->d_bond, atoms_per_frame, ctx->e_emb);
  GAMD_LAUNCH_CHECK();
  prof_mark(ctx, "edge_encode", st);
  WNodeArgs na = node_args(mw);
  na.next = mw.layer[0];
  const int tiles = ceil_div(n_atoms, 16 * R);
  prof_mark(ctx, "node_update", st);
  k_wide_node<R, 0><<<tiles < ctx->sm_count * 2 ? tiles : ctx->sm_count * 2, NT, smem, st>>>(
      d, na, (int)n_atoms, ctx->n_edges, ctx->row_ptr, pos_feat, ctx->agg, ctx->part, ctx->h, ctx->hn, ctx->srcA, ctx->dstA,
      ctx->pd, ctx->pred);
  GAMD_LAUNCH_CHECK();
  prof_mark(ctx, "node_update", st);
  return 0;
}

template <int R>
int launch_edges(gamd_ctx* ctx, int l, cudaStream_t st) {
  const WDims d = dims_of(ctx);
  prof_mark(ctx, "mp_edge", st);
  k_wide_edge<R><<<ctx->sm_count * 2, NT, smem_bytes(R, d.XS), st>>>(d, ctx->mw.layer[l], ctx->e_emb, ctx->row_ptr,
                                                                     ctx->col_idx, ctx->edge_dst, ctx->n_edges, ctx->hn,
                                                                     ctx->srcA, ctx->dstA, ctx->agg, ctx->part);
  GAMD_LAUNCH_CHECK();
  prof_mark(ctx, "mp_edge", st);
  return 0;
}

template <int R>
int launch_nodes(gamd_ctx* ctx, int l, const float4* pos_feat, int64_t n_atoms, cudaStream_t st) {
  const ModelW& mw = ctx->mw;
  const WDims d = dims_of(ctx);
  const size_t smem = smem_bytes(R, d.XS);
  WNodeArgs na = node_args(mw);
  na.cur = mw.layer[l];
  const int tiles = ceil_div(n_atoms, 16 * R);
  const int grid = tiles < ctx->sm_count * 2 ? tiles : ctx->sm_count * 2;
  prof_mark(ctx, "node_update", st);
  if (l + 1 < mw.n_layers) {
    na.next = mw.layer[l + 1];
    k_wide_node<R, 1><<<grid, NT, smem, st>>>(d, na, (int)n_atoms, ctx->n_edges, ctx->row_ptr, pos_feat, ctx->agg, ctx->part,
                                              ctx->h, ctx->hn, ctx->srcA, ctx->dstA, ctx->pd, ctx->pred);
  } else {
    k_wide_node<R, 2><<<grid, NT, smem, st>>>(d, na, (int)n_atoms, ctx->n_edges, ctx->row_ptr, pos_feat, ctx->agg, ctx->part,
                                              ctx->h, ctx->hn, ctx->srcA, ctx->dstA, ctx->pd, ctx->pred);
  }
  GAMD_LAUNCH_CHECK();
  prof_mark(ctx, "node_update", st);
  return 0;
}

}  // namespace

// rows per thread / shared-memory row stride for this model's widest layer; 0 when nothing fits
int wide_plan(int D, int H, int De, int* r_out, int* xs_out) {
  int mx = D > H ? D : H;
  if (De > mx) mx = De;
  if (mx < 128) mx = 128;
  const int XS = mx + 4;
  for (int R = 4; R >= 1; R >>= 1)
    if (smem_bytes(R, XS) <= 200 * 1024) {
      *r_out = R;
      *xs_out = XS;
      return 0;
    }
  return -1;
}

int wide_begin(gamd_ctx* ctx, const float4* pos_feat, const int* orig_id, int64_t n_atoms, int atoms_per_frame,
               const float box[3], cudaStream_t st) {
  int rc = wide_attrs(ctx);
  if (rc) return rc;
  ctx->model_atoms = n_atoms;
  switch (ctx->wide_r) {
    case 4: return launch_begin<4>(ctx, pos_feat, orig_id, n_atoms, atoms_per_frame, box, st);
    case 2: return launch_begin<2>(ctx, pos_feat, orig_id, n_atoms, atoms_per_frame, box, st);
    default: return launch_begin<1>(ctx, pos_feat, orig_id, n_atoms, atoms_per_frame, box, st);
  }
}

int wide_layer_edges(gamd_ctx* ctx, int l, cudaStream_t st) {
  switch (ctx->wide_r) {
    case 4: return launch_edges<4>(ctx, l, st);
    case 2: return launch_edges<2>(ctx, l, st);
    default: return launch_edges<1>(ctx, l, st);
  }
}

int wide_layer_nodes(gamd_ctx* ctx, int l, const float4* pos_feat, int64_t n_atoms, cudaStream_t st) {
  switch (ctx->wide_r) {
    case 4: return launch_nodes<4>(ctx, l, pos_feat, n_atoms, st);
    case 2: return launch_nodes<2>(ctx, l, pos_feat, n_atoms, st);
    default: return launch_nodes<1>(ctx, l, pos_feat, n_atoms, st);
  }
}
