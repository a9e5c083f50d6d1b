// Stage 2, fp32 CUDA-core path (the parity anchor): MDNet forward over a receiver-sorted CSR.
//
// Restates the eval-mode forward of code/nn_module.py (SimpleMDNetNew :672-685,
// WaterMDNetNew :545-558; calc_edge_feat :603-634; SmoothConvLayerNew.forward :108-148;
// SmoothConvBlockNew.forward :198-206) as four fused kernels:
//
//   k_edge_encode   edge features + 3-layer GELU encoder + LayerNorm -> e[E,128]      (K5+K6)
//   k_node_update   (first)  h0 -> LN_0 -> src/dst/phi_dst affines                     (K7)
//   k_mp_edge       edge_affine(e) + srcA[src] + dstA[dst] -> theta_edge -> * hn[src]
//                   -> segmented sum over the receiver-sorted edge run, no atomics      (K8+K9)
//   k_node_update   agg -> phi(...) + residual -> next layer's LN + affines, or the
//                   force decoder on the last layer                                     (K10+K11)
//
// All GEMMs are 64-row x 128-column register-tiled FFMA tiles with the activation tile held
// in shared memory and the (pre-transposed) weights streamed through a double-buffered
// cp.async ring.  Arithmetic is fp32 throughout.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 64;        // rows (edges or nodes) per tile
constexpr int NF = GAMD_NF;   // 128 features
constexpr int XS = 132;       // shared-memory row stride of the activation tile (floats)
constexpr int KC = 32;        // weight k-chunk
constexpr int NT = 256;       // threads per CTA: 16 column groups x 16 row groups

struct Smem {
  float X[TM * XS];
  float W[2][KC * NF];
  int src[TM];
  int dst[TM];
  float dh[TM];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void load_w_chunk(float* Wbuf, const float* __restrict__ Wt, int chunk, int tid) {
  const float* src = Wt + (size_t)chunk * KC * NF;
#pragma unroll
  for (int i = 0; i < (KC * NF / 4) / NT; i++) {
    int idx = tid + i * NT;
    cp_async16(Wbuf + idx * 4, src + idx * 4);
  }
}

// acc[r][c]: rows ty*4+r, columns tx*4+c (c<4) and 64+tx*4+(c-4)
template <int KTOT>
__device__ __forceinline__ void tile_gemm(float (&acc)[4][8], const float* __restrict__ Wt, Smem& sm, int tid) {
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 8; c++) acc[r][c] = 0.f;
  constexpr int NCH = KTOT / KC;
  load_w_chunk(sm.W[0], Wt, 0, tid);
  cp_async_commit();
#pragma unroll 1
  for (int ch = 0; ch < NCH; ch++) {
    if (ch + 1 < NCH) {
      load_w_chunk(sm.W[(ch + 1) & 1], Wt, ch + 1, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* Wb = sm.W[ch & 1];
    const float* Xr = sm.X + (ty * 4) * XS + ch * KC;
#pragma unroll
    for (int k4 = 0; k4 < KC / 4; k4++) {
      float4 a[4];
#pragma unroll
      for (int r = 0; r < 4; r++) a[r] = *reinterpret_cast<const float4*>(Xr + r * XS + k4 * 4);
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        float4 w0 = *reinterpret_cast<const float4*>(Wb + (k4 * 4 + kk) * NF + tx * 4);
        float4 w1 = *reinterpret_cast<const float4*>(Wb + (k4 * 4 + kk) * NF + 64 + tx * 4);
#pragma unroll
        for (int r = 0; r < 4; r++) {
          float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
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

__device__ __forceinline__ float gelu_f(float x) { return x * 0.5f * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

__device__ __forceinline__ int col_of(int tx, int c) { return c < 4 ? tx * 4 + c : 64 + tx * 4 + (c - 4); }

__device__ __forceinline__ void add_bias(float (&acc)[4][8], const float* __restrict__ b, int tx) {
  float4 b0 = *reinterpret_cast<const float4*>(b + tx * 4);
  float4 b1 = *reinterpret_cast<const float4*>(b + 64 + tx * 4);
#pragma unroll
  for (int r = 0; r < 4; r++) {
    acc[r][0] += b0.x; acc[r][1] += b0.y; acc[r][2] += b0.z; acc[r][3] += b0.w;
    acc[r][4] += b1.x; acc[r][5] += b1.y; acc[r][6] += b1.z; acc[r][7] += b1.w;
  }
}

// acc[r][:] += M[row_r][:] for a row-major [*,128] global matrix; row < 0 skips
__device__ __forceinline__ void add_rows(float (&acc)[4][8], const float* __restrict__ M, const int (&rows)[4], int tx) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    if (rows[r] < 0) continue;
    const float* p = M + (size_t)rows[r] * NF;
    float4 v0 = __ldg(reinterpret_cast<const float4*>(p + tx * 4));
    float4 v1 = __ldg(reinterpret_cast<const float4*>(p + 64 + tx * 4));
    acc[r][0] += v0.x; acc[r][1] += v0.y; acc[r][2] += v0.z; acc[r][3] += v0.w;
    acc[r][4] += v1.x; acc[r][5] += v1.y; acc[r][6] += v1.z; acc[r][7] += v1.w;
  }
}

__device__ __forceinline__ void mul_rows(float (&acc)[4][8], const float* __restrict__ M, const int (&rows)[4], int tx) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    if (rows[r] < 0) {
#pragma unroll
      for (int c = 0; c < 8; c++) acc[r][c] = 0.f;
      continue;
    }
    const float* p = M + (size_t)rows[r] * NF;
    float4 v0 = __ldg(reinterpret_cast<const float4*>(p + tx * 4));
    float4 v1 = __ldg(reinterpret_cast<const float4*>(p + 64 + tx * 4));
    acc[r][0] *= v0.x; acc[r][1] *= v0.y; acc[r][2] *= v0.z; acc[r][3] *= v0.w;
    acc[r][4] *= v1.x; acc[r][5] *= v1.y; acc[r][6] *= v1.z; acc[r][7] *= v1.w;
  }
}

__device__ __forceinline__ void store_tile_smem(const float (&acc)[4][8], Smem& sm, int tx, int ty) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    float* p = sm.X + (ty * 4 + r) * XS;
    *reinterpret_cast<float4*>(p + tx * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    *reinterpret_cast<float4*>(p + 64 + tx * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
  }
}

__device__ __forceinline__ void store_rows_global(const float (&acc)[4][8], float* __restrict__ M, int64_t row0,
                                                  int64_t n_rows, int tx, int ty) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    int64_t row = row0 + ty * 4 + r;
    if (row >= n_rows) continue;
    float* p = M + row * NF;
    *reinterpret_cast<float4*>(p + tx * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    *reinterpret_cast<float4*>(p + 64 + tx * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
  }
}

// LayerNorm over the 128 columns of each row (eps 1e-5, biased variance), two-pass
__device__ __forceinline__ void layer_norm_rows(float (&acc)[4][8], const float* __restrict__ w,
                                                const float* __restrict__ b, int tx) {
  float4 w0 = *reinterpret_cast<const float4*>(w + tx * 4), w1 = *reinterpret_cast<const float4*>(w + 64 + tx * 4);
  float4 b0 = *reinterpret_cast<const float4*>(b + tx * 4), b1 = *reinterpret_cast<const float4*>(b + 64 + tx * 4);
  float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int r = 0; r < 4; r++) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; c++) s += acc[r][c];
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float mean = s * (1.f / NF);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      float d = acc[r][c] - mean;
      q += d * d;
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    float rstd = 1.f / sqrtf(q * (1.f / NF) + 1e-5f);
#pragma unroll
    for (int c = 0; c < 8; c++) acc[r][c] = (acc[r][c] - mean) * rstd * wv[c] + bv[c];
  }
}

template <typename F>
__device__ __forceinline__ void apply(float (&acc)[4][8], F f) {
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 8; c++) acc[r][c] = f(acc[r][c]);
}

// ------------------------------------------------------------------------------------------
// K5 + K6: edge features + encoder + LayerNorm
// ------------------------------------------------------------------------------------------
struct EncArgs {
  const float *enc0_t, *enc0_b, *enc2_t, *enc2_b, *enc4_t, *enc4_b, *eln_w, *eln_b, *centers;
  float length_mean, length_std;
  int n_edge_in, use_bond, expand_edge;
  float box[3];
  int dynbox;     // WaterMDDynamicBoxNet: rel = -(min-image of pos[center] - pos[neigh])  (nn_module.py:327, md_module.py:65-66)
};

__global__ void __launch_bounds__(NT) k_edge_encode(EncArgs a, const float4* __restrict__ pos,
                                                    const int* __restrict__ col, const int* __restrict__ edst,
                                                    const int* __restrict__ n_edges_dev,
                                                    const int* __restrict__ orig_id, const int* __restrict__ bond,
                                                    int atoms_per_frame, float* __restrict__ e_out,
                                                    uint8_t* __restrict__ e_blob) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int E = *n_edges_dev;
  const int ntiles = (E + TM - 1) / TM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int e0 = tile * TM;
    if (tid < TM) {
      int e = e0 + tid;
      float ux = 0.f, uy = 0.f, uz = 0.f, dh = 0.f, flag = 0.f;
      if (e < E) {
        int c = edst[e], n = col[e];
        float4 pc = pos[c], pn = pos[n];
        // rel = pos[neigh] - pos[center]; remainder(rel + L/2, L) - L/2   (nn_module.py:615-621)
        float r[3] = {pn.x - pc.x, pn.y - pc.y, pn.z - pc.z};
        if (a.dynbox) { r[0] = pc.x - pn.x; r[1] = pc.y - pn.y; r[2] = pc.z - pn.z; }
#pragma unroll
        for (int d = 0; d < 3; d++) {
          float half = 0.5f * a.box[d];
          float t = __fadd_rn(r[d], half);
          float m = fmodf(t, a.box[d]);
          if (m < 0.f) m = __fadd_rn(m, a.box[d]);
          r[d] = __fsub_rn(m, half);
          if (a.dynbox) r[d] = -r[d];
        }
        float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r[0], r[0]), __fmul_rn(r[1], r[1])), __fmul_rn(r[2], r[2])));
        float den = dist + 1e-8f;
        ux = r[0] / den; uy = r[1] / den; uz = r[2] / den;
        dh = (dist - a.length_mean) / a.length_std;
        if (a.use_bond) {
          int ic = orig_id ? orig_id[c] : c, in = orig_id ? orig_id[n] : n;
          if (ic / atoms_per_frame == in / atoms_per_frame) {
            int lc = ic % atoms_per_frame, ln = in % atoms_per_frame;
#pragma unroll
            for (int k = 0; k < GAMD_MAX_BOND; k++) flag = (bond[lc * GAMD_MAX_BOND + k] == ln) ? 1.f : flag;
          }
        }
      }
      float* x = sm.X + tid * XS;
      x[0] = ux; x[1] = uy; x[2] = uz; x[3] = dh;
      sm.dh[tid] = dh;
      sm.dst[tid] = (e < E) ? 1 : 0;
      int nb = 4 + (a.expand_edge ? GAMD_NRBF : 0);
      if (a.use_bond) x[nb] = flag;
    }
    __syncthreads();
    if (a.expand_edge) {
      for (int idx = tid; idx < TM * GAMD_NRBF; idx += NT) {
        int m = idx / GAMD_NRBF, c = idx - m * GAMD_NRBF;
        float rr = sm.dh[m] - a.centers[c];
        // torch.exp(coef * radial**2), coef = -1/gap = -40 (nn_module.py:261-263)
        sm.X[m * XS + 4 + c] = sm.dst[m] ? expf(-40.f * (rr * rr)) : 0.f;
      }
    }
    for (int idx = tid; idx < TM * (64 - a.n_edge_in); idx += NT) {
      int m = idx / (64 - a.n_edge_in), c = idx - m * (64 - a.n_edge_in);
      sm.X[m * XS + a.n_edge_in + c] = 0.f;
    }
    __syncthreads();
    float acc[4][8];
    tile_gemm<64>(acc, a.enc0_t, sm, tid);
    add_bias(acc, a.enc0_b, tx);
    apply(acc, gelu_f);
    store_tile_smem(acc, sm, tx, ty);
    tile_gemm<NF>(acc, a.enc2_t, sm, tid);
    add_bias(acc, a.enc2_b, tx);
    apply(acc, gelu_f);
    store_tile_smem(acc, sm, tx, ty);
    tile_gemm<NF>(acc, a.enc4_t, sm, tid);
    add_bias(acc, a.enc4_b, tx);
    layer_norm_rows(acc, a.eln_w, a.eln_b, tx);
    if (e_blob == nullptr) {
      store_rows_global(acc, e_out, e0, E, tx, ty);
    } else {
      // tensor-core path: bf16 hi / lo split, 64 KB blob per 128-edge tile laid out
      // [hi | lo][16 k-chunks of 8][128 rows][16 bytes] so that a thread-per-row reader is coalesced
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int e = e0 + ty * 4 + r;
        if (e >= E) continue;
        uint8_t* blob = e_blob + (size_t)(e >> 7) * 65536;
        const int rr = e & 127;
#pragma unroll
        for (int half = 0; half < 2; half++) {
          const int c0 = half * 64 + tx * 4;
          uint32_t h0, l0, h1, l1;
          tc::split_bf16(acc[r][half * 4 + 0], acc[r][half * 4 + 1], h0, l0);
          tc::split_bf16(acc[r][half * 4 + 2], acc[r][half * 4 + 3], h1, l1);
          const size_t off = ((size_t)(c0 >> 3) * 128 + rr) * 16 + (c0 & 7) * 2;
          *reinterpret_cast<uint2*>(blob + off) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(blob + 32768 + off) = make_uint2(l0, l1);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K8 + K9: per-layer edge chain + segmented reduction
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_mp_edge(LayerW w, const float* __restrict__ e_emb,
                                                const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                const int* __restrict__ edst, const int* __restrict__ n_edges_dev,
                                                const float* __restrict__ hn, const float* __restrict__ srcA,
                                                const float* __restrict__ dstA, float* __restrict__ agg,
                                                float* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int E = *n_edges_dev;
  const int ntiles = (E + TM - 1) / TM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int e0 = tile * TM;
    // stage e tile (cp.async) + edge endpoints
#pragma unroll
    for (int i = 0; i < (TM * NF / 4) / NT; i++) {
      int idx = tid + i * NT;
      int m = idx >> 5, c4 = idx & 31;
      float* dstp = sm.X + m * XS + c4 * 4;
      if (e0 + m < E) cp_async16(dstp, e_emb + (size_t)(e0 + m) * NF + c4 * 4);
      else *reinterpret_cast<float4*>(dstp) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
    if (tid < TM) {
      int e = e0 + tid;
      sm.src[tid] = e < E ? col[e] : -1;
      sm.dst[tid] = e < E ? edst[e] : -1;
    }
    cp_async_wait<0>();
    __syncthreads();
    int rs[4], rd[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      rs[r] = sm.src[ty * 4 + r];
      rd[r] = sm.dst[ty * 4 + r];
    }
    float acc[4][8];
    tile_gemm<NF>(acc, w.ea0_t, sm, tid);          // edge_affine.0
    add_bias(acc, w.ea0_b, tx);
    apply(acc, silu_f);
    store_tile_smem(acc, sm, tx, ty);
    tile_gemm<NF>(acc, w.ea2_t, sm, tid);          // edge_affine.2 + src_affine(hn)[src] + dst_affine(hn)[dst]
    add_bias(acc, w.ea2_b, tx);
    add_rows(acc, srcA, rs, tx);
    add_rows(acc, dstA, rd, tx);
    apply(acc, silu_f);                            // theta_edge: act first
    store_tile_smem(acc, sm, tx, ty);
    tile_gemm<NF>(acc, w.te1_t, sm, tid);
    add_bias(acc, w.te1_b, tx);
    apply(acc, silu_f);
    store_tile_smem(acc, sm, tx, ty);
    tile_gemm<NF>(acc, w.te3_t, sm, tid);
    add_bias(acc, w.te3_b, tx);
    mul_rows(acc, hn, rs, tx);                     // message = hn[src] * e_emb
    store_tile_smem(acc, sm, tx, ty);
    __syncthreads();
    // segmented sum over the receiver-sorted run; one thread per feature column
    if (tid < NF) {
      const int tile_end = min(e0 + TM, E);
      int cur = sm.dst[0];
      int seg_start = 0;
      float s = 0.f;
      for (int m = 0; m <= TM; m++) {
        int d = (m < TM) ? sm.dst[m] : -2;
        if (d != cur) {
          if (cur >= 0) {
            bool head = (seg_start == 0) && (row_ptr[cur] < e0);
            bool tail = (e0 + m == tile_end) && (row_ptr[cur + 1] > tile_end);
            if (head) part[((size_t)tile * 2 + 0) * NF + tid] = s;
            else if (tail) part[((size_t)tile * 2 + 1) * NF + tid] = s;
            else agg[(size_t)cur * NF + tid] = s;
          }
          cur = d;
          seg_start = m;
          s = 0.f;
          if (d < 0) break;
        }
        s += sm.X[m * XS + tid];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// K7 / K10 / K11: node update
// ------------------------------------------------------------------------------------------
struct NodeArgs {
  LayerW cur;    // layer whose aggregation was just computed (unused when FIRST)
  LayerW next;   // layer whose LN + affines are produced (unused when LAST)
  const float *dec0_t, *dec0_b, *dec2_w, *dec2_b;
  const float *node_emb, *nenc_w, *nenc_b;
};

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(NT) k_node_update(NodeArgs a, int n_atoms, int agg_tile,
                                                    const int* __restrict__ n_edges_dev,
                                                    const int* __restrict__ row_ptr,
                                                    const float4* __restrict__ pos_feat,
                                                    const float* __restrict__ agg, const float* __restrict__ part,
                                                    float* __restrict__ h, float* __restrict__ hn,
                                                    float* __restrict__ srcA, float* __restrict__ dstA,
                                                    float* __restrict__ pd, float* __restrict__ pred) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ntiles = (n_atoms + TM - 1) / TM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n0 = tile * TM;
    int rows[4];
#pragma unroll
    for (int r = 0; r < 4; r++) rows[r] = (n0 + ty * 4 + r < n_atoms) ? n0 + ty * 4 + r : -1;
    float acc[4][8];
    if (FIRST) {
      // h0 = node_emb.repeat(N,1) (nn_module.py:681) or node_encoder(type) (:554)
#pragma unroll
      for (int r = 0; r < 4; r++) {
        float t = (rows[r] >= 0 && a.nenc_w) ? pos_feat[rows[r]].w : 0.f;
#pragma unroll
        for (int c = 0; c < 8; c++) {
          int cc = col_of(tx, c);
          acc[r][c] = a.nenc_w ? fmaf(t, a.nenc_w[cc], a.nenc_b[cc]) : a.node_emb[cc];
        }
      }
      store_rows_global(acc, h, n0, n_atoms, tx, ty);
    } else {
      // assemble agg rows: whole-run rows come from agg, rows split across edge tiles from part
      for (int idx = tid; idx < TM * (NF / 4); idx += NT) {
        int m = idx >> 5, c4 = idx & 31;
        int i = n0 + m;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n_atoms) {
          int rs = row_ptr[i], re = row_ptr[i + 1];
          if (*n_edges_dev == 0) re = rs;   // empty / overflowed edge list (k_nbr_guard)
          if (re > rs) {
            int t0 = rs / agg_tile, t1 = (re - 1) / agg_tile;
            if (t0 == t1) {
              v = *reinterpret_cast<const float4*>(agg + (size_t)i * NF + c4 * 4);
            } else {
              v = *reinterpret_cast<const float4*>(part + ((size_t)t0 * 2 + 1) * NF + c4 * 4);
              for (int t = t0 + 1; t <= t1; t++) {
                float4 u = *reinterpret_cast<const float4*>(part + ((size_t)t * 2 + 0) * NF + c4 * 4);
                v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
              }
            }
          }
        }
        *reinterpret_cast<float4*>(sm.X + m * XS + c4 * 4) = v;
      }
      __syncthreads();
      tile_gemm<NF>(acc, a.cur.pedge_t, sm, tid);            // phi_edge(agg)
      add_bias(acc, a.cur.pedge_b, tx);
      add_rows(acc, pd, rows, tx);                           // + phi_dst(hn)
      apply(acc, silu_f);
      store_tile_smem(acc, sm, tx, ty);
      tile_gemm<NF>(acc, a.cur.phi_t, sm, tid);              // phi.1
      add_bias(acc, a.cur.phi_b, tx);
      add_rows(acc, h, rows, tx);                            // residual with the un-normalised h
      if (!LAST) store_rows_global(acc, h, n0, n_atoms, tx, ty);
    }
    if (!LAST) {
      layer_norm_rows(acc, a.next.ln_w, a.next.ln_b, tx);
      store_rows_global(acc, hn, n0, n_atoms, tx, ty);
      store_tile_smem(acc, sm, tx, ty);
      tile_gemm<NF>(acc, a.next.src_t, sm, tid);
      add_bias(acc, a.next.src_b, tx);
      store_rows_global(acc, srcA, n0, n_atoms, tx, ty);
      tile_gemm<NF>(acc, a.next.dst_t, sm, tid);
      add_bias(acc, a.next.dst_b, tx);
      store_rows_global(acc, dstA, n0, n_atoms, tx, ty);
      tile_gemm<NF>(acc, a.next.pdst_t, sm, tid);
      add_bias(acc, a.next.pdst_b, tx);
      store_rows_global(acc, pd, n0, n_atoms, tx, ty);
    } else {
      // graph_decoder: Linear -> GELU -> Linear(3)   (nn_module.py:601, :684)
      store_tile_smem(acc, sm, tx, ty);
      tile_gemm<NF>(acc, a.dec0_t, sm, tid);
      add_bias(acc, a.dec0_b, tx);
      apply(acc, gelu_f);
#pragma unroll
      for (int r = 0; r < 4; r++) {
        float o[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          float s = 0.f;
#pragma unroll
          for (int c = 0; c < 8; c++) s = fmaf(acc[r][c], a.dec2_w[k * NF + col_of(tx, c)], s);
#pragma unroll
          for (int off = 8; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
          o[k] = s + a.dec2_b[k];
        }
        if (tx == 0 && rows[r] >= 0) {
          pred[(size_t)rows[r] * 3 + 0] = o[0];
          pred[(size_t)rows[r] * 3 + 1] = o[1];
          pred[(size_t)rows[r] * 3 + 2] = o[2];
        }
      }
    }
    __syncthreads();
  }
}

template <typename K>
int set_smem(gamd_ctx* ctx, K kernel) {
  GAMD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  return 0;
}

}  // namespace

static int model_attrs(gamd_ctx* ctx) {
  if (!(ctx->attr_mask & GAMD_ATTR_FP32)) {
    int rc;
    if ((rc = set_smem(ctx, k_edge_encode))) return rc;
    if ((rc = set_smem(ctx, k_mp_edge))) return rc;
    if ((rc = set_smem(ctx, k_node_update<true, false>))) return rc;
    if ((rc = set_smem(ctx, k_node_update<false, false>))) return rc;
    if ((rc = set_smem(ctx, k_node_update<false, true>))) return rc;
    ctx->attr_mask |= GAMD_ATTR_FP32;
  }
  return 0;
}

static NodeArgs node_args(const ModelW& mw) {
  NodeArgs na{};
  na.dec0_t = mw.dec0_t; na.dec0_b = mw.dec0_b; na.dec2_w = mw.dec2_w; na.dec2_b = mw.dec2_b;
  na.node_emb = mw.node_emb; na.nenc_w = mw.nenc_w; na.nenc_b = mw.nenc_b;
  return na;
}

// edge encoder + layer-0 node prologue (h0, LN_0, src/dst/phi_dst affines) for n_atoms rows
int model_begin(gamd_ctx* ctx, const float4* pos_feat, const int* orig_id, int64_t n_atoms, int atoms_per_frame,
                const float box[3], cudaStream_t st) {
  if (ctx->wide) return wide_begin(ctx, pos_feat, orig_id, n_atoms, atoms_per_frame, box, st);
  int rc = model_attrs(ctx);
  if (rc) return rc;
  ctx->model_atoms = n_atoms;
  const ModelW& mw = ctx->mw;
  const int grid_edge = ctx->sm_count * 3;
  const int node_tiles = ceil_div(n_atoms, TM);
  const int grid_node = node_tiles < ctx->sm_count * 3 ? node_tiles : ctx->sm_count * 3;
  const size_t smem = sizeof(Smem);
  const bool tcpath = ctx->desc.precision != GAMD_PREC_FP32;
  const int agg_tile = tcpath ? 32 : GAMD_EDGE_TILE;
  prof_mark(ctx, "edge_encode", st);
  if (tcpath) {
    if ((rc = edge_encode_tc_launch(ctx, pos_feat, orig_id, atoms_per_frame, box, st))) return rc;
  } else {
    EncArgs ea{mw.enc0_t, mw.enc0_b, mw.enc2_t, mw.enc2_b, mw.enc4_t, mw.enc4_b, mw.eln_w, mw.eln_b, mw.centers,
               mw.length_mean, mw.length_std, mw.n_edge_in, mw.use_bond, mw.expand_edge, {box[0], box[1], box[2]},
               mw.kind == GAMD_MODEL_DYNBOX ? 1 : 0};
    k_edge_encode<<<grid_edge, NT, smem, st>>>(ea, pos_feat, ctx->col_idx, ctx->edge_dst, ctx->n_edges, orig_id,
                                               ctx->d_bond, atoms_per_frame, ctx->e_emb, nullptr);
    GAMD_LAUNCH_CHECK();
  }
  prof_mark(ctx, "edge_encode", st);
  NodeArgs na = node_args(mw);
  na.next = mw.layer[0];
  prof_mark(ctx, "node_update", st);
  if (tcpath) {
    if ((rc = node_update_tc_launch(ctx, 0, 0, pos_feat, n_atoms, st))) return rc;
  } else {
    k_node_update<true, false><<<grid_node, NT, smem, st>>>(na, (int)n_atoms, agg_tile, ctx->n_edges, ctx->row_ptr, pos_feat, ctx->agg,
                                                            ctx->part, ctx->h, ctx->hn, ctx->srcA, ctx->dstA, ctx->pd,
                                                            ctx->pred);
    GAMD_LAUNCH_CHECK();
  }
  prof_mark(ctx, "node_update", st);
  return 0;
}

// message-passing layer l, edge part: edge chain + segmented sum (which: -1 all tiles, 0 / 1 the interior / boundary
// tiles of the domain-decomposition split - tensor-core path only)
int model_layer_edges(gamd_ctx* ctx, int l, cudaStream_t st, int which) {
  const ModelW& mw = ctx->mw;
  const bool tcpath = ctx->desc.precision != GAMD_PREC_FP32;
  if (ctx->wide && which < 0) return wide_layer_edges(ctx, l, st);
  if (!tcpath && which >= 0) {
    ctx->err = "tile-split layers need a tensor-core precision (bf16x3 / bf16)";
    return GAMD_EUNSUPPORTED;
  }
  prof_mark(ctx, "mp_edge", st);
  if (tcpath) {
    int rc = mp_edge_tc_launch(ctx, l, st, which);
    if (rc) return rc;
  } else {
    k_mp_edge<<<ctx->sm_count * 3, NT, sizeof(Smem), st>>>(mw.layer[l], ctx->e_emb, ctx->row_ptr, ctx->col_idx,
                                                            ctx->edge_dst, ctx->n_edges, ctx->hn, ctx->srcA, ctx->dstA,
                                                            ctx->agg, ctx->part);
    GAMD_LAUNCH_CHECK();
  }
  prof_mark(ctx, "mp_edge", st);
  return 0;
}

// message-passing layer l, node part: next layer's LN + affines, or the force decoder after the last layer
int model_layer_nodes(gamd_ctx* ctx, int l, const float4* pos_feat, int64_t n_atoms, cudaStream_t st) {
  if (ctx->wide) return wide_layer_nodes(ctx, l, pos_feat, n_atoms, st);
  const ModelW& mw = ctx->mw;
  const int node_tiles = ceil_div(n_atoms, TM);
  const int grid_node = node_tiles < ctx->sm_count * 3 ? node_tiles : ctx->sm_count * 3;
  const size_t smem = sizeof(Smem);
  const bool tcpath = ctx->desc.precision != GAMD_PREC_FP32;
  const int agg_tile = tcpath ? 32 : GAMD_EDGE_TILE;
  prof_mark(ctx, "node_update", st);
  NodeArgs na = node_args(mw);
  na.cur = mw.layer[l];
  if (tcpath) {
    int rc = node_update_tc_launch(ctx, l + 1 < mw.n_layers ? 1 : 2, l, pos_feat, n_atoms, st);
    if (rc) return rc;
    prof_mark(ctx, "node_update", st);
    return 0;
  }
  if (l + 1 < mw.n_layers) {
    na.next = mw.layer[l + 1];
    k_node_update<false, false><<<grid_node, NT, smem, st>>>(na, (int)n_atoms, agg_tile, ctx->n_edges, ctx->row_ptr, pos_feat, ctx->agg,
                                                             ctx->part, ctx->h, ctx->hn, ctx->srcA, ctx->dstA, ctx->pd,
                                                             ctx->pred);
  } else {
    k_node_update<false, true><<<grid_node, NT, smem, st>>>(na, (int)n_atoms, agg_tile, ctx->n_edges, ctx->row_ptr, pos_feat, ctx->agg,
                                                            ctx->part, ctx->h, ctx->hn, ctx->srcA, ctx->dstA, ctx->pd,
                                                            ctx->pred);
  }
  GAMD_LAUNCH_CHECK();
  prof_mark(ctx, "node_update", st);
  return 0;
}

int model_layer(gamd_ctx* ctx, int l, const float4* pos_feat, int64_t n_atoms, cudaStream_t st) {
  int rc = model_layer_edges(ctx, l, st, -1);
  if (rc) return rc;
  return model_layer_nodes(ctx, l, pos_feat, n_atoms, st);
}

int model_forward(gamd_ctx* ctx, const float4* pos_feat, const float* /*feat*/, const int* orig_id,
                       int64_t n_atoms, int atoms_per_frame, const float box[3], cudaStream_t st) {
  int rc = model_begin(ctx, pos_feat, orig_id, n_atoms, atoms_per_frame, box, st);
  if (rc) return rc;
  for (int l = 0; l < ctx->mw.n_layers; l++)
    if ((rc = model_layer(ctx, l, pos_feat, n_atoms, st))) return rc;
  return 0;
}
