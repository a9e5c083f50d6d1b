// Stage 1: periodic cell-list neighbor search -> receiver-sorted CSR.
//
// Replaces jax-md's partition.neighbor_list + the exact mask + the COO conversion of the
// reference (code/graph_utils.py:29-61, code/LJ/train_network_lj.py:166-199) and the O(N^2)
// torch path (code/md_module.py:63-126).  Pipeline (no host synchronisation anywhere):
//
//   bin    wrap positions (fmod semantics of jnp.mod / np.mod), cell key per atom
//   sort   stable LSD radix sort of (cell key, atom id), 8 bits per pass
//   gather float4 positions in cell order, cell_start table
//   count  one warp per centre, 27-cell sweep over contiguous x-runs, exact fp32 predicate
//   scan   degrees -> row_ptr
//   fill   same sweep, warp-ballot compaction straight into col_idx / edge_dst
//
// The predicate is evaluated for every DIRECTED pair from its own centre, one IEEE rounding
// per operation (no FMA contraction), so the edge set is bit-identical to oracle/neighbor.py.
#include "common.cuh"
#include <algorithm>
#include <cstring>

// ------------------------------------------------------------------------------------------
// arithmetic shared by bin / sweep / export
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float wrap_pos(float x, float L) {
  // jnp.mod / np.mod / torch.remainder: fmod, then add L when the sign differs (L > 0)
  float r = fmodf(x, L);
  if (r < 0.f) r = __fadd_rn(r, L);
  return r + 0.f;  // -0 -> +0
}

template <bool GENERAL>
__device__ __forceinline__ float min_image(float t, float L, float half) {
  // jax-md periodic_displacement: mod(dR + L/2, L) - L/2, each op rounded once
  t = __fadd_rn(t, half);
  if (GENERAL) {
    t = wrap_pos(t, L);
  } else {
    // wrapped inputs: t in [-L/2, 3L/2]; this branch form is bit-identical to fmod-based mod
    if (t < 0.f) t = __fadd_rn(t, L);
    else if (t >= L) t = __fsub_rn(t, L);
  }
  return __fsub_rn(t, half);
}

template <bool GENERAL>
__device__ __forceinline__ float pair_dr2(const float4& pc, const float4& pn, const NbrParams& p) {
  float tx = min_image<GENERAL>(__fsub_rn(pc.x, pn.x), p.box[0], p.half[0]);
  float ty = min_image<GENERAL>(__fsub_rn(pc.y, pn.y), p.box[1], p.half[1]);
  float tz = min_image<GENERAL>(__fsub_rn(pc.z, pn.z), p.box[2], p.half[2]);
  return __fadd_rn(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)), __fmul_rn(tz, tz));
}

__device__ __forceinline__ bool pass_pred(float dr2, const NbrParams& p) {
  if (p.flags & GAMD_NBR_LE) return __fsqrt_rn(dr2) <= p.rc;
  return dr2 < p.rc2;
}

__device__ __forceinline__ uint32_t cell_key(float wx, float wy, float wz, int frame, const NbrParams& p) {
  int cx = min(max((int)(wx * p.inv_cell[0]), 0), p.nc[0] - 1);
  int cy = min(max((int)(wy * p.inv_cell[1]), 0), p.nc[1] - 1);
  int cz = min(max((int)(wz * p.inv_cell[2]), 0), p.nc[2] - 1);
  return (uint32_t)(frame * p.cells_per_frame + (cz * p.nc[1] + cy) * p.nc[0] + cx);
}

// ------------------------------------------------------------------------------------------
// bin
// ------------------------------------------------------------------------------------------
__global__ void k_bin_f32(const float* __restrict__ pos, NbrParams p, float4* __restrict__ pos_nbr,
                          float4* __restrict__ pos_feat, uint32_t* __restrict__ keys,
                          uint32_t* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_atoms) return;
  float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
  float wx = wrap_pos(x, p.box[0]), wy = wrap_pos(y, p.box[1]), wz = wrap_pos(z, p.box[2]);
  bool raw = p.flags & GAMD_NBR_NOWRAP;
  pos_nbr[i] = make_float4(raw ? x : wx, raw ? y : wy, raw ? z : wz, __int_as_float(i));
  pos_feat[i] = make_float4(raw ? x : wx, raw ? y : wy, raw ? z : wz, 0.f);
  keys[i] = cell_key(wx, wy, wz, i / p.atoms_per_frame, p);
  vals[i] = i;
}

// engine path: fp64 state (x * scale = Angstrom).  Restates predict_forces exactly:
//   neighbor positions  = jnp.mod(f32(pos), f32(L))            (train_network_lj.py:188, graph_utils.py:37)
//   feature positions   = f32(np.mod(pos, L)) in float64 first  (train_network_lj.py:141-142)
__global__ void k_bin_f64(const double* __restrict__ x, double scale, double bx, double by, double bz,
                          NbrParams p, float4* __restrict__ pos_nbr, float4* __restrict__ pos_feat,
                          uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, const int* __restrict__ gate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_atoms || (gate && !*gate)) return;
  double px = x[3 * i] * scale, py = x[3 * i + 1] * scale, pz = x[3 * i + 2] * scale;
  float wx = wrap_pos((float)px, p.box[0]), wy = wrap_pos((float)py, p.box[1]), wz = wrap_pos((float)pz, p.box[2]);
  double fx = fmod(px, bx), fy = fmod(py, by), fz = fmod(pz, bz);
  if (fx < 0.0) fx += bx;
  if (fy < 0.0) fy += by;
  if (fz < 0.0) fz += bz;
  pos_nbr[i] = make_float4(wx, wy, wz, __int_as_float(i));
  pos_feat[i] = make_float4((float)fx, (float)fy, (float)fz, 0.f);
  keys[i] = cell_key(wx, wy, wz, i / p.atoms_per_frame, p);
  vals[i] = i;
}

// ------------------------------------------------------------------------------------------
// exclusive scan (int32), n inputs -> n+1 outputs (out[n] = total)
// ------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
  // exclusive scan of one int per thread over SCAN_THREADS threads
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  if (w == 0) {
    int ws = lane < (SCAN_THREADS / 32) ? s_warp[lane] : 0;
    int winc = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < (SCAN_THREADS / 32)) s_warp[lane] = winc - ws;
    if (lane == (SCAN_THREADS / 32) - 1) s_warp[SCAN_THREADS / 32] = winc;
  }
  __syncthreads();
  int res = inc - v + s_warp[w];
  total = s_warp[SCAN_THREADS / 32];
  __syncthreads();
  return res;
}

__global__ void k_scan_partial(const int* __restrict__ in, int64_t n, int* __restrict__ block_sums,
                               const int* __restrict__ gate) {
  __shared__ int s_warp[SCAN_THREADS / 32 + 1];
  if (gate && !*gate) return;
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) s += in[base + k];
  int total;
  block_exclusive_scan(s, s_warp, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void k_scan_top(int* __restrict__ block_sums, int nb, const int* __restrict__ gate) {
  __shared__ int s_warp[SCAN_THREADS / 32 + 1];
  if (gate && !*gate) return;
  int carry = 0;
  for (int base = 0; base < nb; base += SCAN_THREADS) {
    int i = base + threadIdx.x;
    int v = i < nb ? block_sums[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, s_warp, total);
    if (i < nb) block_sums[i] = ex + carry;
    carry += total;
  }
}

__global__ void k_scan_final(const int* __restrict__ in, int* __restrict__ out, int64_t n,
                             const int* __restrict__ block_sums, int* __restrict__ total_out,
                             const int* __restrict__ gate) {
  __shared__ int s_warp[SCAN_THREADS / 32 + 1];
  if (gate && !*gate) return;
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  int total;
  int ex = block_exclusive_scan(s, s_warp, total) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
    if (base + k == n - 1) {
      out[n] = ex;
      if (total_out) *total_out = ex;
    }
  }
}

__global__ void k_scan_empty(int* out, int* total_out, const int* gate) {
  if (gate && !*gate) return;
  out[0] = 0;
  if (total_out) *total_out = 0;
}

static int scan_with_total(gamd_ctx* ctx, const int* d_in, int* d_out, int64_t n, int* d_total, cudaStream_t st,
                           const int* gate = nullptr) {
  if (n == 0) {
    k_scan_empty<<<1, 1, 0, st>>>(d_out, d_total, gate);
    GAMD_LAUNCH_CHECK();
    return 0;
  }
  int nb = ceil_div(n, SCAN_TILE);
  int* sums = (int*)ctx->scan_tmp;
  k_scan_partial<<<nb, SCAN_THREADS, 0, st>>>(d_in, n, sums, gate);
  GAMD_LAUNCH_CHECK();
  k_scan_top<<<1, SCAN_THREADS, 0, st>>>(sums, nb, gate);
  GAMD_LAUNCH_CHECK();
  k_scan_final<<<nb, SCAN_THREADS, 0, st>>>(d_in, d_out, n, sums, d_total, gate);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int exclusive_scan_i32(gamd_ctx* ctx, const int* d_in, int* d_out, int64_t n, cudaStream_t st) {
  return scan_with_total(ctx, d_in, d_out, n, nullptr, st);
}

// ------------------------------------------------------------------------------------------
// stable LSD radix sort of (key, value) pairs, 8 bits per pass
// ------------------------------------------------------------------------------------------
#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_WARPS (RS_THREADS / 32)

__global__ void k_radix_hist(const uint32_t* __restrict__ keys, int n, int shift, int nblocks,
                             uint32_t* __restrict__ hist, const int* __restrict__ gate) {
  __shared__ uint32_t s_h[256];
  if (gate && !*gate) return;
  s_h[threadIdx.x] = 0;
  __syncthreads();
  int base = blockIdx.x * RS_TILE;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    int i = base + k * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * nblocks + blockIdx.x] = s_h[threadIdx.x];   // digit-major
}

__global__ void k_radix_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int n,
                                int shift, int nblocks, const uint32_t* __restrict__ offs,
                                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                const int* __restrict__ gate) {
  // warp w owns the contiguous sub-range [w*256, (w+1)*256) of the tile -> stable
  __shared__ uint32_t s_cnt[RS_WARPS][256];
  if (gate && !*gate) return;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int d = lane; d < 256; d += 32) s_cnt[w][d] = 0;
  __syncwarp();
  int base = blockIdx.x * RS_TILE + w * (RS_TILE / RS_WARPS);
  uint32_t k[RS_ITEMS], v[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    int i = base + r * 32 + lane;
    bool ok = i < n;
    k[r] = ok ? keys[i] : 0xffffffffu;
    v[r] = ok ? vals[i] : 0u;
    uint32_t d = (k[r] >> shift) & 255u;
    uint32_t peers = __match_any_sync(0xffffffffu, ok ? d : 256u);
    uint32_t before = __popc(peers & ((1u << lane) - 1u));
    uint32_t cur = 0;
    if (ok) cur = s_cnt[w][d];
    __syncwarp();
    if (ok && before == 0) s_cnt[w][d] = cur + __popc(peers);
    __syncwarp();
    rank[r] = cur + before;
  }
  __syncthreads();
  {  // per digit: exclusive prefix over warps + global offset
    int d = threadIdx.x;
    uint32_t run = offs[d * nblocks + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ww++) {
      uint32_t c = s_cnt[ww][d];
      s_cnt[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    int i = base + r * 32 + lane;
    if (i < n) {
      uint32_t d = (k[r] >> shift) & 255u;
      uint32_t dst = s_cnt[w][d] + rank[r];
      keys_out[dst] = k[r];
      vals_out[dst] = v[r];
    }
  }
}

// sorts ctx->keys[0]/vals[0]; returns the index (0/1) of the buffer holding the result
static int radix_sort_pairs(gamd_ctx* ctx, int n, int key_bits, cudaStream_t st, int* result_buf,
                            const int* gate = nullptr) {
  int cur = 0;
  int nblocks = ceil_div(n, RS_TILE);
  for (int shift = 0; shift < key_bits; shift += 8) {
    k_radix_hist<<<nblocks, RS_THREADS, 0, st>>>(ctx->keys[cur], n, shift, nblocks, ctx->radix_hist, gate);
    GAMD_LAUNCH_CHECK();
    int rc = scan_with_total(ctx, (const int*)ctx->radix_hist, (int*)ctx->radix_hist + 256 * nblocks + 8,
                             (int64_t)256 * nblocks, nullptr, st, gate);
    if (rc) return rc;
    k_radix_scatter<<<nblocks, RS_THREADS, 0, st>>>(ctx->keys[cur], ctx->vals[cur], n, shift, nblocks,
                                                    ctx->radix_hist + 256 * nblocks + 8, ctx->keys[cur ^ 1],
                                                    ctx->vals[cur ^ 1], gate);
    GAMD_LAUNCH_CHECK();
    cur ^= 1;
  }
  *result_buf = cur;
  return 0;
}

// ------------------------------------------------------------------------------------------
// gather into cell order + cell_start table
// ------------------------------------------------------------------------------------------
__global__ void k_gather_sorted(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int n,
                                int ncells, const float4* __restrict__ pos_nbr, const float4* __restrict__ pos_feat,
                                const float* __restrict__ feat, float4* __restrict__ pos_nbr_s,
                                float4* __restrict__ pos_feat_s, int* __restrict__ perm,
                                int* __restrict__ inv_perm, int* __restrict__ cell_start,
                                const int* __restrict__ gate) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n || (gate && !*gate)) return;
  int i = (int)vals[s];
  pos_nbr_s[s] = pos_nbr[i];
  float4 pf = pos_feat[i];
  pf.w = feat ? feat[i] : 0.f;
  pos_feat_s[s] = pf;
  perm[s] = i;
  inv_perm[i] = s;
  int k = (int)keys[s];
  int kp = s > 0 ? (int)keys[s - 1] : -1;
  for (int c = kp + 1; c <= k; c++) cell_start[c] = s;
  if (s == n - 1)
    for (int c = k + 1; c <= ncells; c++) cell_start[c] = n;
}

// ------------------------------------------------------------------------------------------
// 27-cell sweep, one warp per centre
// ------------------------------------------------------------------------------------------
template <bool WRITE, bool GENERAL>
__device__ __forceinline__ int sweep_range(int lo, int hi, int s, const float4& pc, const NbrParams& p,
                                           const float4* __restrict__ pos, int lane, int base, int cap,
                                           int* __restrict__ col, int* __restrict__ edst) {
  int cnt = 0;
  for (int j0 = lo; j0 < hi; j0 += 32) {
    int j = j0 + lane;
    bool ok = j < hi;
    if (ok) {
      float4 pn = pos[j];
      ok = pass_pred(pair_dr2<GENERAL>(pc, pn, p), p);
      if (j == s) ok = (p.flags & GAMD_NBR_SELF) != 0;
    }
    uint32_t m = __ballot_sync(0xffffffffu, ok);
    if (WRITE && ok) {
      int dst = base + cnt + __popc(m & ((1u << lane) - 1u));
      if (dst < cap) {
        col[dst] = j;
        if (edst) edst[dst] = s;
      }
    }
    cnt += __popc(m);
  }
  return cnt;
}

template <bool WRITE, bool GENERAL>
__device__ __forceinline__ void sweep_centre(int s, int lane, const NbrParams& p, const float4* __restrict__ pos,
                                             const uint32_t* __restrict__ keys, const int* __restrict__ cell_start,
                                             int* __restrict__ deg, const int* __restrict__ row_ptr, int* __restrict__ col,
                                             int* __restrict__ edst, int cap, int* __restrict__ err_flag) {
  float4 pc = pos[s];
  int key = (int)keys[s];
  int frame = key / p.cells_per_frame;
  int c = key - frame * p.cells_per_frame;
  int cx = c % p.nc[0];
  int cy = (c / p.nc[0]) % p.nc[1];
  int cz = c / (p.nc[0] * p.nc[1]);
  if (__float_as_int(pc.w) >= p.n_centers) {   // halo atom: neighbour only
    if (!WRITE && lane == 0) deg[s] = 0;
    return;
  }
  int base = WRITE ? row_ptr[s] : 0;
  if (WRITE && row_ptr[s + 1] > cap) {
    if (lane == 0) atomicOr(err_flag, 1);
    return;
  }
  int cnt = 0;
  int nz = p.nc[2] >= 3 ? 3 : 1, ny = p.nc[1] >= 3 ? 3 : 1;
  for (int iz = 0; iz < nz; iz++) {
    int zz = nz == 3 ? (cz + iz - 1 + p.nc[2]) % p.nc[2] : 0;
    for (int iy = 0; iy < ny; iy++) {
      int yy = ny == 3 ? (cy + iy - 1 + p.nc[1]) % p.nc[1] : 0;
      int rowbase = frame * p.cells_per_frame + (zz * p.nc[1] + yy) * p.nc[0];
      int ncx = p.nc[0];
      if (ncx <= 3) {  // whole x-row (ncx == 3 or single-cell fallback)
        cnt += sweep_range<WRITE, GENERAL>(cell_start[rowbase], cell_start[rowbase + ncx], s, pc, p, pos, lane,
                                            base + cnt, cap, col, edst);
      } else if (cx == 0) {
        cnt += sweep_range<WRITE, GENERAL>(cell_start[rowbase + ncx - 1], cell_start[rowbase + ncx], s, pc, p, pos,
                                            lane, base + cnt, cap, col, edst);
        cnt += sweep_range<WRITE, GENERAL>(cell_start[rowbase], cell_start[rowbase + 2], s, pc, p, pos, lane,
                                            base + cnt, cap, col, edst);
      } else if (cx == ncx - 1) {
        cnt += sweep_range<WRITE, GENERAL>(cell_start[rowbase + cx - 1], cell_start[rowbase + ncx], s, pc, p, pos,
                                            lane, base + cnt, cap, col, edst);
        cnt += sweep_range<WRITE, GENERAL>(cell_start[rowbase], cell_start[rowbase + 1], s, pc, p, pos, lane,
                                            base + cnt, cap, col, edst);
      } else {
        cnt += sweep_range<WRITE, GENERAL>(cell_start[rowbase + cx - 1], cell_start[rowbase + cx + 2], s, pc, p, pos,
                                            lane, base + cnt, cap, col, edst);
      }
    }
  }
  if (!WRITE && lane == 0) deg[s] = cnt;
}

// one warp per centre, grid-stride (a bounded grid: a gated-off launch of the skin path must cost microseconds, not the
// 0.27 ms that 125 000 empty CTAs take at 1 M atoms)
template <bool WRITE, bool GENERAL>
__global__ void __launch_bounds__(256) k_sweep(NbrParams p, const float4* __restrict__ pos,
                                               const uint32_t* __restrict__ keys,
                                               const int* __restrict__ cell_start, int* __restrict__ deg,
                                               const int* __restrict__ row_ptr, int* __restrict__ col,
                                               int* __restrict__ edst, int cap, int* __restrict__ err_flag,
                                               const int* __restrict__ gate) {
  if (gate && !*gate) return;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < p.n_atoms; s += warps)
    sweep_centre<WRITE, GENERAL>(s, lane, p, pos, keys, cell_start, deg, row_ptr, col, edst, cap, err_flag);
}

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------
int nbr_setup_params(gamd_ctx* ctx, int64_t n_atoms, int n_frames, const float box[3], float rc, int flags,
                     NbrParams* p) {
  if (n_atoms <= 0 || n_frames <= 0 || n_atoms % n_frames != 0) {
    ctx->err = "n_atoms must be a positive multiple of n_frames";
    return GAMD_EINVAL;
  }
  if (!(rc > 0.f) || !(box[0] > 0.f) || !(box[1] > 0.f) || !(box[2] > 0.f)) {
    ctx->err = "box and cutoff must be positive";
    return GAMD_EINVAL;
  }
  if (n_atoms > ctx->cap_atoms) {
    ctx->err = "n_atoms exceeds reserved capacity; call gamd_reserve";
    return GAMD_ECAPACITY;
  }
  int64_t cells = 1;
  for (int d = 0; d < 3; d++) {
    p->box[d] = box[d];
    p->half[d] = box[d] * 0.5f;
    // cell edge >= rc*(1+1e-3): a pair two cells apart can never pass the fp32 test by rounding
    int nc = (int)floor((double)box[d] / ((double)rc * 1.001));
    if (nc < 3) nc = 1;  // fewer than 3 cells: treat the axis as one cell (brute force along it)
    p->nc[d] = nc;
    p->inv_cell[d] = (float)((double)nc / (double)box[d]);
    cells *= nc;
  }
  // keep the cell table small relative to the atom count (huge sparse boxes)
  while (cells * n_frames > 4 * n_atoms + 64 && cells > 27) {
    int dmax = 0;
    for (int d = 1; d < 3; d++)
      if (p->nc[d] > p->nc[dmax]) dmax = d;
    if (p->nc[dmax] <= 3) break;
    cells /= p->nc[dmax];
    p->nc[dmax]--;
    cells *= p->nc[dmax];
    p->inv_cell[dmax] = (float)((double)p->nc[dmax] / (double)box[dmax]);
  }
  if (cells * n_frames + 1 > ctx->cap_cells) {
    ctx->err = "cell table exceeds reserved capacity";
    return GAMD_ECAPACITY;
  }
  p->cells_per_frame = (int)cells;
  p->n_atoms = (int)n_atoms;
  p->atoms_per_frame = (int)(n_atoms / n_frames);
  p->n_frames = n_frames;
  p->rc = rc;
  // python: `dr_2 < cutoff ** 2` - the square is taken in double, then rounded to fp32
  p->rc2 = (float)((double)rc * (double)rc);
  p->flags = flags;
  p->n_centers = (int)n_atoms;
  return 0;
}

int nbr_bin_f32(gamd_ctx* ctx, const float* d_pos, const NbrParams& p, cudaStream_t st) {
  k_bin_f32<<<ceil_div(p.n_atoms, 256), 256, 0, st>>>(d_pos, p, ctx->pos_nbr, ctx->pos_feat, ctx->keys[0],
                                                      ctx->vals[0]);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int nbr_bin_f64(gamd_ctx* ctx, const double* d_x, double scale, const double* box64, const NbrParams& p,
                cudaStream_t st) {
  k_bin_f64<<<ceil_div(p.n_atoms, 256), 256, 0, st>>>(d_x, scale, box64[0], box64[1], box64[2], p, ctx->pos_nbr,
                                                      ctx->pos_feat, ctx->keys[0], ctx->vals[0], nullptr);
  GAMD_LAUNCH_CHECK();
  return 0;
}

// edge capacity exceeded: flag it and publish an EMPTY edge list so that no downstream kernel can index past the
// reserved buffers (the error surfaces at the next gamd_check_async_errors as GAMD_ECAPACITY with the needed size)
__global__ void k_nbr_guard(const int* __restrict__ row_ptr, int n, int cap, int* __restrict__ n_edges,
                            int* __restrict__ err_flag) {
  if (row_ptr[n] > cap) {
    atomicOr(err_flag, 1);
    err_flag[1] = row_ptr[n];
    *n_edges = 0;
  }
}

int nbr_sort_and_sweep(gamd_ctx* ctx, const NbrParams& p, const float* d_feat, cudaStream_t st) {
  int n = p.n_atoms;
  int64_t ncells = (int64_t)p.cells_per_frame * p.n_frames;
  int bits = 1;
  while (((int64_t)1 << bits) < ncells) bits++;
  int buf = 0;
  int rc = 0;
  if (ncells > 1) {
    rc = radix_sort_pairs(ctx, n, bits, st, &buf);
    if (rc) return rc;
  }
  k_gather_sorted<<<ceil_div(n, 256), 256, 0, st>>>(ctx->keys[buf], ctx->vals[buf], n, (int)ncells, ctx->pos_nbr,
                                                    ctx->pos_feat, d_feat, ctx->pos_nbr_s, ctx->pos_feat_s, ctx->perm,
                                                    ctx->inv_perm, ctx->cell_start, nullptr);
  GAMD_LAUNCH_CHECK();
  int cap = (int)ctx->cap_edges;
  int blocks = std::min(ceil_div((int64_t)n * 32, 256), ctx->sm_count * 32);
  bool general = (p.flags & GAMD_NBR_NOWRAP) != 0;
  if (general)
    k_sweep<false, true><<<blocks, 256, 0, st>>>(p, ctx->pos_nbr_s, ctx->keys[buf], ctx->cell_start, ctx->deg, nullptr,
                                                 nullptr, nullptr, cap, ctx->err_flag, nullptr);
  else
    k_sweep<false, false><<<blocks, 256, 0, st>>>(p, ctx->pos_nbr_s, ctx->keys[buf], ctx->cell_start, ctx->deg,
                                                  nullptr, nullptr, nullptr, cap, ctx->err_flag, nullptr);
  GAMD_LAUNCH_CHECK();
  rc = scan_with_total(ctx, ctx->deg, ctx->row_ptr, n, ctx->n_edges, st);
  if (rc) return rc;
  k_nbr_guard<<<1, 1, 0, st>>>(ctx->row_ptr, n, cap, ctx->n_edges, ctx->err_flag);
  GAMD_LAUNCH_CHECK();
  if (general)
    k_sweep<true, true><<<blocks, 256, 0, st>>>(p, ctx->pos_nbr_s, ctx->keys[buf], ctx->cell_start, ctx->deg,
                                                ctx->row_ptr, ctx->col_idx, ctx->edge_dst, cap, ctx->err_flag, nullptr);
  else
    k_sweep<true, false><<<blocks, 256, 0, st>>>(p, ctx->pos_nbr_s, ctx->keys[buf], ctx->cell_start, ctx->deg,
                                                 ctx->row_ptr, ctx->col_idx, ctx->edge_dst, cap, ctx->err_flag, nullptr);
  GAMD_LAUNCH_CHECK();
  ctx->last_nbr = p;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Small frames (<= 1024 atoms per frame: LJ-258, TIP3P-774, the replica ensembles): one CTA per frame does the whole
// search - wrap, exact predicate on every ordered pair (with 3 cells per axis the 27-cell sweep visits every atom
// anyway), ballot masks kept in shared memory, block scan, CSR fill.  A single frame needs ONE launch instead of ~20
// (radix passes, scans, two sweeps); the step of these systems is launch-latency bound.
// Atoms keep the caller's order (perm = identity); a row lists its neighbours by ascending atom id.
// ------------------------------------------------------------------------------------------
#define SMALL_MAX 1024
#define SMALL_THREADS 1024

struct SmallSmem {
  int deg[SMALL_MAX + 1];
  int warp_sums[33];
  float4 pos[SMALL_MAX];
};

// pass 0: wrap + exact predicate.  grid = (frames, slices): a CTA loads its frame's positions into shared memory and
// its 32 warps take the centres of one slice (one centre per warp and round); ballot masks and degrees go to global.
__global__ void __launch_bounds__(SMALL_THREADS) k_nbr_small_count(const double* __restrict__ x, double scale, double bx,
                                                                   double by, double bz, NbrParams p,
                                                                   const float* __restrict__ feat, float4* __restrict__ pos_nbr,
                                                                   float4* __restrict__ pos_nbr_s, float4* __restrict__ pos_feat_s,
                                                                   int* __restrict__ perm, int* __restrict__ inv_perm,
                                                                   int* __restrict__ deg_g, uint32_t* __restrict__ gmask) {
  __shared__ float4 spos[SMALL_MAX];
  const int n = p.atoms_per_frame, words = (n + 31) >> 5;
  const int base = blockIdx.x * n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int i = tid; i < n; i += blockDim.x) {
    const int g = base + i;
    const double px = x[3 * g] * scale, py = x[3 * g + 1] * scale, pz = x[3 * g + 2] * scale;
    const float wx = wrap_pos((float)px, p.box[0]), wy = wrap_pos((float)py, p.box[1]), wz = wrap_pos((float)pz, p.box[2]);
    const float4 pn = make_float4(wx, wy, wz, __int_as_float(g));
    spos[i] = pn;
    if (blockIdx.y == 0) {          // one slice publishes the frame's arrays
      double fx = fmod(px, bx), fy = fmod(py, by), fz = fmod(pz, bz);
      if (fx < 0.0) fx += bx;
      if (fy < 0.0) fy += by;
      if (fz < 0.0) fz += bz;
      pos_nbr[g] = pn;
      pos_nbr_s[g] = pn;
      pos_feat_s[g] = make_float4((float)fx, (float)fy, (float)fz, feat ? feat[g] : 0.f);
      perm[g] = g;
      inv_perm[g] = g;
    }
  }
  __syncthreads();
  const int per = (n + gridDim.y - 1) / gridDim.y;
  const int lo = blockIdx.y * per, hi = min(lo + per, n);
  for (int i = lo + warp; i < hi; i += nwarps) {
    const float4 pc = spos[i];
    int cnt = 0;
    for (int w = 0; w < words; w++) {
      const int j = w * 32 + lane;
      bool ok = false;
      if (j < n) {
        ok = pass_pred(pair_dr2<false>(pc, spos[j], p), p);
        if (j == i) ok = (p.flags & GAMD_NBR_SELF) != 0;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) gmask[((size_t)base + i) * words + w] = m;
      cnt += __popc(m);
    }
    if (lane == 0) deg_g[base + i] = cnt;
  }
}

// pass 1: CSR rows from the masks.  SINGLE: one frame - the CTA scans the degrees itself (n <= 1024 = one per thread) and
// publishes row_ptr / n_edges; otherwise row_ptr comes from the global scan.
template <bool SINGLE>
__global__ void __launch_bounds__(SMALL_THREADS) k_nbr_small_fill(NbrParams p, const int* __restrict__ deg_g,
                                                                  int* __restrict__ row_ptr, const uint32_t* __restrict__ gmask,
                                                                  int* __restrict__ col, int* __restrict__ edst, int cap,
                                                                  int* __restrict__ n_edges, int* __restrict__ err_flag) {
  __shared__ int s_off[SMALL_MAX + 1];
  __shared__ int s_warp[33];
  const int n = p.atoms_per_frame, words = (n + 31) >> 5;
  const int base = blockIdx.x * n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  if (SINGLE) {
    const int v = tid < n ? deg_g[tid] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int ws = lane < nwarps ? s_warp[lane] : 0;
      int winc = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      s_warp[lane] = winc - ws;
      if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    const int ex = inc - v + s_warp[warp];
    const int total = s_warp[32];
    if (tid < n) {
      s_off[tid] = ex;
      row_ptr[tid] = ex;
    }
    if (tid == 0) {
      row_ptr[n] = total;
      if (total > cap) {          // publish an empty edge list (see k_nbr_guard)
        atomicOr(err_flag, 1);
        err_flag[1] = total;
        *n_edges = 0;
      } else {
        *n_edges = total;
      }
    }
    __syncthreads();
    if (total > cap) return;
  }
  for (int i = warp; i < n; i += nwarps) {
    int out = SINGLE ? s_off[i] : row_ptr[base + i];
    if (!SINGLE && row_ptr[base + i + 1] > cap) continue;
    for (int w = 0; w < words; w++) {
      const uint32_t m = gmask[((size_t)base + i) * words + w];
      if ((m >> lane) & 1u) {
        const int dst = out + __popc(m & ((1u << lane) - 1u));
        col[dst] = base + w * 32 + lane;
        edst[dst] = base + i;
      }
      out += __popc(m);
    }
  }
}

// engine path for frames of at most SMALL_MAX atoms (dr2 < rc^2, self pairs kept, wrapped positions)
int nbr_small_frames(gamd_ctx* ctx, const double* d_x, double scale, const double* box64, const NbrParams& p,
                     const float* d_feat, cudaStream_t st) {
  const int n = p.atoms_per_frame, words = (n + 31) >> 5;
  const int cap = (int)ctx->cap_edges;
  if ((int64_t)p.n_atoms * words > ctx->vl_cap) {
    ctx->err = "mask scratch too small for the per-frame neighbor search";
    return GAMD_ECAPACITY;
  }
  uint32_t* gm = reinterpret_cast<uint32_t*>(ctx->vl_cand);     // scratch: the skin path is off for small frames
  // a single frame spreads its centres over enough CTAs to fill the machine's latency, many frames take one each
  const int slices = p.n_frames == 1 ? (n + 31) / 32 : (p.n_frames < ctx->sm_count ? 4 : 1);
  k_nbr_small_count<<<dim3(p.n_frames, slices), p.n_frames == 1 ? 1024 : 256, 0, st>>>(
      d_x, scale, box64[0], box64[1], box64[2], p, d_feat, ctx->pos_nbr, ctx->pos_nbr_s, ctx->pos_feat_s, ctx->perm,
      ctx->inv_perm, ctx->deg, gm);
  GAMD_LAUNCH_CHECK();
  if (p.n_frames == 1) {
    k_nbr_small_fill<true><<<1, SMALL_THREADS, 0, st>>>(p, ctx->deg, ctx->row_ptr, gm, ctx->col_idx, ctx->edge_dst, cap,
                                                        ctx->n_edges, ctx->err_flag);
    GAMD_LAUNCH_CHECK();
  } else {
    int rc = scan_with_total(ctx, ctx->deg, ctx->row_ptr, p.n_atoms, ctx->n_edges, st);
    if (rc) return rc;
    k_nbr_guard<<<1, 1, 0, st>>>(ctx->row_ptr, p.n_atoms, cap, ctx->n_edges, ctx->err_flag);
    GAMD_LAUNCH_CHECK();
    k_nbr_small_fill<false><<<p.n_frames, 256, 0, st>>>(p, ctx->deg, ctx->row_ptr, gm, ctx->col_idx, ctx->edge_dst, cap,
                                                        ctx->n_edges, ctx->err_flag);
    GAMD_LAUNCH_CHECK();
  }
  ctx->last_nbr = p;
  ctx->vl_key = 0;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Verlet-skin reuse (the reference: jax-md neighbor_list with dr_threshold = cutoff / 6 rebuilds its candidate
// list only when an atom has moved more than half the skin, and re-applies the exact mask every step,
// code/graph_utils.py:21-25, :36-44, :51-61).  Same here, without a host round trip:
//
//   every step   k_vl_place   positions in the cell order of the LAST rebuild; max displacement since then > 0.45 skin
//                             -> device flag
//   flag set     the cell-list pipeline above, gated on the flag (bin, sort, gather) with cells of edge >= rc + skin,
//                a 27-cell sweep with the predicate dr2 < (rc + skin)^2 -> per-centre candidate rows (padded to 32)
//   every step   k_vl_count   the EXACT predicate on the ~36 candidates of a centre (instead of ~146 atoms of 27 cells),
//                             ballot masks kept; scan; k_vl_fill writes the CSR from the masks (one predicate pass)
//
// The edge SET is that of the exact predicate on the current wrapped fp32 positions - identical to the rebuilt-every-
// step path - because every pair within rc now was within rc + skin at the rebuild (each atom moved < skin / 2).
// ------------------------------------------------------------------------------------------
__global__ void k_vl_place(const double* __restrict__ x, double scale, double bx, double by, double bz, NbrParams p,
                           const int* __restrict__ perm, const float* __restrict__ feat, float4* __restrict__ pos_nbr,
                           float4* __restrict__ pos_nbr_s, float4* __restrict__ pos_feat_s,
                           const float4* __restrict__ pos_ref, float thr2, int* __restrict__ flag) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.n_atoms || *flag) return;     // a forced rebuild (first use, new system) recomputes everything anyway
  const int i = perm[s];
  double px = x[3 * i] * scale, py = x[3 * i + 1] * scale, pz = x[3 * i + 2] * scale;
  float wx = wrap_pos((float)px, p.box[0]), wy = wrap_pos((float)py, p.box[1]), wz = wrap_pos((float)pz, p.box[2]);
  double fx = fmod(px, bx), fy = fmod(py, by), fz = fmod(pz, bz);
  if (fx < 0.0) fx += bx;
  if (fy < 0.0) fy += by;
  if (fz < 0.0) fz += bz;
  pos_nbr_s[s] = make_float4(wx, wy, wz, __int_as_float(i));
  pos_nbr[i] = make_float4(wx, wy, wz, __int_as_float(i));      // caller order (gamd_neighbor_export distances)
  pos_feat_s[s] = make_float4((float)fx, (float)fy, (float)fz, feat ? feat[i] : 0.f);
  const float4 r = pos_ref[s];
  const float dx = min_image<false>(wx - r.x, p.box[0], p.half[0]);
  const float dy = min_image<false>(wy - r.y, p.box[1], p.half[1]);
  const float dz = min_image<false>(wz - r.z, p.box[2], p.half[2]);
  if (dx * dx + dy * dy + dz * dz > thr2) *flag = 1;      // benign race: every writer stores 1
}

__global__ void k_vl_pad32(const int* __restrict__ deg, int n, int* __restrict__ cnt32, const int* __restrict__ gate) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n || !*gate) return;
  cnt32[s] = (deg[s] + 31) & ~31;
}

// candidate rows: the sweep's output compacted at cand_ptr[s]; the padding slots get -1; reference positions saved
__global__ void k_vl_finish_rows(const int* __restrict__ deg, const int* __restrict__ cand_ptr, int n, int cap,
                                 int* __restrict__ cand, const float4* __restrict__ pos_nbr_s,
                                 float4* __restrict__ pos_ref, int* __restrict__ err_flag, const int* __restrict__ gate) {
  const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
  if (!*gate) return;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    if (lane == 0) pos_ref[s] = pos_nbr_s[s];
    const int b = cand_ptr[s], e = cand_ptr[s + 1], d = deg[s];
    if (e > cap) {
      if (lane == 0) atomicOr(err_flag, 4);
      continue;
    }
    for (int k = b + d + lane; k < e; k += 32) cand[k] = -1;
  }
}

__global__ void k_vl_clear(int* flag, unsigned long long* counters) {
  counters[1] += 1ull;                 // steps
  if (*flag) counters[0] += 1ull;      // rebuilds
  *flag = 0;
}

__global__ void __launch_bounds__(256) k_vl_count(NbrParams p, const float4* __restrict__ pos,
                                                  const int* __restrict__ cand_ptr, const int* __restrict__ cand,
                                                  uint32_t* __restrict__ vmask, int* __restrict__ deg) {
  const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < p.n_atoms; s += warps) {
    const float4 pc = pos[s];
    if (__float_as_int(pc.w) >= p.n_centers) {
      if (lane == 0) deg[s] = 0;
      continue;
    }
    const int b = cand_ptr[s], e = cand_ptr[s + 1];
    int cnt = 0;
    for (int k = b; k < e; k += 32) {
      const int j = cand[k + lane];
      bool ok = false;
      if (j >= 0) {
        ok = pass_pred(pair_dr2<false>(pc, pos[j], p), p);
        if (j == s) ok = (p.flags & GAMD_NBR_SELF) != 0;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) vmask[k >> 5] = m;
      cnt += __popc(m);
    }
    if (lane == 0) deg[s] = cnt;
  }
}

__global__ void __launch_bounds__(256) k_vl_fill(int n, const int* __restrict__ cand_ptr, const int* __restrict__ cand,
                                                 const uint32_t* __restrict__ vmask, const int* __restrict__ row_ptr,
                                                 int* __restrict__ col, int* __restrict__ edst, int cap,
                                                 int* __restrict__ err_flag) {
  const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    int out = row_ptr[s];
    if (row_ptr[s + 1] > cap) {
      if (lane == 0 && row_ptr[s + 1] > row_ptr[s]) atomicOr(err_flag, 1);
      continue;
    }
    const int b = cand_ptr[s], e = cand_ptr[s + 1];
    for (int k = b; k < e; k += 32) {
      const uint32_t m = vmask[k >> 5];
      if ((m >> lane) & 1u) {
        const int dst = out + __popc(m & ((1u << lane) - 1u));
        col[dst] = cand[k + lane];
        edst[dst] = s;
      }
      out += __popc(m);
    }
  }
}

// engine path (fp64 state, wrapped positions, dr2 < rc^2, self pairs kept): positions -> CSR with skin reuse
int nbr_step_verlet(gamd_ctx* ctx, const double* d_x, double scale, const double* box64, const NbrParams& p,
                    const float* d_feat, cudaStream_t st) {
  const int n = p.n_atoms;
  const float skin = ctx->vl_skin_frac * p.rc;
  float boxf[3] = {p.box[0], p.box[1], p.box[2]};
  NbrParams pc;                                   // the candidate search: cells and predicate of radius rc + skin
  int rc = nbr_setup_params(ctx, n, p.n_frames, boxf, p.rc + skin, p.flags, &pc);
  if (rc) return rc;
  pc.n_centers = p.n_centers;
  // anything that invalidates the saved cell order / candidate rows forces a rebuild
  uint64_t key = 1469598103934665603ull;
  auto mix = [&](uint64_t v) { key = (key ^ v) * 1099511628211ull; };
  mix((uint64_t)n); mix((uint64_t)p.n_frames); mix((uint64_t)p.flags); mix((uint64_t)ctx->arena); mix((uint64_t)d_feat);
  for (int d = 0; d < 3; d++) { uint32_t b; memcpy(&b, &p.box[d], 4); mix(b); }
  { uint32_t b; memcpy(&b, &p.rc, 4); mix(b); }
  mix(ctx->vl_epoch);
  int* flag = ctx->vl_flag;
  if (key != ctx->vl_key) {
    GAMD_CUDA(cudaMemsetAsync(flag, 0xff, sizeof(int), st));
    ctx->vl_key = key;
  }
  const float thr = 0.45f * skin;
  k_vl_place<<<ceil_div(n, 256), 256, 0, st>>>(d_x, scale, box64[0], box64[1], box64[2], p, ctx->perm, d_feat,
                                               ctx->pos_nbr, ctx->pos_nbr_s, ctx->pos_feat_s, ctx->vl_pos_ref, thr * thr, flag);
  GAMD_LAUNCH_CHECK();
  // ---- rebuild, gated on the flag ----
  k_bin_f64<<<ceil_div(n, 256), 256, 0, st>>>(d_x, scale, box64[0], box64[1], box64[2], pc, ctx->pos_nbr, ctx->pos_feat,
                                              ctx->keys[0], ctx->vals[0], flag);
  GAMD_LAUNCH_CHECK();
  const int64_t ncells = (int64_t)pc.cells_per_frame * pc.n_frames;
  int bits = 1;
  while (((int64_t)1 << bits) < ncells) bits++;
  int buf = 0;
  if (ncells > 1 && (rc = radix_sort_pairs(ctx, n, bits, st, &buf, flag))) return rc;
  k_gather_sorted<<<ceil_div(n, 256), 256, 0, st>>>(ctx->keys[buf], ctx->vals[buf], n, (int)ncells, ctx->pos_nbr,
                                                    ctx->pos_feat, d_feat, ctx->pos_nbr_s, ctx->pos_feat_s, ctx->perm,
                                                    ctx->inv_perm, ctx->cell_start, flag);
  GAMD_LAUNCH_CHECK();
  const int blocks = std::min(ceil_div((int64_t)n * 32, 256), ctx->sm_count * 32);
  const int cap_c = (int)ctx->vl_cap;
  k_sweep<false, false><<<blocks, 256, 0, st>>>(pc, ctx->pos_nbr_s, ctx->keys[buf], ctx->cell_start, ctx->deg, nullptr,
                                                nullptr, nullptr, cap_c, ctx->err_flag, flag);
  GAMD_LAUNCH_CHECK();
  k_vl_pad32<<<ceil_div(n, 256), 256, 0, st>>>(ctx->deg, n, ctx->vl_cnt, flag);
  GAMD_LAUNCH_CHECK();
  if ((rc = scan_with_total(ctx, ctx->vl_cnt, ctx->vl_ptr, n, nullptr, st, flag))) return rc;
  k_sweep<true, false><<<blocks, 256, 0, st>>>(pc, ctx->pos_nbr_s, ctx->keys[buf], ctx->cell_start, ctx->deg, ctx->vl_ptr,
                                               ctx->vl_cand, nullptr, cap_c, ctx->err_flag + 2, flag);
  GAMD_LAUNCH_CHECK();
  k_vl_finish_rows<<<blocks, 256, 0, st>>>(ctx->deg, ctx->vl_ptr, n, cap_c, ctx->vl_cand, ctx->pos_nbr_s, ctx->vl_pos_ref,
                                           ctx->err_flag, flag);
  GAMD_LAUNCH_CHECK();
  k_vl_clear<<<1, 1, 0, st>>>(flag, ctx->vl_counters);
  GAMD_LAUNCH_CHECK();
  // ---- every step: the exact predicate on the candidates ----
  k_vl_count<<<blocks, 256, 0, st>>>(p, ctx->pos_nbr_s, ctx->vl_ptr, ctx->vl_cand, ctx->vl_mask, ctx->deg);
  GAMD_LAUNCH_CHECK();
  if ((rc = scan_with_total(ctx, ctx->deg, ctx->row_ptr, n, ctx->n_edges, st))) return rc;
  const int cap = (int)ctx->cap_edges;
  k_nbr_guard<<<1, 1, 0, st>>>(ctx->row_ptr, n, cap, ctx->n_edges, ctx->err_flag);
  GAMD_LAUNCH_CHECK();
  k_vl_fill<<<blocks, 256, 0, st>>>(n, ctx->vl_ptr, ctx->vl_cand, ctx->vl_mask, ctx->row_ptr, ctx->col_idx, ctx->edge_dst,
                                    cap, ctx->err_flag);
  GAMD_LAUNCH_CHECK();
  ctx->last_nbr = p;
  return 0;
}

// ------------------------------------------------------------------------------------------
// export: CSR in sorted space -> COO in caller ids (centre-major, neighbour ascending)
// ------------------------------------------------------------------------------------------
__global__ void k_deg_to_orig(const int* __restrict__ deg, const int* __restrict__ perm, int n, int* __restrict__ deg_o) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) deg_o[perm[s]] = deg[s];
}

__global__ void k_export_rows(NbrParams p, const int* __restrict__ row_ptr, const int* __restrict__ col,
                              const int* __restrict__ perm, const int* __restrict__ row_ptr_o,
                              const float4* __restrict__ pos_orig, int64_t* __restrict__ out, int64_t cap,
                              float* __restrict__ dist, float* __restrict__ norm) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.n_atoms) return;
  int i = perm[s];
  int b = row_ptr_o[i], d = row_ptr[s + 1] - row_ptr[s];
  if ((int64_t)b + d > cap) return;
  int64_t* nb = out + cap + b;
  for (int k = 0; k < d; k++) {  // insertion sort by caller id
    int64_t j = perm[col[row_ptr[s] + k]];
    int q = k;
    while (q > 0 && nb[q - 1] > j) {
      nb[q] = nb[q - 1];
      q--;
    }
    nb[q] = j;
  }
  float4 pc = pos_orig[i];
  for (int k = 0; k < d; k++) {
    out[b + k] = i;
    if (dist || norm) {
      float4 pn = pos_orig[nb[k]];
      float tx = min_image<true>(__fsub_rn(pc.x, pn.x), p.box[0], p.half[0]);
      float ty = min_image<true>(__fsub_rn(pc.y, pn.y), p.box[1], p.half[1]);
      float tz = min_image<true>(__fsub_rn(pc.z, pn.z), p.box[2], p.half[2]);
      if (dist) {
        dist[3 * (b + k)] = tx;
        dist[3 * (b + k) + 1] = ty;
        dist[3 * (b + k) + 2] = tz;
      }
      if (norm) norm[b + k] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)), __fmul_rn(tz, tz)));
    }
  }
}

int nbr_export(gamd_ctx* ctx, int64_t* d_edge_idx, int64_t cap, float* d_dist, float* d_norm, cudaStream_t st) {
  const NbrParams& p = ctx->last_nbr;
  int n = p.n_atoms;
  if (n <= 0) {
    ctx->err = "gamd_neighbor_export before gamd_neighbor_build";
    return GAMD_ESTATE;
  }
  k_deg_to_orig<<<ceil_div(n, 256), 256, 0, st>>>(ctx->deg, ctx->perm, n, ctx->deg_o);
  GAMD_LAUNCH_CHECK();
  int rc = exclusive_scan_i32(ctx, ctx->deg_o, ctx->row_ptr_o, n, st);
  if (rc) return rc;
  k_export_rows<<<ceil_div(n, 128), 128, 0, st>>>(p, ctx->row_ptr, ctx->col_idx, ctx->perm, ctx->row_ptr_o,
                                                  ctx->pos_nbr, d_edge_idx, cap, d_dist, d_norm);
  GAMD_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// explicit edge list (sorted by centre) -> CSR, caller index space
// ------------------------------------------------------------------------------------------
__global__ void k_coo_to_csr(const int64_t* __restrict__ center, const int64_t* __restrict__ neigh, int64_t n_edges,
                             int n_atoms, int* __restrict__ row_ptr, int* __restrict__ col, int* __restrict__ edst,
                             int* __restrict__ n_edges_dev, int* __restrict__ err_flag) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) *n_edges_dev = (int)n_edges;
  if (e >= n_edges) return;
  int c = (int)center[e];
  int cp = e > 0 ? (int)center[e - 1] : -1;
  if (c < cp || c < 0 || c >= n_atoms || neigh[e] < 0 || neigh[e] >= n_atoms) {
    atomicOr(err_flag, 2);  // not sorted by centre / id out of range
    return;
  }
  col[e] = (int)neigh[e];
  edst[e] = c;
  for (int r = cp + 1; r <= c; r++) row_ptr[r] = (int)e;
  if (e == n_edges - 1)
    for (int r = c + 1; r <= n_atoms; r++) row_ptr[r] = (int)n_edges;
}

int csr_from_sorted_coo(gamd_ctx* ctx, const int64_t* d_center, const int64_t* d_neigh, int64_t n_atoms,
                        int64_t n_edges, cudaStream_t st) {
  if (n_edges > ctx->cap_edges || n_atoms > ctx->cap_atoms) {
    ctx->err = "edge list exceeds reserved capacity; call gamd_reserve";
    return GAMD_ECAPACITY;
  }
  if (n_edges == 0) {
    GAMD_CUDA(cudaMemsetAsync(ctx->row_ptr, 0, sizeof(int) * (n_atoms + 1), st));
    GAMD_CUDA(cudaMemsetAsync(ctx->n_edges, 0, sizeof(int), st));
    return 0;
  }
  k_coo_to_csr<<<ceil_div(n_edges, 256), 256, 0, st>>>(d_center, d_neigh, n_edges, (int)n_atoms, ctx->row_ptr,
                                                       ctx->col_idx, ctx->edge_dst, ctx->n_edges, ctx->err_flag);
  GAMD_LAUNCH_CHECK();
  return 0;
}
