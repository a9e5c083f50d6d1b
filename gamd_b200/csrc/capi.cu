// C ABI of libgamd_b200 (see include/gamd_b200.h for the contract of every entry point).
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>

static std::string g_create_err;

namespace {

struct WSpec {
  std::string name;
  int64_t rows, cols;   // cols == 0 -> vector of `rows`
};

std::vector<WSpec> expected_weights(const gamd_model_desc& d) {
  std::vector<WSpec> v;
  const int D = d.encoding_size, H = d.hidden_dim, De = d.edge_dim;
  const int n_in = 4 + (d.expand_edge ? GAMD_NRBF : 0) + (d.use_bond ? 1 : 0);
  v.push_back({"length_mean", 1, 0});
  v.push_back({"length_std", 1, 0});
  if (d.kind == GAMD_MODEL_LJ) v.push_back({"node_emb", 1, D});
  for (int l = 0; l < d.conv_layer; l++) {
    std::string p = "graph_conv.conv." + std::to_string(l) + ".";
    auto lin = [&](const std::string& n, int64_t out, int64_t in) {
      v.push_back({p + n + ".weight", out, in});
      v.push_back({p + n + ".bias", out, 0});
    };
    if (d.update_edge) {
      v.push_back({p + "edge_layer_norm.weight", De, 0});
      v.push_back({p + "edge_layer_norm.bias", De, 0});
    }
    lin("edge_affine.mlp_layer.0", 128, De);
    lin("edge_affine.mlp_layer.2", H, 128);
    lin("src_affine", H, D);
    lin("dst_affine", H, D);
    lin("theta_edge.mlp_layer.1", H, H);
    lin("theta_edge.mlp_layer.3", D, H);
    lin("phi_dst", H, D);
    lin("phi_edge", H, D);
    lin("phi.mlp_layer.1", D, H);
  }
  for (int l = 0; l < d.conv_layer; l++) {
    v.push_back({"graph_conv.norm_layers." + std::to_string(l) + ".weight", D, 0});
    v.push_back({"graph_conv.norm_layers." + std::to_string(l) + ".bias", D, 0});
    if (d.batch_norm) {
      v.push_back({"graph_conv.norm_layers." + std::to_string(l) + ".running_mean", D, 0});
      v.push_back({"graph_conv.norm_layers." + std::to_string(l) + ".running_var", D, 0});
      v.push_back({"graph_conv.norm_layers." + std::to_string(l) + ".num_batches_tracked", 1, 0});
    }
  }
  if (d.expand_edge) v.push_back({"edge_expand.centers", GAMD_NRBF, 0});
  if (d.kind != GAMD_MODEL_LJ) {
    v.push_back({"node_encoder.weight", D, d.in_feats});
    v.push_back({"node_encoder.bias", D, 0});
  }
  v.push_back({"edge_encoder.mlp_layer.0.weight", H, n_in});
  v.push_back({"edge_encoder.mlp_layer.0.bias", H, 0});
  v.push_back({"edge_encoder.mlp_layer.2.weight", H, H});
  v.push_back({"edge_encoder.mlp_layer.2.bias", H, 0});
  v.push_back({"edge_encoder.mlp_layer.4.weight", De, H});
  v.push_back({"edge_encoder.mlp_layer.4.bias", De, 0});
  v.push_back({"edge_layer_norm.weight", De, 0});
  v.push_back({"edge_layer_norm.bias", De, 0});
  v.push_back({"graph_decoder.mlp_layer.0.weight", H, D});
  v.push_back({"graph_decoder.mlp_layer.0.bias", H, 0});
  v.push_back({"graph_decoder.mlp_layer.2.weight", 3, H});
  v.push_back({"graph_decoder.mlp_layer.2.bias", 3, 0});
  return v;
}

// bump allocator over the arena, 256-byte aligned
struct Carver {
  char* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

void carve(gamd_ctx* ctx, Carver& c, int64_t A, int64_t E) {
  const size_t D = ctx->desc.encoding_size, H = ctx->desc.hidden_dim, De = ctx->desc.edge_dim;
  const int64_t rs_blocks = (A + 2047) / 2048 + 1;
  ctx->cap_cells = 8 * A + 1024;
  for (int i = 0; i < 2; i++) {
    ctx->keys[i] = c.take<uint32_t>(A);
    ctx->vals[i] = c.take<uint32_t>(A);
  }
  ctx->radix_hist = c.take<uint32_t>(2 * (256 * rs_blocks + 16));
  ctx->scan_tmp = c.take<uint32_t>((A + 1) / 1024 + rs_blocks + 4096);
  ctx->pos_nbr = c.take<float4>(A);
  ctx->pos_feat = c.take<float4>(A);
  ctx->pos_nbr_s = c.take<float4>(A);
  ctx->pos_feat_s = c.take<float4>(A);
  ctx->perm = c.take<int>(A);
  ctx->inv_perm = c.take<int>(A);
  ctx->cell_start = c.take<int>(ctx->cap_cells + 1);
  ctx->deg = c.take<int>(A + 1);
  ctx->row_ptr = c.take<int>(A + 1);
  ctx->deg_o = c.take<int>(A + 1);
  ctx->row_ptr_o = c.take<int>(A + 1);
  ctx->n_edges = c.take<int>(4);
  ctx->err_flag = c.take<int>(4);
  ctx->d_stepctr = c.take<int>(4);
  ctx->col_idx = c.take<int>(E);
  ctx->edge_dst = c.take<int>(E);
  ctx->e_emb = c.take<float>((size_t)(E + 256) * De);   // fp32 rows, or 64 KB bf16 hi/lo blobs per 128-edge tile
  ctx->h = c.take<float>((size_t)A * D);
  ctx->hn = c.take<float>((size_t)A * D);
  ctx->srcA = c.take<float>((size_t)A * H);
  ctx->dstA = c.take<float>((size_t)A * H);
  ctx->pd = c.take<float>((size_t)A * H);
  ctx->agg = c.take<float>((size_t)A * D);
  // partial sums of receiver runs cut by an edge tile: 32-edge blocks (tensor path), or the generic path's 16 r-row tiles
  ctx->part = c.take<float>(ctx->wide ? (size_t)(E / (16 * ctx->wide_r) + 4) * 2 * D : (size_t)(E / 32 + 4) * 2 * GAMD_NF);
  ctx->tile_list[0] = c.take<int>(E / 128 + 4);
  ctx->tile_list[1] = c.take<int>(E / 128 + 4);
  ctx->tile_count = c.take<int>(4);
  ctx->pred = c.take<float>((size_t)A * 3);
  ctx->feat_s = c.take<float>(A);
  ctx->stage_a = c.take<double>((size_t)A * 3);
  ctx->stage_b = c.take<double>((size_t)A * 3);
  ctx->stage_c = c.take<double>((size_t)A * 3);
  ctx->stage_m = c.take<double>(A);
  ctx->stage_feat = c.take<float>(A);
  // Verlet-skin candidate rows: (rc + skin)^3 / rc^3 = 1.59 x the edges for skin = rc / 6, rows padded to 32
  ctx->vl_cap = 2 * E + 32 * A;
  ctx->vl_ptr = c.take<int>(A + 1);
  ctx->vl_cnt = c.take<int>(A + 1);
  ctx->vl_cand = c.take<int>(ctx->vl_cap);
  ctx->vl_mask = c.take<uint32_t>(ctx->vl_cap / 32 + 4);
  ctx->vl_pos_ref = c.take<float4>(A);
  ctx->vl_flag = c.take<int>(4);
  ctx->vl_counters = c.take<unsigned long long>(2);
}

__global__ void k_pack_pos_feat(const float* __restrict__ pos, const float* __restrict__ feat, int64_t n,
                                float4* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], feat ? feat[i] : 0.f);
}

int check_ready(gamd_ctx* ctx) {
  if (!ctx) return GAMD_EINVAL;
  if (!ctx->finalized) {
    ctx->err = "weights not finalized: call gamd_load_weight for every tensor, then gamd_finalize_weights";
    return GAMD_ESTATE;
  }
  if (!ctx->arena) {
    ctx->err = "no scratch reserved: call gamd_reserve";
    return GAMD_ESTATE;
  }
  GAMD_ENTER(ctx);
  return 0;
}

// the bond table is indexed by frame-local atom id: it must cover exactly one frame
int check_bonds(gamd_ctx* ctx, int64_t n_atoms, int n_frames) {
  if (ctx->desc.use_bond && n_frames > 0 && n_atoms / n_frames != ctx->bond_atoms) {
    ctx->err = "bond table was built for " + std::to_string(ctx->bond_atoms) + " atoms per frame, got " +
               std::to_string(n_atoms / n_frames) + ": call gamd_set_bonds for this system";
    return GAMD_EINVAL;
  }
  return 0;
}

int positions_to_forces(gamd_ctx* ctx, const double* d_x, double scale, int64_t n, int n_frames, const double box[3],
                        float cutoff, const float* d_feat, cudaStream_t st) {
  float boxf[3] = {(float)box[0], (float)box[1], (float)box[2]};
  NbrParams p;
  int rc = check_bonds(ctx, n, n_frames);
  if (rc) return rc;
  if ((rc = nbr_setup_params(ctx, n, n_frames, boxf, cutoff, GAMD_NBR_LT | GAMD_NBR_SELF, &p))) return rc;
  prof_mark(ctx, "neighbor", st);
  if (ctx->small_frames && p.atoms_per_frame <= 1024) {
    // launch-bound systems: the whole search in one CTA per frame (one launch for a single frame)
    if ((rc = nbr_small_frames(ctx, d_x, scale, box, p, d_feat, st))) return rc;
  } else if (ctx->vl_skin_frac > 0.f && n >= ctx->vl_min_atoms && ctx->vl_cap < (int64_t(1) << 31)) {
    // candidate list with a skin, rebuilt only when an atom moved more than 0.45 skin (graph_utils.py:21-25)
    if ((rc = nbr_step_verlet(ctx, d_x, scale, box, p, d_feat, st))) return rc;
  } else {
    if ((rc = nbr_bin_f64(ctx, d_x, scale, box, p, st))) return rc;
    if ((rc = nbr_sort_and_sweep(ctx, p, d_feat, st))) return rc;
    ctx->vl_key = 0;          // the cell order changed under the saved candidate rows
  }
  prof_mark(ctx, "neighbor", st);
  return model_forward(ctx, ctx->pos_feat_s, nullptr, ctx->perm, n, p.atoms_per_frame, boxf, st);
}

}  // namespace

void prof_mark(gamd_ctx* ctx, const char* stage, cudaStream_t st) {
  if (!ctx->prof_on) return;
  auto& p = ctx->prof[stage];
  if (p.used == p.ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    p.ev.push_back(e);
  }
  cudaEventRecord(p.ev[p.used++], st);
}

__global__ void k_dd_pack(const int* __restrict__ idx, const int* __restrict__ inv_perm, int64_t n,
                          const float* __restrict__ hn, const float* __restrict__ srcA, float* __restrict__ out) {
  // one warp per row: out[k] = [hn row | srcA row] of local atom idx[k]
  int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (k >= n) return;
  int s = inv_perm[idx[k]];
  float4 a = reinterpret_cast<const float4*>(hn + (size_t)s * 128)[lane];
  float4 b = reinterpret_cast<const float4*>(srcA + (size_t)s * 128)[lane];
  reinterpret_cast<float4*>(out + (size_t)k * 256)[lane] = a;
  reinterpret_cast<float4*>(out + (size_t)k * 256 + 128)[lane] = b;
}

__global__ void k_dd_unpack(int64_t first, const int* __restrict__ inv_perm, int64_t n, const float* __restrict__ in,
                            float* __restrict__ hn, float* __restrict__ srcA) {
  int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (k >= n) return;
  int s = inv_perm[first + k];
  reinterpret_cast<float4*>(hn + (size_t)s * 128)[lane] = reinterpret_cast<const float4*>(in + (size_t)k * 256)[lane];
  reinterpret_cast<float4*>(srcA + (size_t)s * 128)[lane] = reinterpret_cast<const float4*>(in + (size_t)k * 256 + 128)[lane];
}

// ---- halo exchange over NVLink peer memory (CUDA IPC): the pack kernel writes straight into the neighbour's buffer ----
__global__ void k_dd_push_bytes(const float4* __restrict__ src, float4* __restrict__ dst, int64_t n16) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// after the pushing kernel (stream order): make its peer writes visible system-wide, then publish the sequence number
__global__ void k_dd_signal(unsigned long long* remote_flag, unsigned long long seq) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote_flag), "l"(seq) : "memory");
}
__global__ void k_dd_wait(const unsigned long long* flag, unsigned long long seq, int* err_flag) {
  unsigned long long v = 0;
  for (unsigned long long it = 0; it < (1ull << 31); it++) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= seq) return;
    __nanosleep(64);
  }
  atomicOr(err_flag, 8);       // the neighbour never delivered (bounded wait instead of a hang)
}

int pack_pos_feat(gamd_ctx* ctx, const float* d_pos, const float* d_feat, int64_t n, float4* out, cudaStream_t st) {
  k_pack_pos_feat<<<ceil_div(n, 256), 256, 0, st>>>(d_pos, d_feat, n, out);
  GAMD_LAUNCH_CHECK();
  return 0;
}

extern "C" {

const char* gamd_version(void) { return "gamd_b200 0.1 (sm_100a)"; }

const char* gamd_last_error(const gamd_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int gamd_create(int device, const gamd_model_desc* desc, gamd_ctx** out) {
  if (!desc || !out) {
    g_create_err = "null argument";
    return GAMD_EINVAL;
  }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    g_create_err = std::string("no usable CUDA device (there is no CPU fallback): ") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
    return GAMD_ENOGPU;
  }
  // GAMD_FORCE_WIDE=r (development): run a 128-wide model on the generic-width fp32 kernels with tiles of 16 r rows
  const int force_wide = getenv("GAMD_FORCE_WIDE") ? atoi(getenv("GAMD_FORCE_WIDE")) : 0;
  const bool wide = desc->encoding_size != GAMD_NF || desc->hidden_dim != GAMD_NF || desc->edge_dim != GAMD_NF ||
                    desc->update_edge || desc->batch_norm || !desc->expand_edge || force_wide > 0;
  int wide_r = 4, wide_xs = GAMD_NF + 4;
  if (wide) {
    for (int v : {desc->encoding_size, desc->hidden_dim, desc->edge_dim})
      if (v < 128 || v > 1024 || v % 128) {
        g_create_err = "encoding_size, hidden_dim and edge_embedding_dim must be multiples of 128 in 128..1024";
        return GAMD_EUNSUPPORTED;
      }
    if (desc->update_edge && desc->edge_dim != desc->encoding_size) {
      g_create_err = "update_edge needs edge_embedding_dim == encoding_size (edge_layer_norm acts on e_emb)";
      return GAMD_EINVAL;
    }
    if (wide_plan(desc->encoding_size, desc->hidden_dim, desc->edge_dim, &wide_r, &wide_xs)) {
      g_create_err = "model too wide for the shared-memory tiles";
      return GAMD_EUNSUPPORTED;
    }
    if (force_wide == 1 || force_wide == 2 || force_wide == 4) wide_r = force_wide < wide_r ? force_wide : wide_r;
  }
  if (desc->conv_layer < 1 || desc->conv_layer > 8) {
    g_create_err = "conv_layer must be in 1..8";
    return GAMD_EUNSUPPORTED;
  }
  if (desc->kind != GAMD_MODEL_LJ && desc->in_feats != 1) {
    g_create_err = "node_encoder in_feats must be 1";
    return GAMD_EUNSUPPORTED;
  }
  if (desc->precision != GAMD_PREC_FP32 && desc->precision != GAMD_PREC_BF16X3 && desc->precision != GAMD_PREC_BF16) {
    g_create_err = "unknown precision mode";
    return GAMD_EINVAL;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    return GAMD_ECUDA;
  }
  gamd_ctx* ctx = new gamd_ctx();
  ctx->device = device;
  ctx->desc = *desc;
  ctx->wide = wide;
  ctx->wide_r = wide_r;
  ctx->wide_xs = wide_xs;
  if (wide) ctx->desc.precision = GAMD_PREC_FP32;   // the generic-width path computes in fp32 whatever was asked for
  ctx->use_graphs = getenv("GAMD_NO_GRAPH") == nullptr;
  ctx->dbg_timeline = getenv("GAMD_TIMELINE") != nullptr;
  if (const char* e = getenv("GAMD_WAIT_HINT_NS")) ctx->wait_hint_ns = atoi(e);
  // GAMD_WAIT_SLEEP_NS: poll the accumulator barrier with __nanosleep(ns) between attempts instead
  if (const char* e = getenv("GAMD_WAIT_SLEEP_NS")) ctx->wait_hint_ns = (int)(0x80000000u | (uint32_t)(atoi(e) & 0xffff));
  if (const char* e = getenv("GAMD_MP_ROW_PREFETCH")) ctx->mp_row_prefetch = atoi(e);
  ctx->dd_reserve_sms = getenv("GAMD_DD_RESERVE_SMS") ? atoi(getenv("GAMD_DD_RESERVE_SMS")) : 0;
  // message-passing edge kernel: CTA pairs (cta_group::2), resident weights, three tiles in flight - 11 = fixed service
  // order in the leader, accumulator block known to the epilogue threads without a published hand-over (default);
  // 8 = dynamic order, single-round-trip tile set-up; 9 / 10 = 8 with the SiLU exponential on the FMA pipe; 6 = 8 with
  // the two-round-trip set-up; 7 = N-split GEMMs; 5 = 6 with a commit wait between GEMMs; 3 / 4 = three tiles, single
  // CTA; 0 = the round-1 two-tile kernel
  ctx->mp_variant = getenv("GAMD_MP_VARIANT") ? atoi(getenv("GAMD_MP_VARIANT")) : 11;
  // neighbor candidate reuse: skin as a fraction of the cutoff (reference: 1/6); GAMD_NBR_SKIN=0 rebuilds every step
  ctx->vl_skin_frac = getenv("GAMD_NBR_SKIN") ? (float)atof(getenv("GAMD_NBR_SKIN")) : (1.f / 6.f);
  if (getenv("GAMD_NBR_SKIN_MIN_ATOMS")) ctx->vl_min_atoms = atoll(getenv("GAMD_NBR_SKIN_MIN_ATOMS"));
  ctx->small_frames = !(getenv("GAMD_NBR_SMALL") && atoi(getenv("GAMD_NBR_SMALL")) == 0);
  if (getenv("GAMD_MP_SMALL_ATOMS")) ctx->mp_small_atoms = atoi(getenv("GAMD_MP_SMALL_ATOMS"));
  if (getenv("GAMD_ENC_VARIANT")) ctx->enc_variant = atoi(getenv("GAMD_ENC_VARIANT"));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  if (ctx->sm_count < 2 && ctx->mp_variant >= 5) ctx->mp_variant = 0;
  *out = ctx;
  return 0;
}

int gamd_destroy(gamd_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->d_wblob) cudaFree(ctx->d_wblob);
  if (ctx->d_wimg) cudaFree(ctx->d_wimg);
  if (ctx->d_wimg2) cudaFree(ctx->d_wimg2);
  if (ctx->d_tc_bias) cudaFree(ctx->d_tc_bias);
  if (ctx->d_wimg_enc) cudaFree(ctx->d_wimg_enc);
  if (ctx->d_tc_bias_enc) cudaFree(ctx->d_tc_bias_enc);
  if (ctx->d_wimg_node) cudaFree(ctx->d_wimg_node);
  if (ctx->d_bond) cudaFree(ctx->d_bond);
  if (ctx->d_nhc) cudaFree(ctx->d_nhc);
  for (void* p : ctx->peer_opened) cudaIpcCloseMemHandle(p);
  for (void* p : ctx->peer_allocs) cudaFree(p);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
  if (ctx->graph_stream) cudaStreamDestroy(ctx->graph_stream);
  if (ctx->graph_ev_in) cudaEventDestroy(ctx->graph_ev_in);
  if (ctx->graph_ev_out) cudaEventDestroy(ctx->graph_ev_out);
  delete ctx;
  return 0;
}

int gamd_reserve(gamd_ctx* ctx, int64_t max_atoms, int64_t max_edges) {
  if (!ctx || max_atoms <= 0 || max_edges < 0) return GAMD_EINVAL;
  if (max_atoms >= (int64_t(1) << 30) || max_edges >= (int64_t(1) << 31) - 64) {
    ctx->err = "capacity exceeds 32-bit indexing";
    return GAMD_EUNSUPPORTED;
  }
  GAMD_CUDA(cudaSetDevice(ctx->device));
  if (max_atoms <= ctx->cap_atoms && max_edges <= ctx->cap_edges) return 0;
  if (max_atoms < ctx->cap_atoms) max_atoms = ctx->cap_atoms;
  if (max_edges < ctx->cap_edges) max_edges = ctx->cap_edges;
  Carver dry{nullptr};
  carve(ctx, dry, max_atoms, max_edges);
  size_t bytes = dry.off + 256;
  GAMD_CUDA(cudaDeviceSynchronize());
  if (ctx->arena) GAMD_CUDA(cudaFree(ctx->arena));
  ctx->arena = nullptr;
  ctx->cap_atoms = ctx->cap_edges = 0;
  cudaError_t e = cudaMalloc(&ctx->arena, bytes);
  if (e != cudaSuccess) {
    ctx->arena = nullptr;
    ctx->err = "cudaMalloc of " + std::to_string(bytes) + " bytes of scratch failed: " + cudaGetErrorString(e);
    return GAMD_ECUDA;
  }
  ctx->arena_bytes = bytes;
  Carver real{static_cast<char*>(ctx->arena)};
  carve(ctx, real, max_atoms, max_edges);
  ctx->cap_atoms = max_atoms;
  ctx->cap_edges = max_edges;
  GAMD_CUDA(cudaMemset(ctx->err_flag, 0, 4 * sizeof(int)));
  GAMD_CUDA(cudaMemset(ctx->n_edges, 0, 4 * sizeof(int)));
  size_t pin = sizeof(double) * 3 * (size_t)max_atoms;
  if (pin > ctx->pinned_bytes) {
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
  }
  ctx->last_nbr = NbrParams{};
  ctx->graph_key = 0;
  ctx->vl_key = 0;
  GAMD_CUDA(cudaMemset(ctx->vl_counters, 0, 2 * sizeof(unsigned long long)));
  return 0;
}

int gamd_load_weight(gamd_ctx* ctx, const char* name, const float* h_data, int64_t n) {
  if (!ctx || !name || !h_data || n <= 0) return GAMD_EINVAL;
  for (const auto& w : expected_weights(ctx->desc)) {
    if (w.name == name) {
      int64_t want = w.rows * (w.cols ? w.cols : 1);
      if (want != n) {
        ctx->err = std::string("size mismatch for ") + name + ": got " + std::to_string(n) + ", expected " +
                   std::to_string(want);
        return GAMD_EINVAL;
      }
      ctx->host_w[name] = std::vector<float>(h_data, h_data + n);
      ctx->finalized = false;
      return 0;
    }
  }
  ctx->err = std::string("unexpected state-dict key: ") + name;
  return GAMD_EINVAL;
}

int gamd_set_scaler(gamd_ctx* ctx, double mean, double var) {
  if (!ctx || !(var >= 0.0)) return GAMD_EINVAL;
  ctx->scaler_mean = mean;
  ctx->scaler_var = var;
  ctx->graph_key = 0;   // the scaler is baked into captured kernel arguments
  return 0;
}

int gamd_set_bonds(gamd_ctx* ctx, const int64_t* h_bonds, int64_t nb, int64_t n_atoms_per_frame) {
  if (!ctx || nb < 0 || n_atoms_per_frame <= 0 || (nb > 0 && !h_bonds)) return GAMD_EINVAL;
  GAMD_CUDA(cudaSetDevice(ctx->device));
  std::vector<int> tab((size_t)n_atoms_per_frame * GAMD_MAX_BOND, -1);
  auto add = [&](int64_t a, int64_t b) -> bool {
    for (int k = 0; k < GAMD_MAX_BOND; k++) {
      int& slot = tab[(size_t)a * GAMD_MAX_BOND + k];
      if (slot == (int)b) return true;
      if (slot < 0) {
        slot = (int)b;
        return true;
      }
    }
    return false;
  };
  for (int64_t i = 0; i < nb; i++) {
    int64_t a = h_bonds[2 * i], b = h_bonds[2 * i + 1];
    if (a < 0 || b < 0 || a >= n_atoms_per_frame || b >= n_atoms_per_frame) {
      ctx->err = "bond index out of range";
      return GAMD_EINVAL;
    }
    if (!add(a, b) || !add(b, a)) {
      ctx->err = "more than 4 bonded partners per atom are not supported";
      return GAMD_EUNSUPPORTED;
    }
  }
  if (ctx->d_bond) cudaFree(ctx->d_bond);
  ctx->d_bond = nullptr;
  GAMD_CUDA(cudaMalloc(&ctx->d_bond, tab.size() * sizeof(int)));
  GAMD_CUDA(cudaMemcpy(ctx->d_bond, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
  ctx->bond_atoms = n_atoms_per_frame;
  ctx->graph_key = 0;
  return 0;
}

int gamd_finalize_weights(gamd_ctx* ctx) {
  if (!ctx) return GAMD_EINVAL;
  GAMD_CUDA(cudaSetDevice(ctx->device));
  const gamd_model_desc& d = ctx->desc;
  auto specs = expected_weights(d);
  for (const auto& w : specs)
    if (!ctx->host_w.count(w.name)) {
      ctx->err = "missing state-dict key: " + w.name;
      return GAMD_ESTATE;
    }
  std::vector<float> blob;
  std::vector<std::pair<const float**, size_t>> fix;  // pointer slots to patch with blob offsets
  auto push = [&](const float** slot, const std::vector<float>& v) {
    while (blob.size() % 64) blob.push_back(0.f);  // 256-byte alignment for cp.async / float4
    fix.push_back({slot, blob.size()});
    blob.insert(blob.end(), v.begin(), v.end());
  };
  auto transposed = [&](const std::string& name, int64_t out, int64_t in, int64_t pad_in) {
    const std::vector<float>& w = ctx->host_w[name];
    std::vector<float> t((size_t)pad_in * out, 0.f);
    for (int64_t o = 0; o < out; o++)
      for (int64_t k = 0; k < in; k++) t[(size_t)k * out + o] = w[(size_t)o * in + k];
    return t;
  };
  ModelW& mw = ctx->mw;
  mw = ModelW{};
  const int D = d.encoding_size, H = d.hidden_dim, De = d.edge_dim;
  for (int l = 0; l < d.conv_layer; l++) {
    std::string p = "graph_conv.conv." + std::to_string(l) + ".";
    std::string nl = "graph_conv.norm_layers." + std::to_string(l) + ".";
    LayerW& L = mw.layer[l];
    push(&L.ea0_t, transposed(p + "edge_affine.mlp_layer.0.weight", 128, De, De));
    push(&L.ea0_b, ctx->host_w[p + "edge_affine.mlp_layer.0.bias"]);
    push(&L.ea2_t, transposed(p + "edge_affine.mlp_layer.2.weight", H, 128, 128));
    push(&L.ea2_b, ctx->host_w[p + "edge_affine.mlp_layer.2.bias"]);
    push(&L.src_t, transposed(p + "src_affine.weight", H, D, D));
    push(&L.src_b, ctx->host_w[p + "src_affine.bias"]);
    push(&L.dst_t, transposed(p + "dst_affine.weight", H, D, D));
    push(&L.dst_b, ctx->host_w[p + "dst_affine.bias"]);
    push(&L.te1_t, transposed(p + "theta_edge.mlp_layer.1.weight", H, H, H));
    push(&L.te1_b, ctx->host_w[p + "theta_edge.mlp_layer.1.bias"]);
    push(&L.te3_t, transposed(p + "theta_edge.mlp_layer.3.weight", D, H, H));
    push(&L.te3_b, ctx->host_w[p + "theta_edge.mlp_layer.3.bias"]);
    push(&L.pdst_t, transposed(p + "phi_dst.weight", H, D, D));
    push(&L.pdst_b, ctx->host_w[p + "phi_dst.bias"]);
    push(&L.pedge_t, transposed(p + "phi_edge.weight", H, D, D));
    push(&L.pedge_b, ctx->host_w[p + "phi_edge.bias"]);
    push(&L.phi_t, transposed(p + "phi.mlp_layer.1.weight", D, H, H));
    push(&L.phi_b, ctx->host_w[p + "phi.mlp_layer.1.bias"]);
    push(&L.ln_w, ctx->host_w[nl + "weight"]);
    push(&L.ln_b, ctx->host_w[nl + "bias"]);
    if (d.batch_norm) {
      push(&L.bn_mean, ctx->host_w[nl + "running_mean"]);
      push(&L.bn_var, ctx->host_w[nl + "running_var"]);
    }
    if (d.update_edge) {
      push(&L.uln_w, ctx->host_w[p + "edge_layer_norm.weight"]);
      push(&L.uln_b, ctx->host_w[p + "edge_layer_norm.bias"]);
    }
  }
  const int n_in = 4 + (d.expand_edge ? GAMD_NRBF : 0) + (d.use_bond ? 1 : 0);
  push(&mw.enc0_t, transposed("edge_encoder.mlp_layer.0.weight", H, n_in, d.expand_edge ? 64 : 32));
  push(&mw.enc0_b, ctx->host_w["edge_encoder.mlp_layer.0.bias"]);
  push(&mw.enc2_t, transposed("edge_encoder.mlp_layer.2.weight", H, H, H));
  push(&mw.enc2_b, ctx->host_w["edge_encoder.mlp_layer.2.bias"]);
  push(&mw.enc4_t, transposed("edge_encoder.mlp_layer.4.weight", De, H, H));
  push(&mw.enc4_b, ctx->host_w["edge_encoder.mlp_layer.4.bias"]);
  push(&mw.eln_w, ctx->host_w["edge_layer_norm.weight"]);
  push(&mw.eln_b, ctx->host_w["edge_layer_norm.bias"]);
  push(&mw.dec0_t, transposed("graph_decoder.mlp_layer.0.weight", H, D, D));
  push(&mw.dec0_b, ctx->host_w["graph_decoder.mlp_layer.0.bias"]);
  push(&mw.dec2_w, ctx->host_w["graph_decoder.mlp_layer.2.weight"]);
  push(&mw.dec2_b, ctx->host_w["graph_decoder.mlp_layer.2.bias"]);
  if (d.kind == GAMD_MODEL_LJ) {
    push(&mw.node_emb, ctx->host_w["node_emb"]);
  } else {
    push(&mw.nenc_w, ctx->host_w["node_encoder.weight"]);
    push(&mw.nenc_b, ctx->host_w["node_encoder.bias"]);
  }
  if (d.expand_edge) push(&mw.centers, ctx->host_w["edge_expand.centers"]);
  mw.length_mean = ctx->host_w["length_mean"][0];
  mw.length_std = ctx->host_w["length_std"][0];
  mw.n_layers = d.conv_layer;
  mw.n_edge_in = n_in;
  mw.use_bond = d.use_bond;
  mw.expand_edge = d.expand_edge;
  mw.kind = d.kind;
  if (ctx->d_wblob) cudaFree(ctx->d_wblob);
  ctx->d_wblob = nullptr;
  GAMD_CUDA(cudaMalloc(&ctx->d_wblob, blob.size() * sizeof(float)));
  GAMD_CUDA(cudaMemcpy(ctx->d_wblob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
  for (auto& f : fix) *f.first = ctx->d_wblob + f.second;
  if (d.use_bond && !ctx->d_bond) {
    ctx->err = "model uses the bond flag: call gamd_set_bonds before gamd_finalize_weights";
    return GAMD_ESTATE;
  }
  if (d.precision != GAMD_PREC_FP32) {
    // tensor-core operand images of the four edge-chain matrices of every layer:
    // B[n][k] = W[n][k] (torch Linear layout is already N x K, K-major), split x = hi + lo in bf16,
    // stored in the UMMA canonical SWIZZLE_128B K-major layout (two 64-wide K blocks of 16 KB)
    auto bf16_rn = [](float x) -> uint16_t {
      uint32_t u;
      memcpy(&u, &x, 4);
      if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
      return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
    };
    auto bf16_f = [](uint16_t h) -> float {
      uint32_t u = (uint32_t)h << 16;
      float f;
      memcpy(&f, &u, 4);
      return f;
    };
    const size_t chunk = 32768;
    std::vector<uint8_t> img((size_t)d.conv_layer * 8 * chunk, 0);
    std::vector<float> tb((size_t)d.conv_layer * 4 * 128, 0.f);
    const char* names[4] = {"edge_affine.mlp_layer.0", "edge_affine.mlp_layer.2", "theta_edge.mlp_layer.1",
                            "theta_edge.mlp_layer.3"};
    for (int l = 0; l < d.conv_layer; l++)
      for (int s = 0; s < 4; s++) {
        std::string p = "graph_conv.conv." + std::to_string(l) + "." + names[s];
        const std::vector<float>& w = ctx->host_w[p + ".weight"];
        const std::vector<float>& b = ctx->host_w[p + ".bias"];
        uint8_t* hi = img.data() + ((size_t)(l * 4 + s) * 2 + 0) * chunk;
        uint8_t* lo = img.data() + ((size_t)(l * 4 + s) * 2 + 1) * chunk;
        for (int n = 0; n < 128; n++)
          for (int k = 0; k < 128; k++) {
            float x = w[(size_t)n * 128 + k];
            uint16_t h = bf16_rn(x);
            uint16_t lw = bf16_rn(x - bf16_f(h));
            size_t off = (size_t)(k >> 6) * 16384 + (size_t)n * 128 + ((((k & 63) >> 3) ^ (n & 7)) << 4) + ((k & 7) << 1);
            memcpy(hi + off, &h, 2);
            memcpy(lo + off, &lw, 2);
          }
        for (int n = 0; n < 128; n++) tb[(size_t)(l * 4 + s) * 128 + n] = b[n];
      }
    // the same matrices for the CTA-pair kernel (tcgen05 cta_group::2): CTA h of a pair holds rows n = 64 h .. 64 h + 63
    // of B, so that a CTA's 128 KB (4 stages x [hi | lo] x 16 KB) are contiguous: [layer][half][stage][part].
    // N-split kernel (GAMD_MP_VARIANT=7): the GEMM runs as two N = 64 halves, half q = output columns 64 q .. 64 q + 63
    // with CTA h supplying columns 64 q + 32 h .. + 31 as rows 32 q .. 32 q + 31 of its image
    {
      const size_t part = 16384;
      std::vector<uint8_t> img2((size_t)d.conv_layer * 2 * 4 * 2 * part, 0);
      for (int l = 0; l < d.conv_layer; l++)
        for (int s = 0; s < 4; s++) {
          const std::vector<float>& w = ctx->host_w["graph_conv.conv." + std::to_string(l) + "." + names[s] + ".weight"];
          for (int n = 0; n < 128; n++) {
            const bool nsplit = ctx->mp_variant == 7;
            const int half = nsplit ? (n >> 5) & 1 : n >> 6, nl = nsplit ? (n >> 6) * 32 + (n & 31) : n & 63;
            uint8_t* hi = img2.data() + ((((size_t)l * 2 + half) * 4 + s) * 2 + 0) * part;
            uint8_t* lo = hi + part;
            for (int k = 0; k < 128; k++) {
              float x = w[(size_t)n * 128 + k];
              uint16_t h = bf16_rn(x);
              uint16_t lw = bf16_rn(x - bf16_f(h));
              size_t off = (size_t)(k >> 6) * 8192 + (size_t)nl * 128 + ((((k & 63) >> 3) ^ (nl & 7)) << 4) + ((k & 7) << 1);
              memcpy(hi + off, &h, 2);
              memcpy(lo + off, &lw, 2);
            }
          }
        }
      if (ctx->d_wimg2) cudaFree(ctx->d_wimg2);
      ctx->d_wimg2 = nullptr;
      GAMD_CUDA(cudaMalloc(&ctx->d_wimg2, img2.size()));
      GAMD_CUDA(cudaMemcpy(ctx->d_wimg2, img2.data(), img2.size(), cudaMemcpyHostToDevice));
    }
    if (ctx->d_wimg) cudaFree(ctx->d_wimg);
    if (ctx->d_tc_bias) cudaFree(ctx->d_tc_bias);
    ctx->d_wimg = nullptr;
    ctx->d_tc_bias = nullptr;
    GAMD_CUDA(cudaMalloc(&ctx->d_wimg, img.size()));
    GAMD_CUDA(cudaMemcpy(ctx->d_wimg, img.data(), img.size(), cudaMemcpyHostToDevice));
    GAMD_CUDA(cudaMalloc(&ctx->d_tc_bias, tb.size() * sizeof(float)));
    GAMD_CUDA(cudaMemcpy(ctx->d_tc_bias, tb.data(), tb.size() * sizeof(float), cudaMemcpyHostToDevice));
    // node matrices of every layer (+ the decoder's first Linear in slot 5 of every slab)
    {
      std::vector<uint8_t> nimg((size_t)d.conv_layer * 6 * 2 * chunk, 0);
      const char* nn[5] = {"phi_edge", "phi.mlp_layer.1", "src_affine", "dst_affine", "phi_dst"};
      for (int l = 0; l < d.conv_layer; l++)
        for (int mtx = 0; mtx < 6; mtx++) {
          const std::string key = mtx < 5 ? "graph_conv.conv." + std::to_string(l) + "." + nn[mtx] + ".weight"
                                          : std::string("graph_decoder.mlp_layer.0.weight");
          const std::vector<float>& w = ctx->host_w[key];
          uint8_t* hi = nimg.data() + ((size_t)(l * 6 + mtx) * 2 + 0) * chunk;
          uint8_t* lo = hi + chunk;
          for (int n = 0; n < 128; n++)
            for (int k = 0; k < 128; k++) {
              float x = w[(size_t)n * 128 + k];
              uint16_t h = bf16_rn(x);
              uint16_t lw = bf16_rn(x - bf16_f(h));
              size_t off = (size_t)(k >> 6) * 16384 + (size_t)n * 128 + ((((k & 63) >> 3) ^ (n & 7)) << 4) + ((k & 7) << 1);
              memcpy(hi + off, &h, 2);
              memcpy(lo + off, &lw, 2);
            }
        }
      if (ctx->d_wimg_node) cudaFree(ctx->d_wimg_node);
      ctx->d_wimg_node = nullptr;
      GAMD_CUDA(cudaMalloc(&ctx->d_wimg_node, nimg.size()));
      GAMD_CUDA(cudaMemcpy(ctx->d_wimg_node, nimg.data(), nimg.size(), cudaMemcpyHostToDevice));
    }
    // edge encoder: enc0 [128 x n_in] zero-padded to K = 64 (one K block), enc2 / enc4 [128 x 128]
    {
      std::vector<uint8_t> eimg(2 * 16384 + 4 * 32768, 0);
      std::vector<float> eb(3 * 128, 0.f);
      const char* en[3] = {"edge_encoder.mlp_layer.0", "edge_encoder.mlp_layer.2", "edge_encoder.mlp_layer.4"};
      size_t base = 0;
      for (int s = 0; s < 3; s++) {
        const std::vector<float>& w = ctx->host_w[std::string(en[s]) + ".weight"];
        const std::vector<float>& b = ctx->host_w[std::string(en[s]) + ".bias"];
        const int K = s == 0 ? n_in : 128;
        const size_t part = s == 0 ? 16384 : 32768;
        for (int n = 0; n < 128; n++)
          for (int k = 0; k < K; k++) {
            float x = w[(size_t)n * K + k];
            uint16_t h = bf16_rn(x);
            uint16_t lw = bf16_rn(x - bf16_f(h));
            size_t off = (size_t)(k >> 6) * 16384 + (size_t)n * 128 + ((((k & 63) >> 3) ^ (n & 7)) << 4) + ((k & 7) << 1);
            memcpy(eimg.data() + base + off, &h, 2);
            memcpy(eimg.data() + base + part + off, &lw, 2);
          }
        for (int n = 0; n < 128; n++) eb[(size_t)s * 128 + n] = b[n];
        base += 2 * part;
      }
      if (ctx->d_wimg_enc) cudaFree(ctx->d_wimg_enc);
      if (ctx->d_tc_bias_enc) cudaFree(ctx->d_tc_bias_enc);
      ctx->d_wimg_enc = nullptr;
      ctx->d_tc_bias_enc = nullptr;
      GAMD_CUDA(cudaMalloc(&ctx->d_wimg_enc, eimg.size()));
      GAMD_CUDA(cudaMemcpy(ctx->d_wimg_enc, eimg.data(), eimg.size(), cudaMemcpyHostToDevice));
      GAMD_CUDA(cudaMalloc(&ctx->d_tc_bias_enc, eb.size() * sizeof(float)));
      GAMD_CUDA(cudaMemcpy(ctx->d_tc_bias_enc, eb.data(), eb.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
  }
  ctx->finalized = true;
  ctx->graph_key = 0;
  return 0;
}

int gamd_check_async_errors(gamd_ctx* ctx, void* stream) {
  if (!ctx || !ctx->arena) return GAMD_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  int flag[4] = {0, 0, 0, 0};
  int ne = 0;
  GAMD_CUDA(cudaMemcpyAsync(flag, ctx->err_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  GAMD_CUDA(cudaMemcpyAsync(&ne, ctx->n_edges, sizeof(int), cudaMemcpyDeviceToHost, st));
  GAMD_CUDA(cudaStreamSynchronize(st));
  if (flag[0]) GAMD_CUDA(cudaMemsetAsync(ctx->err_flag, 0, 2 * sizeof(int), st));
  if (flag[0] & 1) {
    if (flag[1] > 0) ne = flag[1];
    ctx->err = "edge capacity exceeded: " + std::to_string(ne) + " edges needed, " + std::to_string(ctx->cap_edges) +
               " reserved; call gamd_reserve with a larger max_edges";
    return GAMD_ECAPACITY;
  }
  if (flag[0] & 2) {
    ctx->err = "edge list must be sorted by centre with ids in [0, n_atoms)";
    return GAMD_EINVAL;
  }
  if (flag[0] & 8) {
    ctx->err = "halo exchange: the neighbouring rank did not deliver its rows (peer-memory flag wait timed out)";
    return GAMD_ESTATE;
  }
  if (flag[0] & 4) {
    ctx->vl_key = 0;
    ctx->err = "neighbor candidate capacity exceeded (" + std::to_string(ctx->vl_cap) + " slots): call gamd_reserve with a "
               "larger max_edges";
    return GAMD_ECAPACITY;
  }
  return 0;
}

int gamd_neighbor_build(gamd_ctx* ctx, const float* d_pos, int64_t n_atoms, int32_t n_frames, const double h_box[3],
                        float cutoff, int32_t flags, void* stream) {
  if (!ctx || !d_pos || !h_box) return GAMD_EINVAL;
  if (!ctx->arena) {
    ctx->err = "no scratch reserved: call gamd_reserve";
    return GAMD_ESTATE;
  }
  GAMD_ENTER(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  float boxf[3] = {(float)h_box[0], (float)h_box[1], (float)h_box[2]};
  NbrParams p;
  int rc = nbr_setup_params(ctx, n_atoms, n_frames, boxf, cutoff, flags, &p);
  if (rc) return rc;
  if ((rc = nbr_bin_f32(ctx, d_pos, p, st))) return rc;
  return nbr_sort_and_sweep(ctx, p, nullptr, st);
}

int gamd_neighbor_count_host(gamd_ctx* ctx, int64_t* n_edges, void* stream) {
  if (!ctx || !n_edges || !ctx->arena) return GAMD_EINVAL;
  int rc = gamd_check_async_errors(ctx, stream);
  int ne = 0;
  GAMD_CUDA(cudaMemcpy(&ne, ctx->n_edges, sizeof(int), cudaMemcpyDeviceToHost));
  *n_edges = ne;
  return rc;
}

int gamd_neighbor_export(gamd_ctx* ctx, int64_t* d_edge_idx, int64_t cap, float* d_dist, float* d_norm,
                         void* stream) {
  if (!ctx || !d_edge_idx || !ctx->arena) return GAMD_EINVAL;
  return nbr_export(ctx, d_edge_idx, cap, d_dist, d_norm, (cudaStream_t)stream);
}

int gamd_model_forward(gamd_ctx* ctx, const float* d_pos, int64_t n_atoms, int32_t n_frames, const double h_box[3],
                       const int64_t* d_center, const int64_t* d_neigh, int64_t n_edges, const float* d_feat,
                       float* d_out, void* stream) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!d_pos || !h_box || !d_out || n_atoms <= 0 || n_frames <= 0 || n_atoms % n_frames || n_edges < 0 ||
      (n_edges > 0 && (!d_center || !d_neigh)))
    return GAMD_EINVAL;
  if (ctx->desc.kind != GAMD_MODEL_LJ && !d_feat) {
    ctx->err = "this model needs the node feature vector";
    return GAMD_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = check_bonds(ctx, n_atoms, n_frames))) return rc;
  if ((rc = csr_from_sorted_coo(ctx, d_center, d_neigh, n_atoms, n_edges, st))) return rc;
  if ((rc = pack_pos_feat(ctx, d_pos, d_feat, n_atoms, ctx->pos_feat_s, st))) return rc;
  float boxf[3] = {(float)h_box[0], (float)h_box[1], (float)h_box[2]};
  if ((rc = model_forward(ctx, ctx->pos_feat_s, nullptr, nullptr, n_atoms, (int)(n_atoms / n_frames), boxf, st)))
    return rc;
  GAMD_CUDA(cudaMemcpyAsync(d_out, ctx->pred, sizeof(float) * 3 * n_atoms, cudaMemcpyDeviceToDevice, st));
  return 0;
}

__global__ void k_unpermute_pred(const float* __restrict__ pred, const int* __restrict__ perm, int64_t n,
                                 float* __restrict__ out) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int64_t i = perm[s];
  out[3 * i] = pred[3 * s];
  out[3 * i + 1] = pred[3 * s + 1];
  out[3 * i + 2] = pred[3 * s + 2];
}

int gamd_dynbox_forward(gamd_ctx* ctx, const float* d_pos, int64_t n_atoms, const double h_box[3], float cutoff,
                        const float* d_feat, float* d_out, void* stream) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!d_pos || !h_box || !d_out || !d_feat || n_atoms <= 0) return GAMD_EINVAL;
  if (ctx->desc.kind != GAMD_MODEL_DYNBOX) {
    ctx->err = "gamd_dynbox_forward needs a GAMD_MODEL_DYNBOX context";
    return GAMD_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = check_bonds(ctx, n_atoms, 1))) return rc;
  float boxf[3] = {(float)h_box[0], (float)h_box[1], (float)h_box[2]};
  NbrParams p;
  // md_module.get_neighbor: |d| <= rc, no self pair, positions exactly as given (md_module.py:103-121)
  if ((rc = nbr_setup_params(ctx, n_atoms, 1, boxf, cutoff, GAMD_NBR_LE | GAMD_NBR_NOWRAP, &p))) return rc;
  prof_mark(ctx, "neighbor", st);
  if ((rc = nbr_bin_f32(ctx, d_pos, p, st))) return rc;
  if ((rc = nbr_sort_and_sweep(ctx, p, d_feat, st))) return rc;
  ctx->vl_key = 0;
  prof_mark(ctx, "neighbor", st);
  if ((rc = model_forward(ctx, ctx->pos_feat_s, nullptr, ctx->perm, n_atoms, (int)n_atoms, boxf, st))) return rc;
  k_unpermute_pred<<<ceil_div(n_atoms, 256), 256, 0, st>>>(ctx->pred, ctx->perm, n_atoms, d_out);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int gamd_compute_forces(gamd_ctx* ctx, const double* d_pos, int64_t n_atoms, int32_t n_frames, const double h_box[3],
                        float cutoff, const float* d_feat, double* d_force, void* stream) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!d_pos || !h_box || !d_force) return GAMD_EINVAL;
  if (ctx->desc.kind != GAMD_MODEL_LJ && !d_feat) {
    ctx->err = "this model needs the node feature vector";
    return GAMD_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = positions_to_forces(ctx, d_pos, 1.0, n_atoms, n_frames, h_box, cutoff, d_feat, st))) return rc;
  return integ_denorm_scatter(ctx, ctx->perm, d_force, nullptr, nullptr, 0.0, n_atoms, nullptr, st);
}

int gamd_compute_forces_host(gamd_ctx* ctx, const double* h_pos, int64_t n_atoms, int32_t n_frames,
                             const double h_box[3], float cutoff, const float* h_feat, double* h_force) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!h_pos || !h_force || n_atoms <= 0) return GAMD_EINVAL;
  if (n_atoms > ctx->cap_atoms) {
    ctx->err = "n_atoms exceeds reserved capacity; call gamd_reserve";
    return GAMD_ECAPACITY;
  }
  GAMD_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = 0;
  size_t b3 = sizeof(double) * 3 * (size_t)n_atoms;
  GAMD_CUDA(cudaMemcpyAsync(ctx->stage_a, h_pos, b3, cudaMemcpyHostToDevice, st));
  if (h_feat) GAMD_CUDA(cudaMemcpyAsync(ctx->stage_feat, h_feat, sizeof(float) * n_atoms, cudaMemcpyHostToDevice, st));
  if ((rc = gamd_compute_forces(ctx, ctx->stage_a, n_atoms, n_frames, h_box, cutoff, h_feat ? ctx->stage_feat : nullptr,
                                ctx->stage_b, st)))
    return rc;
  GAMD_CUDA(cudaMemcpyAsync(h_force, ctx->stage_b, b3, cudaMemcpyDeviceToHost, st));
  return gamd_check_async_errors(ctx, st);
}

int gamd_vv_first_half(gamd_ctx* ctx, double* d_x, double* d_v, const double* d_f, const double* d_mass,
                       int64_t n_atoms, double dt, void* stream) {
  if (!ctx || !d_x || !d_v || !d_f || !d_mass || n_atoms <= 0) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  return integ_first_half(ctx, d_x, d_v, d_f, d_mass, n_atoms, dt, (cudaStream_t)stream);
}

int gamd_vv_second_half(gamd_ctx* ctx, double* d_v, const double* d_f, const double* d_mass, int64_t n_atoms,
                        double dt, void* stream) {
  if (!ctx || !d_v || !d_f || !d_mass || n_atoms <= 0) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  return integ_second_half(ctx, d_v, d_f, d_mass, n_atoms, dt, (cudaStream_t)stream);
}

int gamd_md_run(gamd_ctx* ctx, double* d_x, double* d_v, double* d_f, const double* d_mass, int64_t n_atoms,
                int32_t n_frames, const double h_box[3], float cutoff, const float* d_feat, double dt,
                int32_t n_steps, double* d_ke, void* stream) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!d_x || !d_v || !d_f || !d_mass || !h_box || n_steps < 0) return GAMD_EINVAL;
  if (ctx->desc.kind != GAMD_MODEL_LJ && !d_feat) {
    ctx->err = "this model needs the node feature vector";
    return GAMD_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (n_steps == 0) return 0;
  if (d_ke) GAMD_CUDA(cudaMemsetAsync(d_ke, 0, sizeof(double) * n_steps, st));
  GAMD_CUDA(cudaMemsetAsync(ctx->d_stepctr, 0, sizeof(int), st));
  // the step program: plain velocity Verlet, or the thermostat / constraint programs chosen by gamd_md_configure
  const gamd_md_options& md = ctx->md;
  const bool nhc = md.thermostat == GAMD_THERMO_NHC && md.chain_length > 0;
  const bool langevin = md.thermostat == GAMD_THERMO_LANGEVIN;
  const bool rigid = md.rigid_water != 0;
  const double d_oh = md.d_oh > 0 ? md.d_oh : 0.09572, d_hh = md.d_hh > 0 ? md.d_hh : 0.15139006545273223;
  if (rigid && (n_atoms % 3 || n_frames < 1)) {
    ctx->err = "rigid_water needs [O,H,H] triplets";
    return GAMD_EINVAL;
  }
  if (langevin && rigid) {
    ctx->err = "the device-resident loop runs Langevin without constraints; rigid water: use Nose-Hoover or NVE";
    return GAMD_EUNSUPPORTED;
  }
  if ((nhc || langevin) && !ctx->d_nhc) {
    ctx->err = "gamd_md_configure was not called";
    return GAMD_ESTATE;
  }
  const int64_t n_mol = n_atoms / 3;
  if (nhc) {
    // KE2 = sum m v^2 of the caller's velocities for the first chain propagation; afterwards the sum is carried
    // analytically (scale^2 * KE2) or comes out of the kick kernel's reduction
    if ((rc = thermo_ke2(ctx, d_v, d_mass, n_atoms, ctx->d_ke2_acc, st))) return rc;
  }
  // fresh: the first chain of this call reads the reduction above, later ones the cached post-scaling sum
  auto one_step = [&](cudaStream_t s_, bool fresh) -> int {
    int r;
    if (nhc) {                                                    // propagateNHC (hack_integrator.py:271)
      if (fresh) r = thermo_chain(ctx, ctx->d_nhc, ctx->d_ke2_acc, dt, 0, nullptr, nullptr, s_);
      else r = thermo_chain_cached(ctx, ctx->d_nhc, dt, 0, s_);
      if (r) return r;
    }
    prof_mark(ctx, "integrate", s_);
    if (rigid) r = thermo_vv_first_rigid(ctx, d_x, d_v, d_f, d_mass, n_mol, dt, nhc, d_oh, d_hh, s_);
    else if (nhc) r = thermo_vv_first_scaled(ctx, d_x, d_v, d_f, d_mass, n_atoms, dt, s_);
    else if (langevin) r = thermo_langevin_first(ctx, d_x, d_v, d_f, d_mass, n_atoms, dt, md.kT, md.friction, nullptr, s_);
    else r = integ_first_half(ctx, d_x, d_v, d_f, d_mass, n_atoms, dt, s_);
    prof_mark(ctx, "integrate", s_);
    if (r) return r;
    // forces at x*10 Angstrom (test_nosehoover.py:112: value_in_unit(angstrom))
    if ((r = positions_to_forces(ctx, d_x, 10.0, n_atoms, n_frames, h_box, cutoff, d_feat, s_))) return r;
    // second half: kick (+ constrain velocities) (+ propagateNHC, bath energies); kinetic energy of the final velocities
    double* ke_kick = (rigid || nhc) ? nullptr : d_ke;
    double* acc_kick = (nhc && !rigid) ? ctx->d_ke2_acc : nullptr;
    if ((r = integ_denorm_scatter(ctx, ctx->perm, d_f, d_v, d_mass, dt, n_atoms, ke_kick, s_, -1, ctx->d_stepctr, acc_kick)))
      return r;
    if (rigid && (r = thermo_settle_vel(ctx, d_x, d_v, d_mass, n_mol, nhc ? ctx->d_ke2_acc : nullptr, nhc ? nullptr : d_ke,
                                        ctx->d_stepctr, s_)))
      return r;
    if (nhc) {
      if ((r = thermo_chain(ctx, ctx->d_nhc, ctx->d_ke2_acc, dt, 1, d_ke, ctx->d_stepctr, s_))) return r;
      if ((r = thermo_scale_v(ctx, ctx->d_nhc, d_v, n_atoms, s_))) return r;
    }
    return d_ke ? integ_inc_counter(ctx, ctx->d_stepctr, s_) : 0;
  };
  // the first step always runs eagerly (it also sets the one-time kernel attributes)
  const int64_t l0 = ctx->launches;
  if ((rc = one_step(st, true))) return rc;
  ctx->launches_last_step = ctx->launches - l0;
  int done = 1;
  // launch-bound systems: capture one step into a CUDA graph and replay it (every grid is shape-static: tile
  // loops read the edge count from device memory).  Large systems gain nothing and keep the eager path.
  const bool graph_ok = ctx->use_graphs && !ctx->prof_on && n_steps >= 4 && n_atoms <= 200000;
  if (graph_ok) {
    uint64_t key = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { key = (key ^ v) * 1099511628211ull; };
    mix((uint64_t)d_x); mix((uint64_t)d_v); mix((uint64_t)d_f); mix((uint64_t)d_mass); mix((uint64_t)d_feat);
    mix((uint64_t)d_ke); mix((uint64_t)n_atoms); mix((uint64_t)n_frames); mix((uint64_t)ctx->arena);
    uint64_t bits;
    for (int d = 0; d < 3; d++) { memcpy(&bits, &h_box[d], 8); mix(bits); }
    memcpy(&bits, &dt, 8); mix(bits);
    uint32_t cb; memcpy(&cb, &cutoff, 4); mix(cb);
    mix(ctx->md_generation);
    // (scaler, weights, bonds: gamd_set_scaler / gamd_finalize_weights / gamd_set_bonds reset graph_key themselves)
    if (!ctx->graph_stream) {
      GAMD_CUDA(cudaStreamCreateWithFlags(&ctx->graph_stream, cudaStreamNonBlocking));
      GAMD_CUDA(cudaEventCreateWithFlags(&ctx->graph_ev_in, cudaEventDisableTiming));
      GAMD_CUDA(cudaEventCreateWithFlags(&ctx->graph_ev_out, cudaEventDisableTiming));
    }
    if (key != ctx->graph_key || !ctx->graph_exec) {
      if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
      ctx->graph_exec = nullptr;
      ctx->graph_key = 0;
      cudaGraph_t graph = nullptr;
      const int64_t launches0 = ctx->launches;
      GAMD_CUDA(cudaStreamBeginCapture(ctx->graph_stream, cudaStreamCaptureModeThreadLocal));
      rc = one_step(ctx->graph_stream, false);
      cudaError_t ce = cudaStreamEndCapture(ctx->graph_stream, &graph);
      ctx->launches = launches0;   // captured, not launched
      if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      if (ce != cudaSuccess) {
        ctx->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce);
        return GAMD_ECUDA;
      }
      ce = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) {
        ctx->graph_exec = nullptr;
        ctx->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce);
        return GAMD_ECUDA;
      }
      ctx->graph_key = key;
      ctx->graph_launches_per_step = 0;
    }
    if (!ctx->graph_launches_per_step) {
      // kernels per replayed step = what one eager step issued (for gamd_launch_count bookkeeping)
      ctx->graph_launches_per_step = ctx->launches_last_step;
    }
    GAMD_CUDA(cudaEventRecord(ctx->graph_ev_in, st));
    GAMD_CUDA(cudaStreamWaitEvent(ctx->graph_stream, ctx->graph_ev_in, 0));
    for (; done < n_steps; done++) {
      GAMD_CUDA(cudaGraphLaunch(ctx->graph_exec, ctx->graph_stream));
      ctx->launches += ctx->graph_launches_per_step;
    }
    GAMD_CUDA(cudaEventRecord(ctx->graph_ev_out, ctx->graph_stream));
    GAMD_CUDA(cudaStreamWaitEvent(st, ctx->graph_ev_out, 0));
  }
  for (; done < n_steps; done++)
    if ((rc = one_step(st, false))) return rc;
  return 0;
}

int gamd_md_step_host(gamd_ctx* ctx, double* h_x, double* h_v, double* h_f, const double* h_mass, int64_t n_atoms,
                      int32_t n_frames, const double h_box[3], float cutoff, const float* h_feat, double dt) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!h_x || !h_v || !h_f || !h_mass || n_atoms <= 0) return GAMD_EINVAL;
  if (n_atoms > ctx->cap_atoms) {
    ctx->err = "n_atoms exceeds reserved capacity; call gamd_reserve";
    return GAMD_ECAPACITY;
  }
  GAMD_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = 0;
  size_t b3 = sizeof(double) * 3 * (size_t)n_atoms;
  GAMD_CUDA(cudaMemcpyAsync(ctx->stage_a, h_x, b3, cudaMemcpyHostToDevice, st));
  GAMD_CUDA(cudaMemcpyAsync(ctx->stage_b, h_v, b3, cudaMemcpyHostToDevice, st));
  GAMD_CUDA(cudaMemcpyAsync(ctx->stage_c, h_f, b3, cudaMemcpyHostToDevice, st));
  GAMD_CUDA(cudaMemcpyAsync(ctx->stage_m, h_mass, sizeof(double) * n_atoms, cudaMemcpyHostToDevice, st));
  if (h_feat) GAMD_CUDA(cudaMemcpyAsync(ctx->stage_feat, h_feat, sizeof(float) * n_atoms, cudaMemcpyHostToDevice, st));
  if ((rc = gamd_md_run(ctx, ctx->stage_a, ctx->stage_b, ctx->stage_c, ctx->stage_m, n_atoms, n_frames, h_box, cutoff,
                        h_feat ? ctx->stage_feat : nullptr, dt, 1, nullptr, st)))
    return rc;
  GAMD_CUDA(cudaMemcpyAsync(h_x, ctx->stage_a, b3, cudaMemcpyDeviceToHost, st));
  GAMD_CUDA(cudaMemcpyAsync(h_v, ctx->stage_b, b3, cudaMemcpyDeviceToHost, st));
  GAMD_CUDA(cudaMemcpyAsync(h_f, ctx->stage_c, b3, cudaMemcpyDeviceToHost, st));
  return gamd_check_async_errors(ctx, st);
}

int gamd_md_configure(gamd_ctx* ctx, const gamd_md_options* opt) {
  if (!ctx || !opt) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  if (opt->thermostat < GAMD_THERMO_NONE || opt->thermostat > GAMD_THERMO_LANGEVIN) return GAMD_EINVAL;
  int rc = thermo_alloc(ctx);
  if (rc) return rc;
  ctx->md = *opt;
  ctx->md_generation++;
  ctx->graph_key = 0;
  GAMD_CUDA(cudaMemsetAsync(ctx->d_rng_ctr, 0, sizeof(unsigned long long), 0));
  if (opt->thermostat == GAMD_THERMO_NHC) {
    if (opt->ndf <= 0) {
      ctx->err = "Nose-Hoover needs the number of degrees of freedom (3 N - constraints [- 3])";
      return GAMD_EINVAL;
    }
    return gamd_nhc_init_state(ctx, nullptr, opt->chain_length, opt->num_mts, opt->num_ys, opt->kT, opt->frequency, opt->ndf, 0);
  }
  return 0;
}

int gamd_nhc_init_state(gamd_ctx* ctx, gamd_nhc_state* d_state, int32_t chain_length, int32_t num_mts, int32_t num_ys,
                        double kT, double frequency, double ndf, void* stream) {
  if (!ctx) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  if (chain_length < 0 || chain_length > GAMD_NHC_MAX) {
    ctx->err = "Nose-Hoover chain length must be in 0.." + std::to_string(GAMD_NHC_MAX);
    return GAMD_EINVAL;
  }
  if (num_ys != 1 && num_ys != 3 && num_ys != 5) {
    ctx->err = "Invalid Yoshida-Suzuki value. Allowed values are: 1,3,5";
    return GAMD_EINVAL;
  }
  if (num_mts < 1 || !(frequency > 0) || !(kT > 0)) return GAMD_EINVAL;
  if (!d_state) {
    int rc = thermo_alloc(ctx);
    if (rc) return rc;
    d_state = ctx->d_nhc;
  }
  gamd_nhc_state h{};
  h.M = chain_length; h.n_c = num_mts; h.n_ys = num_ys;
  h.kT = kT; h.ndf = ndf; h.Qbase = kT / (frequency * frequency);
  for (int i = 0; i < chain_length; i++) {
    h.G[i] = -frequency * frequency;                                  // hack_integrator.py:257
    h.Q[i] = i == 0 ? ndf * h.Qbase : h.Qbase;
  }
  h.scale = 1.0;
  GAMD_CUDA(cudaMemcpyAsync(d_state, &h, sizeof(h), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  GAMD_CUDA(cudaStreamSynchronize((cudaStream_t)stream));              // `h` lives on this stack frame
  return 0;
}

int gamd_nhc_propagate(gamd_ctx* ctx, gamd_nhc_state* d_state, double* d_v, const double* d_mass, int64_t n_atoms,
                       double dt, int32_t bath, void* stream) {
  if (!ctx || !d_v || !d_mass || n_atoms <= 0) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  int rc = thermo_alloc(ctx);
  if (rc) return rc;
  if (!d_state) d_state = ctx->d_nhc;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = thermo_ke2(ctx, d_v, d_mass, n_atoms, ctx->d_ke2_acc, st))) return rc;
  if ((rc = thermo_chain(ctx, d_state, ctx->d_ke2_acc, dt, bath, nullptr, nullptr, st))) return rc;
  return thermo_scale_v(ctx, d_state, d_v, n_atoms, st);
}

int gamd_nhc_get_state(gamd_ctx* ctx, const gamd_nhc_state* d_state, gamd_nhc_state* h_out, void* stream) {
  if (!ctx || !h_out) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  if (!d_state) d_state = ctx->d_nhc;
  if (!d_state) return GAMD_ESTATE;
  GAMD_CUDA(cudaMemcpyAsync(h_out, d_state, sizeof(gamd_nhc_state), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  GAMD_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int gamd_nhc_set_state(gamd_ctx* ctx, gamd_nhc_state* d_state, const gamd_nhc_state* h_in, void* stream) {
  if (!ctx || !h_in) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  if (!d_state) d_state = ctx->d_nhc;
  if (!d_state) return GAMD_ESTATE;
  GAMD_CUDA(cudaMemcpyAsync(d_state, h_in, sizeof(gamd_nhc_state), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  GAMD_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int gamd_langevin_first_half(gamd_ctx* ctx, double* d_x, double* d_v, const double* d_f, const double* d_mass,
                             int64_t n_atoms, double dt, double kT, double friction, const double* d_gaussian,
                             void* stream) {
  if (!ctx || !d_x || !d_v || !d_f || !d_mass || n_atoms <= 0) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  int rc = thermo_alloc(ctx);
  if (rc) return rc;
  return thermo_langevin_first(ctx, d_x, d_v, d_f, d_mass, n_atoms, dt, kT, friction, d_gaussian, (cudaStream_t)stream);
}

int gamd_andersen_collide(gamd_ctx* ctx, double* d_v, const double* d_mass, int64_t n_atoms, double kT,
                          double p_collision, const double* d_uniform, const double* d_gaussian, void* stream) {
  if (!ctx || !d_v || !d_mass || n_atoms <= 0 || ((d_uniform == nullptr) != (d_gaussian == nullptr))) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  int rc = thermo_alloc(ctx);
  if (rc) return rc;
  return thermo_andersen(ctx, d_v, d_mass, n_atoms, kT, p_collision, d_uniform, d_gaussian, (cudaStream_t)stream);
}

int gamd_settle_positions(gamd_ctx* ctx, const double* d_x0, double* d_x, double* d_v, const double* d_mass,
                          int64_t n_mol, double dt_corr, double d_oh, double d_hh, void* stream) {
  if (!ctx || !d_x0 || !d_x || !d_mass || n_mol <= 0 || (d_v && !(dt_corr > 0))) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  return thermo_settle_pos(ctx, d_x0, d_x, d_v, d_mass, n_mol, dt_corr, d_oh > 0 ? d_oh : 0.09572,
                           d_hh > 0 ? d_hh : 0.15139006545273223, (cudaStream_t)stream);
}

int gamd_settle_velocities(gamd_ctx* ctx, const double* d_x, double* d_v, const double* d_mass, int64_t n_mol,
                           void* stream) {
  if (!ctx || !d_x || !d_v || !d_mass || n_mol <= 0) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  return thermo_settle_vel(ctx, d_x, d_v, d_mass, n_mol, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int gamd_tip4p_strip(gamd_ctx* ctx, const double* d_x4, double* d_x3, int64_t n_mol, void* stream) {
  if (!ctx || !d_x4 || !d_x3 || n_mol <= 0) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  return integ_tip4p_strip(ctx, d_x4, d_x3, n_mol, (cudaStream_t)stream);
}

int gamd_tip4p_unstrip(gamd_ctx* ctx, const double* d_a3, double* d_a4, int64_t n_mol, double w_o, double w_h,
                       int32_t place_m, void* stream) {
  if (!ctx || !d_a3 || !d_a4 || n_mol <= 0) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  return integ_tip4p_unstrip(ctx, d_a3, d_a4, n_mol, w_o, w_h, place_m, (cudaStream_t)stream);
}

int gamd_dd_begin(gamd_ctx* ctx, const double* d_pos, int64_t n_own, int64_t n_local, const double h_box[3],
                  float cutoff, const float* d_feat, void* stream) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!d_pos || !h_box || n_own <= 0 || n_local < n_own) return GAMD_EINVAL;
  if (ctx->desc.kind != GAMD_MODEL_LJ && !d_feat) {
    ctx->err = "this model needs the node feature vector";
    return GAMD_EINVAL;
  }
  if (ctx->wide) {
    ctx->err = "domain decomposition is built for the 128-wide tensor-core models only";
    return GAMD_EUNSUPPORTED;
  }
  if (ctx->desc.use_bond) {
    // the bond flag is looked up by frame-local atom id; a rank's local numbering is not the global one
    ctx->err = "domain decomposition of a model with the bond flag is not built (bond look-up needs global atom ids)";
    return GAMD_EUNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float boxf[3] = {(float)h_box[0], (float)h_box[1], (float)h_box[2]};
  NbrParams p;
  if ((rc = nbr_setup_params(ctx, n_local, 1, boxf, cutoff, GAMD_NBR_LT | GAMD_NBR_SELF, &p))) return rc;
  p.n_centers = (int)n_own;
  prof_mark(ctx, "neighbor", st);
  if (ctx->vl_skin_frac > 0.f && n_local >= ctx->vl_min_atoms && ctx->vl_cap < (int64_t(1) << 31)) {
    // candidate reuse while the caller keeps the local atom set unchanged (gamd_neighbor_invalidate after a change)
    if ((rc = nbr_step_verlet(ctx, d_pos, 1.0, h_box, p, d_feat, st))) return rc;
  } else {
    if ((rc = nbr_bin_f64(ctx, d_pos, 1.0, h_box, p, st))) return rc;
    if ((rc = nbr_sort_and_sweep(ctx, p, d_feat, st))) return rc;
    ctx->vl_key = 0;
  }
  prof_mark(ctx, "neighbor", st);
  ctx->dd_n_own = n_own;
  ctx->dd_n_loc = n_local;
  return model_begin(ctx, ctx->pos_feat_s, ctx->perm, n_local, (int)n_local, boxf, st);
}

int gamd_dd_layer(gamd_ctx* ctx, int32_t layer, void* stream) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (ctx->dd_n_loc <= 0 || layer < 0 || layer >= ctx->mw.n_layers) return GAMD_EINVAL;
  return model_layer(ctx, layer, ctx->pos_feat_s, ctx->dd_n_loc, (cudaStream_t)stream);
}

// one warp per 128-edge tile of the CSR: does any edge of the tile have a halo atom (local index >= n_own) as its
// source?  Tiles are appended to the interior (0) or boundary (1) list; the order inside a list does not matter
// (every tile writes its own receiver rows / partial-sum blocks).
__global__ void k_dd_split_tiles(const int* __restrict__ col, const int* __restrict__ perm,
                                 const int* __restrict__ n_edges, int n_own, int* __restrict__ list0,
                                 int* __restrict__ list1, int* __restrict__ counts) {
  const int tile = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int E = *n_edges;
  if ((int64_t)tile * 128 >= E) return;
  bool halo = false;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int e = tile * 128 + k * 32 + lane;
    if (e < E) halo |= perm[col[e]] >= n_own;
  }
  halo = __any_sync(0xffffffffu, halo);
  if (lane == 0) {
    const int i = atomicAdd(counts + (halo ? 1 : 0), 1);
    (halo ? list1 : list0)[i] = tile;
  }
}

int gamd_dd_split_tiles(gamd_ctx* ctx, void* stream) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (ctx->dd_n_loc <= 0) return GAMD_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  GAMD_CUDA(cudaMemsetAsync(ctx->tile_count, 0, 4 * sizeof(int), st));
  const int64_t max_tiles = ctx->cap_edges / 128 + 1;
  k_dd_split_tiles<<<ceil_div(max_tiles * 32, 256), 256, 0, st>>>(ctx->col_idx, ctx->perm, ctx->n_edges,
                                                                   (int)ctx->dd_n_own, ctx->tile_list[0],
                                                                   ctx->tile_list[1], ctx->tile_count);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int gamd_dd_layer_edges(gamd_ctx* ctx, int32_t layer, int32_t which, void* stream) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (ctx->dd_n_loc <= 0 || layer < 0 || layer >= ctx->mw.n_layers || which < -1 || which > 1) return GAMD_EINVAL;
  return model_layer_edges(ctx, layer, (cudaStream_t)stream, which);
}

int gamd_dd_layer_nodes(gamd_ctx* ctx, int32_t layer, void* stream) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (ctx->dd_n_loc <= 0 || layer < 0 || layer >= ctx->mw.n_layers) return GAMD_EINVAL;
  return model_layer_nodes(ctx, layer, ctx->pos_feat_s, ctx->dd_n_loc, (cudaStream_t)stream);
}

int gamd_dd_pack_rows(gamd_ctx* ctx, const int32_t* d_local_idx, int64_t n, float* d_out, void* stream) {
  if (!ctx || ctx->dd_n_loc <= 0 || n < 0 || (n > 0 && (!d_local_idx || !d_out))) return GAMD_EINVAL;
  if (n == 0) return 0;
  GAMD_ENTER(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  k_dd_pack<<<ceil_div(n * 32, 256), 256, 0, st>>>(d_local_idx, ctx->inv_perm, n, ctx->hn, ctx->srcA, d_out);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int gamd_dd_unpack_rows(gamd_ctx* ctx, int64_t first_local_idx, int64_t n, const float* d_in, void* stream) {
  if (!ctx || ctx->dd_n_loc <= 0 || n < 0 || first_local_idx < 0 || first_local_idx + n > ctx->dd_n_loc ||
      (n > 0 && !d_in))
    return GAMD_EINVAL;
  if (n == 0) return 0;
  GAMD_ENTER(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  k_dd_unpack<<<ceil_div(n * 32, 256), 256, 0, st>>>(first_local_idx, ctx->inv_perm, n, d_in, ctx->hn, ctx->srcA);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int gamd_peer_alloc(gamd_ctx* ctx, int64_t n_bytes, void** d_ptr, uint8_t h_handle[64]) {
  if (!ctx || n_bytes <= 0 || !d_ptr || !h_handle) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  GAMD_CUDA(cudaMalloc(&p, (size_t)n_bytes));
  GAMD_CUDA(cudaMemset(p, 0, (size_t)n_bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    ctx->err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e);
    return GAMD_ECUDA;
  }
  memcpy(h_handle, &h, 64);
  ctx->peer_allocs.push_back(p);
  *d_ptr = p;
  return 0;
}

int gamd_peer_open(gamd_ctx* ctx, const uint8_t h_handle[64], void** d_ptr) {
  if (!ctx || !h_handle || !d_ptr) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  cudaIpcMemHandle_t h;
  memcpy(&h, h_handle, 64);
  void* p = nullptr;
  GAMD_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  ctx->peer_opened.push_back(p);
  *d_ptr = p;
  return 0;
}

int gamd_dd_push_rows(gamd_ctx* ctx, const int32_t* d_local_idx, int64_t n, float* d_remote_rows,
                      unsigned long long* d_remote_flag, uint64_t seq, void* stream) {
  if (!ctx || ctx->dd_n_loc <= 0 || n < 0 || !d_remote_flag || (n > 0 && (!d_local_idx || !d_remote_rows))) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  if (n > 0) {
    k_dd_pack<<<ceil_div(n * 32, 256), 256, 0, st>>>(d_local_idx, ctx->inv_perm, n, ctx->hn, ctx->srcA, d_remote_rows);
    GAMD_LAUNCH_CHECK();
  }
  k_dd_signal<<<1, 1, 0, st>>>(d_remote_flag, (unsigned long long)seq);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int gamd_dd_arm_push(gamd_ctx* ctx, const int32_t* d_slot_left, float* d_remote_rows_left, const int32_t* d_slot_right,
                     float* d_remote_rows_right, int64_t n_own) {
  if (!ctx || ctx->dd_n_loc <= 0 || n_own != ctx->dd_n_own) return GAMD_EINVAL;
  if ((d_slot_left && !d_remote_rows_left) || (d_slot_right && !d_remote_rows_right)) return GAMD_EINVAL;
  if (ctx->desc.precision == GAMD_PREC_FP32) return GAMD_EUNSUPPORTED;     // the tensor-core node kernel carries the push
  ctx->dd_push_slot[0] = d_slot_left;
  ctx->dd_push_slot[1] = d_slot_right;
  ctx->dd_push_rows[0] = d_remote_rows_left;
  ctx->dd_push_rows[1] = d_remote_rows_right;
  ctx->dd_push_n = n_own;
  ctx->dd_push_armed = d_slot_left || d_slot_right;
  return 0;
}

int gamd_dd_push_bytes(gamd_ctx* ctx, const void* d_src, int64_t n_bytes, void* d_remote_dst,
                       unsigned long long* d_remote_flag, uint64_t seq, void* stream) {
  if (!ctx || n_bytes < 0 || (n_bytes & 15) || !d_remote_flag || (n_bytes > 0 && (!d_src || !d_remote_dst))) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_bytes > 0) {
    const int64_t n16 = n_bytes / 16;
    const int blocks = (int)std::min<int64_t>(ceil_div(n16, 256), 4 * ctx->sm_count);
    k_dd_push_bytes<<<blocks, 256, 0, st>>>(static_cast<const float4*>(d_src), static_cast<float4*>(d_remote_dst), n16);
    GAMD_LAUNCH_CHECK();
  }
  k_dd_signal<<<1, 1, 0, st>>>(d_remote_flag, (unsigned long long)seq);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int gamd_dd_wait_flag(gamd_ctx* ctx, const unsigned long long* d_flag, uint64_t seq, void* stream) {
  if (!ctx || !d_flag || !ctx->arena) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  k_dd_wait<<<1, 1, 0, (cudaStream_t)stream>>>(d_flag, (unsigned long long)seq, ctx->err_flag);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int gamd_dd_finish(gamd_ctx* ctx, double* d_force, double* d_v, const double* d_mass, double dt, double* d_ke,
                   void* stream) {
  if (!ctx || ctx->dd_n_loc <= 0 || !d_force || (d_v && !d_mass)) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  return integ_denorm_scatter(ctx, ctx->perm, d_force, d_v, d_mass, dt, ctx->dd_n_loc, d_ke, (cudaStream_t)stream,
                              ctx->dd_n_own);
}

int gamd_debug_ptr(gamd_ctx* ctx, const char* name, void** d_ptr, int64_t* n_bytes) {
  if (!ctx || !name || !d_ptr || !ctx->arena) return GAMD_EINVAL;
  std::string n(name);
  const int64_t A = ctx->cap_atoms, E = ctx->cap_edges;
  int64_t nb = 0;
  void* p = nullptr;
  if (n == "row_ptr") p = ctx->row_ptr, nb = 4 * (A + 1);
  else if (n == "col_idx") p = ctx->col_idx, nb = 4 * E;
  else if (n == "edge_dst") p = ctx->edge_dst, nb = 4 * E;
  else if (n == "perm") p = ctx->perm, nb = 4 * A;
  else if (n == "n_edges") p = ctx->n_edges, nb = 4;
  else if (n == "pos_sorted") p = ctx->pos_feat_s, nb = 16 * A;
  else if (n == "pos_nbr_sorted") p = ctx->pos_nbr_s, nb = 16 * A;
  else if (n == "e_emb") p = ctx->e_emb, nb = 4 * E * ctx->desc.edge_dim;
  else if (n == "h") p = ctx->h, nb = 4 * A * ctx->desc.encoding_size;
  else if (n == "hn") p = ctx->hn, nb = 4 * A * ctx->desc.encoding_size;
  else if (n == "agg") p = ctx->agg, nb = 4 * A * ctx->desc.encoding_size;
  else if (n == "pred") p = ctx->pred, nb = 4 * A * 3;
  else if (n == "dbg") p = ctx->e_emb + (size_t)E * ctx->desc.edge_dim, nb = 256 * GAMD_NF * 4;
  else {
    ctx->err = "unknown debug buffer: " + n;
    return GAMD_EINVAL;
  }
  *d_ptr = p;
  if (n_bytes) *n_bytes = nb;
  return 0;
}

int64_t gamd_launch_count(const gamd_ctx* ctx) { return ctx ? ctx->launches : 0; }

int gamd_neighbor_invalidate(gamd_ctx* ctx) {
  if (!ctx) return GAMD_EINVAL;
  ctx->vl_epoch++;
  return 0;
}

int gamd_neighbor_stats(gamd_ctx* ctx, int64_t* n_rebuilds, int64_t* n_searches, void* stream) {
  if (!ctx || !ctx->arena) return GAMD_EINVAL;
  GAMD_ENTER(ctx);
  unsigned long long h[2] = {0, 0};
  GAMD_CUDA(cudaMemcpyAsync(h, ctx->vl_counters, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  GAMD_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (n_rebuilds) *n_rebuilds = (int64_t)h[0];
  if (n_searches) *n_searches = (int64_t)h[1];
  return 0;
}

int gamd_profile_enable(gamd_ctx* ctx, int32_t on) {
  if (!ctx) return GAMD_EINVAL;
  ctx->prof_on = on != 0;
  return 0;
}

int gamd_profile_read(gamd_ctx* ctx, const char* stage, double* total_ms, int64_t* launches) {
  if (!ctx || !stage) return GAMD_EINVAL;
  GAMD_CUDA(cudaDeviceSynchronize());
  auto& p = ctx->prof[stage];
  for (size_t i = 0; i + 1 < p.used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.ev[i], p.ev[i + 1]) == cudaSuccess) {
      p.total_ms += ms;
      p.launches++;
    }
  }
  p.used = 0;
  if (total_ms) *total_ms = p.total_ms;
  if (launches) *launches = p.launches;
  p.total_ms = 0.0;
  p.launches = 0;
  return 0;
}

}  // extern "C"
