// Shared declarations of libgamd_b200: context, scratch arena, launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <map>
#include <vector>
#include "../../include/gamd_b200.h"

#define GAMD_NF 128          // feature width the tensor-core / 128-wide fp32 kernels are built for (D = H = De = 128)
#define GAMD_EDGE_TILE 64    // edges per tile of the fp32 edge kernels
#define GAMD_NODE_TILE 64    // nodes per tile of the node kernels
#define GAMD_MAX_BOND 4      // bonded partners kept per atom (water: O has 2, H has 1)
#define GAMD_NRBF 40

struct NbrParams {
  float box[3], half[3], inv_cell[3];
  int nc[3];
  int cells_per_frame;
  int n_atoms, atoms_per_frame, n_frames;
  int n_centers;   // atoms with caller index >= n_centers are neighbours only (halo atoms): no CSR row
  float rc, rc2;
  int flags;
};

// device copies of one message-passing layer's weights (transposed: Wt[k][n] = W[n][k])
struct LayerW {
  const float *ea0_t, *ea0_b, *ea2_t, *ea2_b;
  const float *src_t, *src_b, *dst_t, *dst_b;
  const float *te1_t, *te1_b, *te3_t, *te3_b;
  const float *pdst_t, *pdst_b, *pedge_t, *pedge_b;
  const float *phi_t, *phi_b;
  const float *ln_w, *ln_b;           // norm_layers[l]: LayerNorm, or BatchNorm1d when bn_mean != nullptr
  const float *bn_mean, *bn_var;      // BatchNorm1d running statistics (use_layer_norm = False), else nullptr
  const float *uln_w, *uln_b;         // conv[l].edge_layer_norm (update_edge), else nullptr
};

struct ModelW {
  LayerW layer[8];
  const float *enc0_t, *enc0_b, *enc2_t, *enc2_b, *enc4_t, *enc4_b, *eln_w, *eln_b;  // enc0_t is [64][128], zero padded
  const float *dec0_t, *dec0_b, *dec2_w, *dec2_b;   // dec2_w stays [3][128]
  const float *node_emb;                             // [128] (LJ)
  const float *nenc_w, *nenc_b;                      // [128] each (water, in_feats = 1)
  const float *centers;                              // [40]
  float length_mean, length_std;
  int n_layers, n_edge_in, use_bond, expand_edge, kind;
};

struct gamd_ctx {
  int device = 0;
  gamd_model_desc desc{};
  std::string err;
  int64_t launches = 0;

  // host-side weights by reference state-dict name
  std::map<std::string, std::vector<float>> host_w;
  bool finalized = false;
  float* d_wblob = nullptr;
  ModelW mw{};
  double scaler_mean = 0.0, scaler_var = 1.0;
  uint8_t* d_wimg = nullptr;      // tcgen05 weight images: [layer][4 stages][hi,lo][32 KB] SW128 K-major bf16
  float* d_tc_bias = nullptr;     // [layer][4][128]
  uint8_t* d_wimg2 = nullptr;     // CTA-pair images: [layer][2 halves of the 128 rows of B][4 stages][hi,lo][16 KB]
  uint8_t* d_wimg_enc = nullptr;  // edge encoder images: enc0 hi|lo (16 KB each, K=64), enc2 hi|lo, enc4 hi|lo (32 KB each)
  float* d_tc_bias_enc = nullptr; // [3][128]
  uint8_t* d_wimg_node = nullptr; // node matrices: [layer][pedge, phi, src, dst, pdst, dec0][hi|lo][32 KB]
  int* d_bond = nullptr;          // [atoms_per_frame][GAMD_MAX_BOND] frame-local partner ids, -1 padded
  int64_t bond_atoms = 0;

  // scratch
  int64_t cap_atoms = 0, cap_edges = 0;
  void* arena = nullptr;
  size_t arena_bytes = 0;
  // neighbor
  uint32_t *keys[2] = {nullptr, nullptr}, *vals[2] = {nullptr, nullptr};
  uint32_t* radix_hist = nullptr;
  uint32_t* scan_tmp = nullptr;
  float4 *pos_nbr = nullptr, *pos_feat = nullptr;             // caller order
  float4 *pos_nbr_s = nullptr, *pos_feat_s = nullptr;         // cell-sorted order
  int* perm = nullptr;                                        // sorted index -> caller index
  int* cell_start = nullptr;
  int64_t cap_cells = 0;
  int *deg = nullptr, *row_ptr = nullptr, *col_idx = nullptr, *edge_dst = nullptr;
  int* n_edges = nullptr;                                     // device scalar (== row_ptr[n])
  int* err_flag = nullptr;                                    // device: bit0 edge overflow
  // model
  float *e_emb = nullptr, *h = nullptr, *hn = nullptr, *srcA = nullptr, *dstA = nullptr, *pd = nullptr;
  float *agg = nullptr, *part = nullptr, *pred = nullptr;
  int* inv_perm = nullptr;                                    // caller index -> sorted index (domain decomposition)
  float* feat_s = nullptr;                                    // node type feature in sorted order
  // Verlet-skin reuse of the candidate list (neighbor.cu: nbr_step_verlet)
  float vl_skin_frac = 0.f;               // skin = vl_skin_frac * cutoff (the reference: dr_threshold = cutoff / 6); 0 = off
  bool small_frames = true;               // one-CTA-per-frame search for frames of <= 1024 atoms (GAMD_NBR_SMALL=0: off)
  int64_t vl_min_atoms = 20000;           // smaller systems rebuild every step (launch-bound: the gated pipeline costs more)
  int64_t vl_cap = 0;                     // candidate capacity
  int *vl_ptr = nullptr, *vl_cnt = nullptr, *vl_cand = nullptr, *vl_flag = nullptr;
  uint32_t* vl_mask = nullptr;
  float4* vl_pos_ref = nullptr;
  unsigned long long* vl_counters = nullptr;   // [rebuilds, steps]
  uint64_t vl_key = 0, vl_epoch = 0;
  // export helpers
  int *deg_o = nullptr, *row_ptr_o = nullptr;
  // host staging
  double *stage_a = nullptr, *stage_b = nullptr, *stage_c = nullptr, *stage_m = nullptr;
  float* stage_feat = nullptr;
  int64_t stage_atoms = 0;
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  NbrParams last_nbr{};
  // CUDA-graph replay of the MD step (launch-bound small systems)
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_key = 0;
  cudaStream_t graph_stream = nullptr;
  cudaEvent_t graph_ev_in = nullptr, graph_ev_out = nullptr;
  int* d_stepctr = nullptr;
  bool use_graphs = true;
  int64_t graph_launches_per_step = 0, launches_last_step = 0;
  int64_t dd_n_own = 0, dd_n_loc = 0;   // domain decomposition: owned / owned + halo atoms of the step in flight
  // 128-edge tiles split by "has a halo source": [0] interior, [1] boundary (lists + device counters)
  int* tile_list[2] = {nullptr, nullptr};
  int* tile_count = nullptr;
  int sm_count = 148;
  // one-time per-(device, kernel) opt-ins (cudaFuncSetAttribute is per device: tracked per context, not per process)
  uint32_t attr_mask = 0;
  // development switches read once at gamd_create (never on the launch path)
  bool dbg_timeline = false;
  int dd_reserve_sms = 0;
  // fused halo push (gamd_dd_arm_push): consumed by the next node kernel launch of gamd_dd_layer / gamd_dd_layer_nodes
  bool dd_push_armed = false;
  const int* dd_push_slot[2] = {nullptr, nullptr};
  float* dd_push_rows[2] = {nullptr, nullptr};
  int64_t dd_push_n = 0;
  int mp_row_prefetch = 0;   // GAMD_MP_ROW_PREFETCH
  int wait_hint_ns = 0;   // GAMD_WAIT_HINT_NS: mbarrier try_wait suspend-time hint in the tensor kernels' epilogues
  int mp_variant = 0;
  int enc_variant = 4;         // edge encoder: 4 = three tiles in flight, fixed service order (default); 3 = the same with
                               // the dynamic event loop; 0 = two tiles
  int64_t model_atoms = 0;     // atoms of the forward pass in flight (model_begin)
  // generic-width fp32 path (model_wide.cu): widths other than 128, update_edge, no RBF expansion, BatchNorm
  bool wide = false;
  int wide_r = 4, wide_xs = 132;   // rows per thread (tile = 16 r rows) and shared-memory row stride
  int mp_small_atoms = 4096;   // systems up to this size use the two-tile single-CTA edge kernel (GAMD_MP_SMALL_ATOMS)

  // halo exchange over peer memory: buffers this rank exposes (cudaMalloc + IPC handle) / neighbours' buffers it opened
  std::vector<void*> peer_allocs, peer_opened;
  // thermostat / constraints of the device-resident loop (gamd_md_configure)
  gamd_md_options md{};
  uint64_t md_generation = 0;
  gamd_nhc_state* d_nhc = nullptr;            // chain state + [ke2 accumulator][variate step counter] behind it
  double* d_ke2_acc = nullptr;
  unsigned long long* d_rng_ctr = nullptr;

  // optional per-stage CUDA-event timers (gamd_profile_enable / gamd_profile_read)
  struct StageProf {
    std::vector<cudaEvent_t> ev;   // begin/end pairs
    size_t used = 0;
    double total_ms = 0.0;
    int64_t launches = 0;
  };
  std::map<std::string, StageProf> prof;
  bool prof_on = false;
};

void prof_mark(gamd_ctx* ctx, const char* stage, cudaStream_t st);   // call before and after the stage

#define GAMD_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                     \
      return GAMD_ECUDA;                                                                 \
    }                                                                                    \
  } while (0)

#define GAMD_LAUNCH_CHECK()                                                              \
  do {                                                                                   \
    ctx->launches++;                                                                     \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      ctx->err = std::string("kernel launch at ") + __FILE__ + ":" + std::to_string(__LINE__) + ": " + cudaGetErrorString(_e); \
      return GAMD_ECUDA;                                                                 \
    }                                                                                    \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

enum { GAMD_ATTR_FP32 = 1, GAMD_ATTR_MP_TC = 2, GAMD_ATTR_ENC_TC = 4, GAMD_ATTR_NODE_TC = 8, GAMD_ATTR_MP_TC3 = 16, GAMD_ATTR_MP_TC2 = 32, GAMD_ATTR_NBR_SMALL = 64, GAMD_ATTR_ENC_TC3 = 128, GAMD_ATTR_WIDE = 256 };

// every stream entry point runs on the context's device, whatever device the calling thread had current
#define GAMD_ENTER(ctx)                                                                  \
  do {                                                                                   \
    int _dev = -1;                                                                       \
    if (cudaGetDevice(&_dev) != cudaSuccess || _dev != (ctx)->device) GAMD_CUDA(cudaSetDevice((ctx)->device)); \
  } while (0)

// ---- stage entry points implemented in the .cu files (host functions) ----
int nbr_setup_params(gamd_ctx* ctx, int64_t n_atoms, int n_frames, const float box[3], float rc, int flags, NbrParams* p);
int nbr_bin_f32(gamd_ctx* ctx, const float* d_pos, const NbrParams& p, cudaStream_t st);
int nbr_bin_f64(gamd_ctx* ctx, const double* d_pos_or_x, double scale, const double* box64, const NbrParams& p, cudaStream_t st);
int nbr_sort_and_sweep(gamd_ctx* ctx, const NbrParams& p, const float* d_feat, cudaStream_t st);
int nbr_small_frames(gamd_ctx* ctx, const double* d_x, double scale, const double* box64, const NbrParams& p,
                     const float* d_feat, cudaStream_t st);
int nbr_step_verlet(gamd_ctx* ctx, const double* d_x, double scale, const double* box64, const NbrParams& p,
                    const float* d_feat, cudaStream_t st);
int nbr_export(gamd_ctx* ctx, int64_t* d_edge_idx, int64_t cap, float* d_dist, float* d_norm, cudaStream_t st);
int csr_from_sorted_coo(gamd_ctx* ctx, const int64_t* d_center, const int64_t* d_neigh, int64_t n_atoms, int64_t n_edges, cudaStream_t st);
int exclusive_scan_i32(gamd_ctx* ctx, const int* d_in, int* d_out, int64_t n, cudaStream_t st);

// which: -1 every tile; 0 / 1 only the interior / boundary tiles of the domain-decomposition split (ctx->tile_list)
int mp_edge_tc_launch(gamd_ctx* ctx, int layer, cudaStream_t st, int which = -1);
int mp_edge_tc3_launch(gamd_ctx* ctx, int layer, cudaStream_t st, int which, bool safe_war);
int mp_edge_tc2_launch(gamd_ctx* ctx, int layer, cudaStream_t st, int which, bool safe_war, bool nsplit);
int node_update_tc_launch(gamd_ctx* ctx, int mode, int layer, const float4* pos_feat, int64_t n_atoms, cudaStream_t st);
int edge_encode_tc_launch(gamd_ctx* ctx, const float4* pos_feat, const int* orig_id, int atoms_per_frame,
                          const float box[3], cudaStream_t st);
int model_begin(gamd_ctx* ctx, const float4* pos_feat, const int* orig_id, int64_t n_atoms, int atoms_per_frame,
                const float box[3], cudaStream_t st);
int model_layer(gamd_ctx* ctx, int l, const float4* pos_feat, int64_t n_atoms, cudaStream_t st);
int model_layer_edges(gamd_ctx* ctx, int l, cudaStream_t st, int which);
int model_layer_nodes(gamd_ctx* ctx, int l, const float4* pos_feat, int64_t n_atoms, cudaStream_t st);
int model_forward(gamd_ctx* ctx, const float4* pos_feat, const float* feat, const int* orig_id,
                       int64_t n_atoms, int atoms_per_frame, const float box[3], cudaStream_t st);

// generic-width fp32 path (model_wide.cu)
int wide_plan(int D, int H, int De, int* r_out, int* xs_out);
int wide_begin(gamd_ctx* ctx, const float4* pos_feat, const int* orig_id, int64_t n_atoms, int atoms_per_frame,
               const float box[3], cudaStream_t st);
int wide_layer_edges(gamd_ctx* ctx, int l, cudaStream_t st);
int wide_layer_nodes(gamd_ctx* ctx, int l, const float4* pos_feat, int64_t n_atoms, cudaStream_t st);

int integ_first_half(gamd_ctx* ctx, double* x, double* v, const double* f, const double* mass, int64_t n, double dt, cudaStream_t st);
int integ_second_half(gamd_ctx* ctx, double* v, const double* f, const double* mass, int64_t n, double dt, cudaStream_t st);
int integ_denorm_scatter(gamd_ctx* ctx, const int* perm, double* f_out, double* v, const double* mass, double dt, int64_t n, double* ke_out, cudaStream_t st, int64_t n_own = -1, const int* ke_slot = nullptr, double* ke2_acc = nullptr);
int integ_inc_counter(gamd_ctx* ctx, int* counter, cudaStream_t st);
int integ_tip4p_strip(gamd_ctx* ctx, const double* x4, double* x3, int64_t n_mol, cudaStream_t st);
int integ_tip4p_unstrip(gamd_ctx* ctx, const double* a3, double* a4, int64_t n_mol, double wo, double wh, int place_m, cudaStream_t st);
int thermo_alloc(gamd_ctx* ctx);
int thermo_ke2(gamd_ctx* ctx, const double* v, const double* mass, int64_t n, double* acc, cudaStream_t st);
int thermo_chain(gamd_ctx* ctx, gamd_nhc_state* d_state, double* ke2_acc, double dt, int bath, double* ke_out,
                 const int* ke_slot, cudaStream_t st);
int thermo_chain_cached(gamd_ctx* ctx, gamd_nhc_state* d_state, double dt, int bath, cudaStream_t st);
int thermo_scale_v(gamd_ctx* ctx, const gamd_nhc_state* d_state, double* v, int64_t n, cudaStream_t st);
int thermo_vv_first_scaled(gamd_ctx* ctx, double* x, double* v, const double* f, const double* mass, int64_t n, double dt,
                           cudaStream_t st);
int thermo_langevin_first(gamd_ctx* ctx, double* x, double* v, const double* f, const double* mass, int64_t n, double dt,
                          double kT, double friction, const double* gaussian, cudaStream_t st);
int thermo_andersen(gamd_ctx* ctx, double* v, const double* mass, int64_t n, double kT, double p_coll,
                    const double* uniform, const double* gaussian, cudaStream_t st);
int thermo_settle_pos(gamd_ctx* ctx, const double* x0, double* x, double* v, const double* mass, int64_t n_mol,
                      double dt_corr, double d_oh, double d_hh, cudaStream_t st);
int thermo_vv_first_rigid(gamd_ctx* ctx, double* x, double* v, const double* f, const double* mass, int64_t n_mol,
                          double dt, bool scaled, double d_oh, double d_hh, cudaStream_t st);
int thermo_settle_vel(gamd_ctx* ctx, const double* x, double* v, const double* mass, int64_t n_mol, double* ke2_acc,
                      double* ke_out, const int* ke_slot, cudaStream_t st);
int pack_pos_feat(gamd_ctx* ctx, const float* d_pos, const float* d_feat, int64_t n, float4* out, cudaStream_t st);
