/* gamd_b200 - C ABI of the B200-native GAMD hot path
 *
 * positions + box --> periodic neighbor search --> MDNet message passing --> per-atom
 * forces --> velocity-Verlet update.
 *
 * The reference (BaratiLab/GAMD) has no FFI: its boundary for this path is four Python call
 * surfaces.  Each entry point below names the reference interface it replaces (paths are
 * relative to the reference tree, file:line).  INTEGRATION.md shows the ctypes stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative GAMD_E* code otherwise;
 *     gamd_last_error(ctx) returns a human-readable message.  No C++ exception crosses.
 *   - pointers named d_* are caller-owned DEVICE pointers (e.g. torch tensors); h_* are HOST
 *     pointers.  The library never frees caller memory.  Its own scratch memory is sized by
 *     gamd_reserve() and lives until gamd_destroy().
 *   - all work is enqueued on the cudaStream_t passed as `stream` (void*, 0 = default
 *     stream); nothing synchronises the host unless the function name ends in _host or
 *     the comment says so.
 *   - one ctx per device and per host thread; distinct ctxs are independent.
 *   - model shapes: encoding_size = hidden_dim = edge_embedding_dim = 128 with LayerNorm and the RBF expansion
 *     (every LJ / TIP3P / TIP4P config the reference ships, code/LJ/test_script/test_nosehoover.py:66-69, and the
 *     128-wide dynamic-box model) runs on the tensor-core kernels in the requested precision.  Every other shape -
 *     widths that are multiples of 128 up to 1024 (the 256 / 128 / 256 x 5 DFT-water model of
 *     code/water/test_script/test_nosehoover_hb.py:69-81 and wider), update_edge, expand_edge = 0, BatchNorm -
 *     runs on a generic-width CUDA-core path in fp32 whatever precision was requested; such a context does not
 *     take part in domain decomposition (gamd_dd_* return GAMD_EUNSUPPORTED).
 */
#ifndef GAMD_B200_H
#define GAMD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gamd_ctx gamd_ctx;

enum {
  GAMD_OK = 0,
  GAMD_EINVAL = -1,        /* bad argument */
  GAMD_ECUDA = -2,         /* CUDA runtime error (text in gamd_last_error) */
  GAMD_EUNSUPPORTED = -3,  /* model shape / option not built */
  GAMD_ECAPACITY = -4,     /* edge or atom capacity exceeded: the analogue of jax-md's
                              did_buffer_overflow (code/graph_utils.py:41); call gamd_reserve
                              with a larger capacity and retry */
  GAMD_ESTATE = -5,        /* call order violated (e.g. weights not finalized) */
  GAMD_ENOGPU = -6         /* no usable CUDA device: there is NO CPU fallback */
};

/* neighbor predicate flags */
enum {
  GAMD_NBR_LT = 0,        /* dr2 <  rc*rc     (code/graph_utils.py:59, jax-md path)      */
  GAMD_NBR_LE = 1,        /* |dr| <= rc       (code/md_module.py:111, torch brute force) */
  GAMD_NBR_SELF = 2,      /* keep the i==i pair (mask_self=False, code/graph_utils.py:25) */
  GAMD_NBR_NOWRAP = 4     /* predicate uses the raw positions (md_module.get_neighbor
                             applies no jnp.mod); cells are still assigned from wrapped ones */
};

/* model kinds: which reference nn.Module the weights belong to */
enum {
  GAMD_MODEL_LJ = 0,      /* SimpleMDNetNew        code/nn_module.py:561-685 */
  GAMD_MODEL_WATER = 1,   /* WaterMDNetNew         code/nn_module.py:410-558 */
  GAMD_MODEL_DYNBOX = 2   /* WaterMDDynamicBoxNet  code/nn_module.py:266-407 */
};

/* arithmetic of the edge-sized GEMMs */
enum {
  GAMD_PREC_FP32 = 0,     /* CUDA-core FFMA, fp32 everywhere (parity anchor)            */
  GAMD_PREC_BF16X3 = 1,   /* tcgen05, 3-pass split-bf16, fp32 accumulate ("exact" mode) */
  GAMD_PREC_BF16 = 2      /* tcgen05, single-pass bf16 operands ("fast" mode)           */
};

typedef struct {
  int32_t kind;            /* GAMD_MODEL_* */
  int32_t encoding_size;   /* D  */
  int32_t hidden_dim;      /* H  */
  int32_t edge_dim;        /* De */
  int32_t conv_layer;      /* L  */
  int32_t in_feats;        /* node_encoder input width (water: 1), 0 for LJ */
  int32_t use_bond;        /* 1: last edge feature is the bond flag (45 inputs) */
  int32_t expand_edge;     /* 1: 40-centre RBF expansion */
  int32_t precision;       /* GAMD_PREC_* */
  int32_t update_edge;     /* 1: every layer replaces the edge embedding by its edge_layer_norm(e_emb)
                              (code/nn_module.py:89-90, :139-146); needs edge_dim == encoding_size */
  int32_t batch_norm;      /* 1: norm_layers are eval-mode BatchNorm1d (use_layer_norm=False,
                              code/nn_module.py:193-196) instead of LayerNorm */
} gamd_model_desc;

/* ---- lifetime ------------------------------------------------------------------------ */
/* replaces: model construction, code/LJ/train_network_lj.py:68-88 (build_model) */
int gamd_create(int device, const gamd_model_desc* desc, gamd_ctx** out);
int gamd_destroy(gamd_ctx* ctx);
const char* gamd_last_error(const gamd_ctx* ctx);   /* ctx may be NULL: last create error */
const char* gamd_version(void);

/* scratch for up to max_atoms nodes and max_edges directed edges (device allocations
 * happen here, never on the hot path).  May be called again to grow. */
int gamd_reserve(gamd_ctx* ctx, int64_t max_atoms, int64_t max_edges);

/* ---- weights ------------------------------------------------------------------------- */
/* replaces: model.load_state_dict / load_from_checkpoint, code/LJ/train_network_lj.py:85-87,
 * code/LJ/test_script/test_nosehoover.py:77.  `name` is the reference state-dict key
 * (SURVEY.md section 8a note 9), `h_data` the fp32 tensor in torch's row-major layout. */
int gamd_load_weight(gamd_ctx* ctx, const char* name, const float* h_data, int64_t n);
/* replaces: load_training_stats (scaler.npz), code/LJ/train_network_lj.py:119-123 */
int gamd_set_scaler(gamd_ctx* ctx, double mean, double var);
/* bond list for the water model's bond flag: nb pairs (i,j) of frame-local atom ids,
 * made symmetric internally. replaces: build_bond_graph, code/nn_module.py:529-534 */
int gamd_set_bonds(gamd_ctx* ctx, const int64_t* h_bonds, int64_t nb, int64_t n_atoms_per_frame);
/* checks that every tensor arrived, builds the transposed / split device copies */
int gamd_finalize_weights(gamd_ctx* ctx);

/* ---- stage 1: neighbor search ---------------------------------------------------------- */
/* replaces: NeighborSearcher.init_new_neighbor_lst / update_neighbor_lst
 *           (code/graph_utils.py:29-44), nbrlst_to_edge_mask (code/graph_utils.py:51-61),
 *           search_for_neighbor + get_edge_idx (code/LJ/train_network_lj.py:166-199) and
 *           get_neighbor (code/md_module.py:93-126, flags LE|NOWRAP, no SELF).
 * d_pos: fp32 [n_atoms,3] Angstrom, any image; h_box is the box in double (the reference's
 * python float); the fp32 box used by the kernels is (float)h_box[d], as jnp.array does.  n_frames independent frames of
 * n_atoms/n_frames atoms each share the box (block-diagonal batch, code/nn_module.py:655-661).
 * Builds, inside ctx, a receiver-sorted CSR in cell-sorted index space. */
int gamd_neighbor_build(gamd_ctx* ctx, const float* d_pos, int64_t n_atoms, int32_t n_frames,
                        const double h_box[3], float cutoff, int32_t flags, void* stream);
/* number of edges of the last build; synchronises `stream`. */
int gamd_neighbor_count_host(gamd_ctx* ctx, int64_t* n_edges, void* stream);
/* COO export in the reference's format: int64 [2,cap] row 0 centre, row 1 neighbour, caller
 * atom ids, centre-major, neighbour ascending (code/LJ/train_network_lj.py:182-185).
 * d_dist / d_norm (optional, may be NULL): fp32 [cap,3] min-image pos[centre]-pos[neigh] and
 * its norm, as get_neighbor returns (code/md_module.py:119-121). */
int gamd_neighbor_export(gamd_ctx* ctx, int64_t* d_edge_idx, int64_t cap, float* d_dist,
                         float* d_norm, void* stream);

/* ---- stage 2: MDNet forward on an explicit edge list ----------------------------------- */
/* replaces: SimpleMDNetNew.forward (code/nn_module.py:672-685), WaterMDNetNew.forward
 *           (:545-558).  d_pos fp32 [n,3] already wrapped by the caller (as
 *           train_network_lj.py:141 does), d_center / d_neigh int64 [n_edges] sorted by
 *           centre (the order get_edge_idx produces), d_feat fp32 [n] node type feature
 *           (water) or NULL (LJ), frames are concatenated with frame-local ids already
 *           offset.  d_out fp32 [n,3] normalised force. */
int gamd_model_forward(gamd_ctx* ctx, const float* d_pos, int64_t n_atoms, int32_t n_frames,
                       const double h_box[3], const int64_t* d_center, const int64_t* d_neigh,
                       int64_t n_edges, const float* d_feat, float* d_out, void* stream);

/* ---- dynamic-box model: neighbor search inside the model ---------------------------------- */
/* replaces: WaterMDDynamicBoxNet.forward for ONE frame (code/nn_module.py:391-407 -> build_graph :338-365 ->
 *           md_module.get_neighbor, code/md_module.py:93-126): per-axis box h_box, |d| <= cutoff, no self edges,
 *           positions exactly as given (the facade wraps them first, code/water/train_network_real_large.py:150),
 *           edge direction -(min-image of pos[center] - pos[neigh]) (:327).  d_feat fp32 [n] node feature,
 *           d_out fp32 [n,3] normalised force in the caller's atom order.  Needs a GAMD_MODEL_DYNBOX context. */
int gamd_dynbox_forward(gamd_ctx* ctx, const float* d_pos, int64_t n_atoms, const double h_box[3], float cutoff,
                        const float* d_feat, float* d_out, void* stream);

/* ---- stage 1+2 fused: positions -> forces ---------------------------------------------- */
/* replaces: ParticleNetLightning.predict_forces (code/LJ/train_network_lj.py:133-157,
 *           code/water/train_network_tip3p.py:142-159): neighbor search on fp32(pos), model
 *           on fp32(mod(pos, L)), de-normalisation in fp64.
 * d_pos fp64 [n,3] Angstrom; d_force fp64 [n,3] kJ/mol/nm. */
int gamd_compute_forces(gamd_ctx* ctx, const double* d_pos, int64_t n_atoms, int32_t n_frames,
                        const double h_box[3], float cutoff, const float* d_feat, double* d_force,
                        void* stream);
/* same through HOST buffers (H2D + D2H inside, synchronous) - the predict_forces call shape */
int gamd_compute_forces_host(gamd_ctx* ctx, const double* h_pos, int64_t n_atoms, int32_t n_frames,
                             const double h_box[3], float cutoff, const float* h_feat,
                             double* h_force);

/* ---- stage 3: integrator hook ------------------------------------------------------------ */
/* replaces: HackNoseHooverIntegrator step body with chain_length=0
 *           (code/hack_integrator.py:271-277: v+=0.5*dt*force_last/m; x+=dt*v) */
int gamd_vv_first_half(gamd_ctx* ctx, double* d_x, double* d_v, const double* d_f,
                       const double* d_mass, int64_t n_atoms, double dt, void* stream);
/* replaces: HackHalfVelocityIntegrator (code/hack_integrator.py:171-178:
 *           v+=(dt/2)*gnn_force/m) */
int gamd_vv_second_half(gamd_ctx* ctx, double* d_v, const double* d_f, const double* d_mass,
                        int64_t n_atoms, double dt, void* stream);

/* ---- whole MD step, device resident ------------------------------------------------------ */
/* replaces: the driver loop body code/LJ/test_script/test_nosehoover.py:100-118 (NVE program):
 *   first half-kick + drift with d_f, forces at the new positions, second half-kick.
 * d_x (nm), d_v (nm/ps), d_f (kJ/mol/nm, holds F(x) on entry and F(x') on exit) fp64 [n,3],
 * d_mass fp64 [n] (Da).  Runs n_steps steps; no host synchronisation.
 * d_ke (optional): fp64 [n_steps] total kinetic energy after each step (kJ/mol). */
int gamd_md_run(gamd_ctx* ctx, double* d_x, double* d_v, double* d_f, const double* d_mass,
                int64_t n_atoms, int32_t n_frames, const double h_box[3], float cutoff,
                const float* d_feat, double dt, int32_t n_steps, double* d_ke, void* stream);
/* one step through HOST buffers (H2D of x,v,f + step + D2H of x,v,f), synchronous */
int gamd_md_step_host(gamd_ctx* ctx, double* h_x, double* h_v, double* h_f, const double* h_mass,
                      int64_t n_atoms, int32_t n_frames, const double h_box[3], float cutoff,
                      const float* h_feat, double dt);

/* ---- thermostats and rigid-water constraints of the integrator hook ------------------------------------- */
/* Nose-Hoover chain state (the globals xi{i}, vxi{i}, G{i}, Q{i}, scale, KE2, bathKE, bathPE of
 * code/hack_integrator.py:249-261).  Lives in DEVICE memory: either inside the ctx (gamd_md_configure, used by
 * gamd_md_run) or in a caller-owned device buffer of sizeof(gamd_nhc_state) bytes (one per Hack*Integrator object,
 * so that copy_state_from_integrator is a device-to-device copy).  gamd_nhc_get_state / _set_state move it to / from
 * the host (synchronous). */
#define GAMD_NHC_MAX 16
typedef struct {
  int32_t M, n_c, n_ys, pad_;    /* chain_length (<= GAMD_NHC_MAX), num_mts, num_yoshidasuzuki (1, 3, 5) */
  double kT, ndf, Qbase;         /* kJ/mol; degrees of freedom; Q = kT / frequency^2 (Q0 = ndf * Q, Q_i = Q) */
  double xi[GAMD_NHC_MAX], vxi[GAMD_NHC_MAX], G[GAMD_NHC_MAX], Q[GAMD_NHC_MAX];
  double scale, ke2_in, ke2, bathKE, bathPE;   /* last velocity scale; sum m v^2 before / after it; bath energies */
} gamd_nhc_state;

enum { GAMD_THERMO_NONE = 0, GAMD_THERMO_NHC = 1, GAMD_THERMO_LANGEVIN = 2 };
typedef struct {
  int32_t thermostat;            /* GAMD_THERMO_* */
  int32_t chain_length, num_mts, num_ys;   /* NHC: hack_integrator.py:191-192 defaults 5, 5, 5 */
  double kT;                     /* kJ/mol */
  double frequency;              /* NHC collision_frequency, 1/ps */
  double ndf;                    /* degrees of freedom; <= 0: 3 n_atoms (minus 3 per rigid molecule) */
  double friction;               /* Langevin collision_rate, 1/ps */
  uint64_t seed;                 /* Philox key of the Langevin / Andersen variates */
  int32_t rigid_water;           /* 1: atoms are [O,H,H] triplets held rigid (constrained=True) */
  int32_t pad_;
  double d_oh, d_hh;             /* nm; <= 0: TIP3P 0.09572 / 0.15139 */
} gamd_md_options;

/* replaces: the constructors of HackNoseHooverIntegrator / HackHalfNoseHooverIntegrator / HackLangevinIntegrator
 * (code/hack_integrator.py:190-277, 344-425, 93-165) for the device-resident loop: afterwards gamd_md_run runs that
 * thermostat's two half-step programs (and ConstrainPositions / ConstrainVelocities when rigid_water) instead of
 * plain velocity Verlet.  Resets the chain (xi = vxi = 0, G = -frequency^2) and the variate counter. */
int gamd_md_configure(gamd_ctx* ctx, const gamd_md_options* opt);
/* replaces: propagateNHC (code/hack_integrator.py:289-316 / :454-481): KE2 = sum m v^2, chain, v *= scale;
 * bath != 0 also evaluates computeEnergies (:483-493).  d_state NULL = the ctx's own chain. */
int gamd_nhc_init_state(gamd_ctx* ctx, gamd_nhc_state* d_state, int32_t chain_length, int32_t num_mts, int32_t num_ys,
                        double kT, double frequency, double ndf, void* stream);
int gamd_nhc_propagate(gamd_ctx* ctx, gamd_nhc_state* d_state, double* d_v, const double* d_mass, int64_t n_atoms,
                       double dt, int32_t bath, void* stream);
int gamd_nhc_get_state(gamd_ctx* ctx, const gamd_nhc_state* d_state, gamd_nhc_state* h_out, void* stream);
int gamd_nhc_set_state(gamd_ctx* ctx, gamd_nhc_state* d_state, const gamd_nhc_state* h_in, void* stream);
/* replaces: HackLangevinIntegrator step body (code/hack_integrator.py:141-165), no constraints:
 * v += dt/2 f/m; x += dt/2 v; v = a v + b sqrt(kT/m) gaussian; x += dt/2 v.
 * d_gaussian fp64 [n,3] standard normals, or NULL: Philox4x32-10 keyed by (seed, step counter, atom). */
int gamd_langevin_first_half(gamd_ctx* ctx, double* d_x, double* d_v, const double* d_f, const double* d_mass,
                             int64_t n_atoms, double dt, double kT, double friction, const double* d_gaussian,
                             void* stream);
/* replaces: the collision stage of HackAndersenVVIntegrator (code/hack_integrator.py:66-68), per DOF:
 * collision = step(p_collision - uniform); v = (1 - collision) v + collision sqrt(kT/m) gaussian.
 * d_uniform / d_gaussian fp64 [n,3] or both NULL (Philox). */
int gamd_andersen_collide(gamd_ctx* ctx, double* d_v, const double* d_mass, int64_t n_atoms, double kT,
                          double p_collision, const double* d_uniform, const double* d_gaussian, void* stream);
/* replaces: OpenMM ConstrainPositions / ConstrainVelocities for rigid 3-site water (addConstrainPositions /
 * addConstrainVelocities, code/hack_integrator.py:84-86, 146-165, 274-277, 421-422): analytic SETTLE.
 * Atoms are [O,H,H] triplets.  d_x0: positions that satisfy the constraints (start of the step); d_x: the
 * unconstrained new positions, constrained in place; d_v (may be NULL): v += (x_constrained - x)/dt_corr. */
int gamd_settle_positions(gamd_ctx* ctx, const double* d_x0, double* d_x, double* d_v, const double* d_mass,
                          int64_t n_mol, double dt_corr, double d_oh, double d_hh, void* stream);
int gamd_settle_velocities(gamd_ctx* ctx, const double* d_x, double* d_v, const double* d_mass, int64_t n_mol,
                           void* stream);

/* ---- TIP4P virtual sites ------------------------------------------------------------------------ */
/* replaces: the M-site strip of WaterDataNew (code/train_utils.py:58-64: the model sees rows with
 * arange % 4 < 3 of a 4-site water box [O,H,H,M]) and OpenMM's average3 virtual-site placement of the
 * TIP4P-Ew system (dataset/generate_tip4p_data.py:55-57).  d_x4 fp64 [4*n_mol,3], d_x3 fp64 [3*n_mol,3].
 * unstrip copies the three massive sites back and sets the M row to w_o*O + w_h*(H1+H2) (place_m = 1, positions)
 * or to zero (place_m = 0, forces / velocities of the massless site). */
int gamd_tip4p_strip(gamd_ctx* ctx, const double* d_x4, double* d_x3, int64_t n_mol, void* stream);
int gamd_tip4p_unstrip(gamd_ctx* ctx, const double* d_a3, double* d_a4, int64_t n_mol, double w_o, double w_h,
                       int32_t place_m, void* stream);

/* ---- spatial domain decomposition (multi-GPU, one ctx per rank) ------------------------------- */
/* The reference has no counterpart (its MD loop is single-GPU); these entry points
 * split gamd_compute_forces at the points where a rank needs data of atoms it does not own:
 *   gamd_dd_begin   local atoms = n_own owned + (n_local - n_own) halo atoms (neighbours only: no CSR row);
 *                   neighbor search with the GLOBAL periodic box, edge encoder, layer-0 node prologue
 *   gamd_dd_layer   message-passing layer `layer` + node update (decoder after the last layer)
 *   gamd_dd_split_tiles / gamd_dd_layer_edges / gamd_dd_layer_nodes   the same layer in pieces, so that the halo
 *                   exchange of layer l-1's rows can run (on another stream) underneath most of layer l's edge work:
 *                   split_tiles (once per step, after gamd_dd_begin) sorts the 128-edge tiles of the CSR into
 *                   "interior" (no halo source) and "boundary"; layer_edges(which = 0) runs the edge chain of
 *                   nn_module.py:135-142 on the interior tiles and needs no halo row, layer_edges(which = 1) the
 *                   rest (after gamd_dd_unpack_rows); layer_nodes finishes the layer (nn_module.py:143-148).
 *                   which = -1 runs every tile.  Tensor-core precisions only (GAMD_EUNSUPPORTED for fp32).
 *   gamd_dd_pack_rows / gamd_dd_unpack_rows   rows [hn | src_affine(hn)] (2 x 128 fp32) of the listed owned
 *                   atoms -> send buffer; received rows -> the halo atoms first_local_idx .. +n-1.  The caller
 *                   moves the buffers between ranks (NCCL send/recv over NVLink) after layers 0 .. L-2.
 *   gamd_dd_finish  de-normalised forces of the owned atoms (local order) [+ second half-kick, + kinetic energy]
 * d_pos fp64 [n_local,3] Angstrom, owned atoms first. */
int gamd_dd_begin(gamd_ctx* ctx, const double* d_pos, int64_t n_own, int64_t n_local, const double h_box[3],
                  float cutoff, const float* d_feat, void* stream);
int gamd_dd_layer(gamd_ctx* ctx, int32_t layer, void* stream);
int gamd_dd_split_tiles(gamd_ctx* ctx, void* stream);
int gamd_dd_layer_edges(gamd_ctx* ctx, int32_t layer, int32_t which, void* stream);
int gamd_dd_layer_nodes(gamd_ctx* ctx, int32_t layer, void* stream);
int gamd_dd_pack_rows(gamd_ctx* ctx, const int32_t* d_local_idx, int64_t n, float* d_out, void* stream);
int gamd_dd_unpack_rows(gamd_ctx* ctx, int64_t first_local_idx, int64_t n, const float* d_in, void* stream);
int gamd_dd_finish(gamd_ctx* ctx, double* d_force, double* d_v, const double* d_mass, double dt, double* d_ke,
                   void* stream);
/* Halo exchange over NVLink / NVSwitch PEER MEMORY instead of NCCL send/recv (one process per GPU; the buffers are
 * shared with CUDA IPC):
 *   gamd_peer_alloc  device buffer of this rank that neighbours may write (zeroed) + its 64-byte IPC handle
 *   gamd_peer_open   map a neighbour's buffer from its handle (exchanged by the caller, e.g. all_gather)
 *   gamd_dd_push_rows  the rows [hn | src_affine(hn)] of the listed owned atoms are written by the pack kernel
 *                    STRAIGHT INTO the neighbour's buffer (no staging copy, no collective); a system-scope release
 *                    store of `seq` to the neighbour's flag follows in stream order
 *   gamd_dd_arm_push   FUSED form of the above: the next gamd_dd_layer / gamd_dd_layer_nodes call stores the rows of the
 *                    owned atoms with d_slot_left[i] / d_slot_right[i] >= 0 (i = local atom index < n_own) into slot
 *                    d_slot_*[i] of the left / right neighbour's buffer FROM THE NODE KERNEL'S EPILOGUE, as it produces
 *                    them (no pack kernel, the transfer runs under the kernel's own GEMMs); either side may be NULL.
 *                    The caller then publishes with gamd_dd_push_rows(ctx, NULL, 0, NULL, flag, seq, stream)
 *   gamd_dd_push_bytes the same for a plain device buffer (halo positions); n_bytes must be a multiple of 16
 *   gamd_dd_wait_flag  stream-ordered wait until this rank's own flag has reached `seq` (bounded: a neighbour that never
 *                    delivers raises GAMD_ESTATE at the next gamd_check_async_errors instead of hanging the GPU) */
int gamd_peer_alloc(gamd_ctx* ctx, int64_t n_bytes, void** d_ptr, uint8_t h_handle[64]);
int gamd_peer_open(gamd_ctx* ctx, const uint8_t h_handle[64], void** d_ptr);
int gamd_dd_push_rows(gamd_ctx* ctx, const int32_t* d_local_idx, int64_t n, float* d_remote_rows,
                      unsigned long long* d_remote_flag, uint64_t seq, void* stream);
int gamd_dd_arm_push(gamd_ctx* ctx, const int32_t* d_slot_left, float* d_remote_rows_left, const int32_t* d_slot_right,
                     float* d_remote_rows_right, int64_t n_own);
int gamd_dd_push_bytes(gamd_ctx* ctx, const void* d_src, int64_t n_bytes, void* d_remote_dst,
                       unsigned long long* d_remote_flag, uint64_t seq, void* stream);
int gamd_dd_wait_flag(gamd_ctx* ctx, const unsigned long long* d_flag, uint64_t seq, void* stream);

/* synchronises `stream` and reports errors raised asynchronously on the device since the
 * last check (GAMD_ECAPACITY: edge capacity exceeded; GAMD_EINVAL: edge list not sorted by
 * centre or ids out of range).  The *_host entry points call it themselves. */
int gamd_check_async_errors(gamd_ctx* ctx, void* stream);

/* ---- introspection (tests, profiling) ---------------------------------------------------- */
/* device pointers into ctx scratch, valid until the next gamd_reserve: names
 * "row_ptr","col_idx","edge_dst","perm","n_edges","pos_sorted","e_emb","h","agg","pred". */
int gamd_debug_ptr(gamd_ctx* ctx, const char* name, void** d_ptr, int64_t* n_bytes);
/* candidate-list reuse of the fused force path (code/graph_utils.py:21-25: jax-md rebuilds its neighbor candidates
 * only when an atom has moved more than half the skin, dr_threshold = cutoff / 6, and applies the exact mask every
 * step): number of candidate rebuilds and of searches since gamd_reserve.  Synchronises `stream`. */
int gamd_neighbor_stats(gamd_ctx* ctx, int64_t* n_rebuilds, int64_t* n_searches, void* stream);
/* the atoms behind the indices changed (hand-over between domains, new halo membership): the saved candidate rows
 * describe other atoms and must be rebuilt by the next search.  gamd_dd_begin reuses candidates only between two
 * calls of this function; the single-domain entry points detect a new system by its size / box / cutoff. */
int gamd_neighbor_invalidate(gamd_ctx* ctx);
/* number of kernel launches issued by this ctx since creation */
int64_t gamd_launch_count(const gamd_ctx* ctx);
/* per-stage CUDA-event timers on the launching stream.  Stages: "neighbor", "edge_encode",
 * "mp_edge", "node_update", "integrate".  gamd_profile_read synchronises the device, returns
 * the device time and the number of timed launches accumulated since the previous read of
 * that stage, and resets them. */
int gamd_profile_enable(gamd_ctx* ctx, int32_t on);
int gamd_profile_read(gamd_ctx* ctx, const char* stage, double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* GAMD_B200_H */
