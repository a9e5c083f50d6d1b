// Stage 3: integrator hook kernels (fp64 state, OpenMM units: nm, ps, Da, kJ/mol/nm).
//
// Restates the per-DOF programs of code/hack_integrator.py:
//   first half   :271-277   v += 0.5*dt*force_last/m ; x += dt*v      (no constraints: the
//                            v += (x-x1)/dt correction is identically zero)
//   second half  :171-178 / :421-422   v += (dt/2)*gnn_force/m
// and the force de-normalisation of code/LJ/train_network_lj.py:128-131, :153-155
// (fp32 prediction * sqrt(var) + mean, in double), fused with the second half-kick and the
// un-permutation from cell-sorted order back to the caller's atom order.
#include "common.cuh"

__global__ void k_vv_first(double* __restrict__ x, double* __restrict__ v, const double* __restrict__ f,
                           const double* __restrict__ mass, int64_t n, double dt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double m = mass[i];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double vv = v[3 * i + k] + 0.5 * dt * f[3 * i + k] / m;
    v[3 * i + k] = vv;
    x[3 * i + k] = x[3 * i + k] + dt * vv;
  }
}

__global__ void k_vv_second(double* __restrict__ v, const double* __restrict__ f, const double* __restrict__ mass,
                            int64_t n, double dt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double m = mass[i];
#pragma unroll
  for (int k = 0; k < 3; k++) v[3 * i + k] = v[3 * i + k] + (dt / 2) * f[3 * i + k] / m;
}

// pred (fp32, sorted order) -> f (fp64, caller order) [+ second half-kick] [+ kinetic energy]
__global__ void k_denorm_scatter(const float* __restrict__ pred, const int* __restrict__ perm, int64_t n, int64_t n_own,
                                 double sigma, double mean, double* __restrict__ f, double* __restrict__ v,
                                 const double* __restrict__ mass, double dt, double* __restrict__ ke,
                                 const int* __restrict__ ke_slot, double* __restrict__ ke2_acc) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double local = 0.0;
  int64_t i = s < n ? (perm ? perm[s] : s) : n_own;
  if (i < n_own) {   // rows of halo atoms (caller index >= n_own) carry no force
    double m = v ? mass[i] : 1.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      double F = (double)pred[3 * s + k] * sigma + mean;
      f[3 * i + k] = F;
      if (v) {
        double vv = v[3 * i + k] + (dt / 2) * F / m;
        v[3 * i + k] = vv;
        local += 0.5 * m * vv * vv;
      }
    }
  }
  if (ke || ke2_acc) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ double sw[8];
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sw[w];
      if (ke) atomicAdd(ke + (ke_slot ? *ke_slot : 0), t);
      if (ke2_acc) atomicAdd(ke2_acc, 2.0 * t);      // sum m v^2 for the thermostat chain that follows the kick
    }
  }
}

int integ_first_half(gamd_ctx* ctx, double* x, double* v, const double* f, const double* mass, int64_t n, double dt,
                     cudaStream_t st) {
  prof_mark(ctx, "integrate", st);
  k_vv_first<<<ceil_div(n, 256), 256, 0, st>>>(x, v, f, mass, n, dt);
  GAMD_LAUNCH_CHECK();
  prof_mark(ctx, "integrate", st);
  return 0;
}

int integ_second_half(gamd_ctx* ctx, double* v, const double* f, const double* mass, int64_t n, double dt,
                      cudaStream_t st) {
  k_vv_second<<<ceil_div(n, 256), 256, 0, st>>>(v, f, mass, n, dt);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int integ_denorm_scatter(gamd_ctx* ctx, const int* perm, double* f_out, double* v, const double* mass, double dt,
                         int64_t n, double* ke_out, cudaStream_t st, int64_t n_own, const int* ke_slot, double* ke2_acc) {
  if (n_own < 0) n_own = n;
  // ke_slot != nullptr: ke_out is a per-step trace zeroed by the caller, *ke_slot selects the entry (CUDA-graph replay)
  if (ke_out && !ke_slot) GAMD_CUDA(cudaMemsetAsync(ke_out, 0, sizeof(double), st));
  prof_mark(ctx, "integrate", st);
  k_denorm_scatter<<<ceil_div(n, 256), 256, 0, st>>>(ctx->pred, perm, n, n_own, sqrt(ctx->scaler_var), ctx->scaler_mean,
                                                     f_out, v, mass, dt, ke_out, ke_slot, ke2_acc);
  GAMD_LAUNCH_CHECK();
  prof_mark(ctx, "integrate", st);
  return 0;
}

__global__ void k_inc_counter(int* c) { *c += 1; }

int integ_inc_counter(gamd_ctx* ctx, int* counter, cudaStream_t st) {
  k_inc_counter<<<1, 1, 0, st>>>(counter);
  GAMD_LAUNCH_CHECK();
  return 0;
}

// ---- TIP4P virtual sites -------------------------------------------------
// 4-site molecules [O, H, H, M]: the model sees only O, H, H (code/train_utils.py:58-64: arange % 4 < 3);
// the massless M site is re-placed from them (OpenMM "average3" virtual site).
__global__ void k_tip4p_strip(const double* __restrict__ x4, double* __restrict__ x3, int64_t n_mol) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (molecule, site<3, component)
  if (i >= n_mol * 9) return;
  int64_t mol = i / 9;
  int r = (int)(i - mol * 9);
  x3[i] = x4[mol * 12 + r];
}

__global__ void k_tip4p_unstrip(const double* __restrict__ a3, double* __restrict__ a4, int64_t n_mol, double wo,
                                double wh, int place_m) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (molecule, component)
  if (i >= n_mol * 3) return;
  int64_t mol = i / 3;
  int k = (int)(i - mol * 3);
  double o = a3[mol * 9 + k], h1 = a3[mol * 9 + 3 + k], h2 = a3[mol * 9 + 6 + k];
  a4[mol * 12 + k] = o;
  a4[mol * 12 + 3 + k] = h1;
  a4[mol * 12 + 6 + k] = h2;
  a4[mol * 12 + 9 + k] = place_m ? wo * o + wh * h1 + wh * h2 : 0.0;
}

int integ_tip4p_strip(gamd_ctx* ctx, const double* x4, double* x3, int64_t n_mol, cudaStream_t st) {
  k_tip4p_strip<<<ceil_div(n_mol * 9, 256), 256, 0, st>>>(x4, x3, n_mol);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int integ_tip4p_unstrip(gamd_ctx* ctx, const double* a3, double* a4, int64_t n_mol, double wo, double wh, int place_m,
                        cudaStream_t st) {
  k_tip4p_unstrip<<<ceil_div(n_mol * 3, 256), 256, 0, st>>>(a3, a4, n_mol, wo, wh, place_m);
  GAMD_LAUNCH_CHECK();
  return 0;
}
