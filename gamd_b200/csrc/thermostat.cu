// Thermostats and rigid-water constraints of the integrator hook, device resident (fp64, OpenMM units).
//
// Restates the step programs of code/hack_integrator.py that the shipped drivers run (NVT):
//   propagateNHC                     :289-316 / :454-481   Nose-Hoover chain, Yoshida-Suzuki n_ys in {1,3,5}, n_c sub-steps
//   HackNoseHooverIntegrator         :267-277   propagateNHC; v += 0.5 dt f/m; x += dt v; constrain; v += (x - x1)/dt
//   HackHalfNoseHooverIntegrator     :419-425   v += 0.5 dt f/m; constrain v; propagateNHC; bath energies :483-493
//   HackLangevinIntegrator           :141-165   B (dt/2), A (dt/2), O, A (dt/2), each followed by its constraint stage
//   HackAndersenVVIntegrator         :66-86     per-particle collisions, then velocity Verlet with test1 / test2
// and, for the water drivers (constrained=True, code/water/test_script/test_nosehoover.py:33-37), OpenMM's
// ConstrainPositions / ConstrainVelocities for rigid 3-site water: the analytic SETTLE of Miyamoto & Kollman
// (J. Comput. Chem. 13, 952, 1992) for positions and the closed-form rigid-triangle velocity projection.
//
// The chain is a handful of scalars: one thread propagates it from a device-resident kinetic-energy sum, the velocity
// scaling is fused into the kick kernels - nothing synchronises the host (the round-1 host loop paid a .item() per
// half step).  Gaussian / uniform variates come from Philox4x32-10 keyed by (seed, step, atom) unless the caller
// injects them (parity tests: OpenMM's generator cannot be reproduced).
#include "common.cuh"
#include <cmath>

namespace {

__constant__ double c_ys1[1] = {1.0};
__constant__ double c_ys3[3] = {0.8289815435887510, -0.6579630871775020, 0.8289815435887510};
__constant__ double c_ys5[5] = {0.2967324292201065, 0.2967324292201065, -0.1869297168804260, 0.2967324292201065,
                                0.2967324292201065};

// ---- reductions -----------------------------------------------------------------------------------------------
// acc[0] += sum_i m_i |v_i|^2  (KE2 of propagateNHC: addComputeSum("KE2", "m*v^2"))
__global__ void k_ke2(const double* __restrict__ v, const double* __restrict__ mass, int64_t n, double* __restrict__ acc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double local = 0.0;
  if (i < n) {
    const double m = mass[i];
    local = m * (v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double sw[8];
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sw[w];
    atomicAdd(acc, t);
  }
}

// ---- Nose-Hoover chain ------------------------------------------------------------------------------------------
// One thread.  ke2_in: sum m v^2 of the velocities the chain acts on.  Writes the velocity scale factor to
// st->scale, leaves st->ke2 = scale^2 * ke2_in (the sum AFTER scaling, exact up to rounding), optionally the bath
// energies (second half) and 0.5 * ke2 into ke_out[*ke_slot].
__global__ void k_nhc_chain(gamd_nhc_state* __restrict__ st, const double* __restrict__ ke2_src, int zero_src, double dt,
                            int bath, double* __restrict__ ke_out, const int* __restrict__ ke_slot) {
  if (threadIdx.x || blockIdx.x) return;
  const int M = st->M, n_c = st->n_c, n_ys = st->n_ys;
  const double kT = st->kT, ndf = st->ndf;
  double ke2 = *ke2_src;
  if (zero_src) *const_cast<double*>(ke2_src) = 0.0;
  double scale = 1.0;
  if (M > 0) {
    double xi[GAMD_NHC_MAX], vxi[GAMD_NHC_MAX], G[GAMD_NHC_MAX], Q[GAMD_NHC_MAX];
    for (int j = 0; j < M; j++) {
      xi[j] = st->xi[j];
      vxi[j] = st->vxi[j];
      G[j] = st->G[j];
      Q[j] = j == 0 ? ndf * st->Qbase : st->Qbase;       // addComputeGlobal("Q0", "ndf*Q"); Q_i = Q  (:283-287)
    }
    const double* w = n_ys == 1 ? c_ys1 : (n_ys == 3 ? c_ys3 : c_ys5);
    G[0] = (ke2 - ndf * kT) / Q[0];
    for (int nc = 0; nc < n_c; nc++)
      for (int ys = 0; ys < n_ys; ys++) {
        const double wdt = w[ys] * dt / n_c;
        vxi[M - 1] = vxi[M - 1] + 0.25 * wdt * G[M - 1];
        for (int j = M - 2; j >= 0; j--) {
          const double aa = exp(-0.125 * wdt * vxi[j + 1]);
          vxi[j] = aa * (aa * vxi[j] + 0.25 * wdt * G[j]);
        }
        const double aa = exp(-0.5 * wdt * vxi[0]);
        scale = scale * aa;
        for (int j = 0; j < M; j++) xi[j] = xi[j] + 0.5 * wdt * vxi[j];
        G[0] = (scale * scale * ke2 - ndf * kT) / Q[0];
        for (int j = 0; j < M - 1; j++) {
          const double ab = exp(-0.125 * wdt * vxi[j + 1]);
          vxi[j] = ab * (ab * vxi[j] + 0.25 * wdt * G[j]);
          G[j + 1] = (Q[j] * vxi[j] * vxi[j] - kT) / Q[j + 1];
        }
        vxi[M - 1] = vxi[M - 1] + 0.25 * wdt * G[M - 1];
      }
    double bke = 0.0, bpe = 0.0;
    for (int j = 0; j < M; j++) {
      st->xi[j] = xi[j];
      st->vxi[j] = vxi[j];
      st->G[j] = G[j];
      st->Q[j] = Q[j];
      bke += 0.5 * Q[j] * vxi[j] * vxi[j];
      bpe += j == 0 ? ndf * xi[0] : xi[j];
    }
    if (bath) {                                           // computeEnergies (:483-493)
      st->bathKE = bke;
      st->bathPE = kT * bpe;
    }
  }
  st->scale = scale;
  st->ke2_in = ke2;
  st->ke2 = scale * scale * ke2;
  if (ke_out) ke_out[ke_slot ? *ke_slot : 0] = 0.5 * st->ke2;
}

__global__ void k_scale_v(double* __restrict__ v, int64_t n3, const gamd_nhc_state* __restrict__ st) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n3) v[i] *= st->scale;
}

// first half with the chain's scale fused: v = scale v + 0.5 dt f/m; x += dt v
__global__ void k_vv_first_scaled(double* __restrict__ x, double* __restrict__ v, const double* __restrict__ f,
                                  const double* __restrict__ mass, int64_t n, double dt,
                                  const gamd_nhc_state* __restrict__ st) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double m = mass[i], s = st->scale;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double vs = s * v[3 * i + k];
    const double vv = vs + 0.5 * dt * f[3 * i + k] / m;
    v[3 * i + k] = vv;
    x[3 * i + k] = x[3 * i + k] + dt * vv;
  }
}

// ---- Philox4x32-10 ------------------------------------------------------------------------------------------------
struct Philox {
  uint32_t c[4];
};
__device__ __forceinline__ Philox philox(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox p;
  p.c[0] = c0; p.c[1] = c1; p.c[2] = c2; p.c[3] = c3;
  return p;
}
__device__ __forceinline__ double u01(uint32_t a) { return ((double)a + 0.5) * (1.0 / 4294967296.0); }   // (0, 1)
// three standard normals for (stream, step, atom): two Box-Muller pairs
__device__ __forceinline__ void gauss3(uint64_t seed, uint64_t step, uint64_t atom, uint32_t stream, double (&g)[3]) {
  const Philox p = philox(seed, atom, (step << 8) | stream);
  const double r0 = sqrt(-2.0 * log(u01(p.c[0]))), r1 = sqrt(-2.0 * log(u01(p.c[2])));
  double s0, c0, s1, c1;
  sincospi(2.0 * u01(p.c[1]), &s0, &c0);
  sincospi(2.0 * u01(p.c[3]), &s1, &c1);
  g[0] = r0 * c0;
  g[1] = r0 * s0;
  g[2] = r1 * c1;
}

// ---- Langevin, first integrator (HackLangevinIntegrator: B, A, O, A), no constraints -----------------------------
__global__ void k_langevin_first(double* __restrict__ x, double* __restrict__ v, const double* __restrict__ f,
                                 const double* __restrict__ mass, int64_t n, double dt, double kT, double a, double b,
                                 const double* __restrict__ gaussian, uint64_t seed,
                                 const unsigned long long* __restrict__ step_ctr) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double m = mass[i], sigma = sqrt(kT / m);
  double g[3];
  if (gaussian) {
    g[0] = gaussian[3 * i]; g[1] = gaussian[3 * i + 1]; g[2] = gaussian[3 * i + 2];
  } else {
    gauss3(seed, *step_ctr, (uint64_t)i, 1u, g);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double vv = v[3 * i + k] + (dt / 2) * f[3 * i + k] / m;        // B
    double xx = x[3 * i + k] + ((dt / 2) * vv);                     // A
    vv = (a * vv) + (b * sigma * g[k]);                             // O
    xx = xx + ((dt / 2) * vv);                                      // A
    v[3 * i + k] = vv;
    x[3 * i + k] = xx;
  }
}

// ---- Andersen collisions: collision = step(p - uniform); v = (1 - collision) v + collision sigma_v gaussian -------
// (per DOF, exactly as the per-DOF program is written: every component draws its own uniform)
__global__ void k_andersen(double* __restrict__ v, const double* __restrict__ mass, int64_t n, double kT, double p_coll,
                           const double* __restrict__ uniform, const double* __restrict__ gaussian, uint64_t seed,
                           const unsigned long long* __restrict__ step_ctr) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double sigma = sqrt(kT / mass[i]);
  double g[3], u[3];
  if (gaussian && uniform) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      g[k] = gaussian[3 * i + k];
      u[k] = uniform[3 * i + k];
    }
  } else {
    gauss3(seed, *step_ctr, (uint64_t)i, 2u, g);
    const Philox p = philox(seed, (uint64_t)i, (*step_ctr << 8) | 3u);
    u[0] = u01(p.c[0]); u[1] = u01(p.c[1]); u[2] = u01(p.c[2]);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double coll = (p_coll - u[k] >= 0.0) ? 1.0 : 0.0;        // OpenMM step(x): 0 if x < 0, 1 otherwise
    v[3 * i + k] = (1.0 - coll) * v[3 * i + k] + coll * sigma * g[k];
  }
}

__global__ void k_inc_u64(unsigned long long* c) { *c += 1ull; }

// ---- SETTLE ---------------------------------------------------------------------------------------------------------
struct V3 {
  double x, y, z;
};
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ V3 ld3(const double* p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ void st3(double* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

// positions of one [O,H,H] molecule: a0* satisfy the constraints, a1* are the unconstrained new positions;
// returns the constrained new positions (displacements along the OLD bond vectors, centre of mass unchanged)
__device__ __forceinline__ void settle_one(V3 a00, V3 a01, V3 a02, V3& a10, V3& a11, V3& a12, double m0, double m1, double m2,
                                           double d_oh, double d_hh) {
  const V3 xp0 = a10 - a00, xp1 = a11 - a01, xp2 = a12 - a02;
  const V3 xb0 = a01 - a00, xc0 = a02 - a00;
  const double inv_mt = 1.0 / (m0 + m1 + m2);
  const V3 xcom = (xp0 * m0 + (xb0 + xp1) * m1 + (xc0 + xp2) * m2) * inv_mt;
  const V3 xa1 = xp0 - xcom, xb1 = xb0 + xp1 - xcom, xc1 = xc0 + xp2 - xcom;
  const V3 akz = cross(xb0, xc0), akx = cross(xa1, akz), aky = cross(akz, akx);
  const V3 t1 = akx * (1.0 / sqrt(dot(akx, akx))), t2 = aky * (1.0 / sqrt(dot(aky, aky))), t3 = akz * (1.0 / sqrt(dot(akz, akz)));
  const double xb0d = dot(t1, xb0), yb0d = dot(t2, xb0), xc0d = dot(t1, xc0), yc0d = dot(t2, xc0);
  const double za1d = dot(t3, xa1);
  const double xb1d = dot(t1, xb1), yb1d = dot(t2, xb1), zb1d = dot(t3, xb1);
  const double xc1d = dot(t1, xc1), yc1d = dot(t2, xc1), zc1d = dot(t3, xc1);
  const double rc = 0.5 * d_hh;
  double rb = sqrt(d_oh * d_oh - rc * rc);
  const double ra = rb * (m1 + m2) * inv_mt;
  rb -= ra;
  const double sinphi = za1d / ra, cosphi = sqrt(1.0 - sinphi * sinphi);
  const double sinpsi = (zb1d - zc1d) / (2.0 * rc * cosphi), cospsi = sqrt(1.0 - sinpsi * sinpsi);
  const double ya2d = ra * cosphi;
  double xb2d = -rc * cospsi;
  const double yb2d = -rb * cosphi - rc * sinpsi * sinphi, yc2d = -rb * cosphi + rc * sinpsi * sinphi;
  const double xb2d2 = xb2d * xb2d;
  const double hh2 = 4.0 * xb2d2 + (yb2d - yc2d) * (yb2d - yc2d) + (zb1d - zc1d) * (zb1d - zc1d);
  const double deltx = 2.0 * xb2d + sqrt(4.0 * xb2d2 - hh2 + d_hh * d_hh);
  xb2d -= deltx * 0.5;
  const double alpha = xb2d * (xb0d - xc0d) + yb0d * yb2d + yc0d * yc2d;
  const double beta = xb2d * (yc0d - yb0d) + xb0d * yb2d + xc0d * yc2d;
  const double gamma = xb0d * yb1d - xb1d * yb0d + xc0d * yc1d - xc1d * yc0d;
  const double al2be2 = alpha * alpha + beta * beta;
  const double sintheta = (alpha * gamma - beta * sqrt(al2be2 - gamma * gamma)) / al2be2;
  const double costheta = sqrt(1.0 - sintheta * sintheta);
  const double xa3d = -ya2d * sintheta, ya3d = ya2d * costheta, za3d = za1d;
  const double xb3d = xb2d * costheta - yb2d * sintheta, yb3d = xb2d * sintheta + yb2d * costheta, zb3d = zb1d;
  const double xc3d = -xb2d * costheta - yc2d * sintheta, yc3d = -xb2d * sintheta + yc2d * costheta, zc3d = zc1d;
  const V3 xa3 = t1 * xa3d + t2 * ya3d + t3 * za3d;
  const V3 xb3 = t1 * xb3d + t2 * yb3d + t3 * zb3d;
  const V3 xc3 = t1 * xc3d + t2 * yc3d + t3 * zc3d;
  a10 = a00 + xcom + xa3;
  a11 = a01 + xcom + xb3 - xb0;
  a12 = a02 + xcom + xc3 - xc0;
}

// remove the relative velocity along the three bonds with impulses along them (3 x 3 linear solve, Cramer)
__device__ __forceinline__ void settle_vel_one(V3 a, V3 b, V3 c, V3& va, V3& vb, V3& vc, double ma, double mb, double mc) {
  V3 eab = b - a, ebc = c - b, eca = a - c;
  eab = eab * (1.0 / sqrt(dot(eab, eab)));
  ebc = ebc * (1.0 / sqrt(dot(ebc, ebc)));
  eca = eca * (1.0 / sqrt(dot(eca, eca)));
  const double vab = dot(vb - va, eab), vbc = dot(vc - vb, ebc), vca = dot(va - vc, eca);
  const double cab_bc = dot(eab, ebc), cab_ca = dot(eab, eca), cbc_ca = dot(ebc, eca);
  // A t = -[vab, vbc, vca]
  const double A00 = -(1.0 / ma + 1.0 / mb), A01 = cab_bc / mb, A02 = cab_ca / ma;
  const double A10 = cab_bc / mb, A11 = -(1.0 / mb + 1.0 / mc), A12 = cbc_ca / mc;
  const double A20 = cab_ca / ma, A21 = cbc_ca / mc, A22 = -(1.0 / mc + 1.0 / ma);
  const double r0 = -vab, r1 = -vbc, r2 = -vca;
  const double det = A00 * (A11 * A22 - A12 * A21) - A01 * (A10 * A22 - A12 * A20) + A02 * (A10 * A21 - A11 * A20);
  const double tab = (r0 * (A11 * A22 - A12 * A21) - A01 * (r1 * A22 - A12 * r2) + A02 * (r1 * A21 - A11 * r2)) / det;
  const double tbc = (A00 * (r1 * A22 - A12 * r2) - r0 * (A10 * A22 - A12 * A20) + A02 * (A10 * r2 - r1 * A20)) / det;
  const double tca = (A00 * (A11 * r2 - r1 * A21) - A01 * (A10 * r2 - r1 * A20) + r0 * (A10 * A21 - A11 * A20)) / det;
  va = va + (eab * tab - eca * tca) * (1.0 / ma);
  vb = vb + (ebc * tbc - eab * tab) * (1.0 / mb);
  vc = vc + (eca * tca - ebc * tbc) * (1.0 / mc);
}

// x0: positions satisfying the constraints; x: unconstrained new positions (updated in place); v (optional):
// v += (x_constrained - x_unconstrained) / dt_corr   (hack_integrator.py:277 / :156, :164)
__global__ void k_settle_pos(const double* __restrict__ x0, double* __restrict__ x, double* __restrict__ v,
                             const double* __restrict__ mass, int64_t n_mol, double dt_corr, double d_oh, double d_hh) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_mol) return;
  const double* p0 = x0 + 9 * k;
  double* p1 = x + 9 * k;
  V3 b0 = ld3(p1), b1 = ld3(p1 + 3), b2 = ld3(p1 + 6);
  const V3 u0 = b0, u1 = b1, u2 = b2;
  settle_one(ld3(p0), ld3(p0 + 3), ld3(p0 + 6), b0, b1, b2, mass[3 * k], mass[3 * k + 1], mass[3 * k + 2], d_oh, d_hh);
  st3(p1, b0); st3(p1 + 3, b1); st3(p1 + 6, b2);
  if (v) {
    double* q = v + 9 * k;
    const double s = 1.0 / dt_corr;
    st3(q, ld3(q) + (b0 - u0) * s);
    st3(q + 3, ld3(q + 3) + (b1 - u1) * s);
    st3(q + 6, ld3(q + 6) + (b2 - u2) * s);
  }
}

// first half of the rigid-water step fused per molecule: [scale] kick, drift, SETTLE, velocity correction
__global__ void k_vv_first_rigid(double* __restrict__ x, double* __restrict__ v, const double* __restrict__ f,
                                 const double* __restrict__ mass, int64_t n_mol, double dt,
                                 const gamd_nhc_state* __restrict__ st, double d_oh, double d_hh) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_mol) return;
  const double s = st ? st->scale : 1.0;
  V3 x0[3], x1[3], vv[3];
  double m[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    m[a] = mass[3 * k + a];
    x0[a] = ld3(x + 9 * k + 3 * a);
    vv[a] = ld3(v + 9 * k + 3 * a) * s + ld3(f + 9 * k + 3 * a) * (0.5 * dt / m[a]);
    x1[a] = x0[a] + vv[a] * dt;
  }
  const V3 u0 = x1[0], u1 = x1[1], u2 = x1[2];
  settle_one(x0[0], x0[1], x0[2], x1[0], x1[1], x1[2], m[0], m[1], m[2], d_oh, d_hh);
  const double id = 1.0 / dt;
  vv[0] = vv[0] + (x1[0] - u0) * id;
  vv[1] = vv[1] + (x1[1] - u1) * id;
  vv[2] = vv[2] + (x1[2] - u2) * id;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    st3(x + 9 * k + 3 * a, x1[a]);
    st3(v + 9 * k + 3 * a, vv[a]);
  }
}

// ConstrainVelocities for rigid water + (optionally) the kinetic-energy sums of the constrained velocities
__global__ void k_settle_vel(const double* __restrict__ x, double* __restrict__ v, const double* __restrict__ mass,
                             int64_t n_mol, double* __restrict__ ke2_acc, double* __restrict__ ke_out,
                             const int* __restrict__ ke_slot) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double local = 0.0;
  if (k < n_mol) {
    const double ma = mass[3 * k], mb = mass[3 * k + 1], mc = mass[3 * k + 2];
    V3 va = ld3(v + 9 * k), vb = ld3(v + 9 * k + 3), vc = ld3(v + 9 * k + 6);
    settle_vel_one(ld3(x + 9 * k), ld3(x + 9 * k + 3), ld3(x + 9 * k + 6), va, vb, vc, ma, mb, mc);
    st3(v + 9 * k, va); st3(v + 9 * k + 3, vb); st3(v + 9 * k + 6, vc);
    local = ma * dot(va, va) + mb * dot(vb, vb) + mc * dot(vc, vc);
  }
  if (ke2_acc || ke_out) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ double sw[8];
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sw[w];
      if (ke2_acc) atomicAdd(ke2_acc, t);
      if (ke_out) atomicAdd(ke_out + (ke_slot ? *ke_slot : 0), 0.5 * t);
    }
  }
}

}  // namespace

// ---- host entry points (called from capi.cu) ------------------------------------------------------------------------
int thermo_alloc(gamd_ctx* ctx) {
  if (ctx->d_nhc) return 0;
  GAMD_CUDA(cudaMalloc(&ctx->d_nhc, sizeof(gamd_nhc_state) + 64));
  GAMD_CUDA(cudaMemset(ctx->d_nhc, 0, sizeof(gamd_nhc_state) + 64));
  ctx->d_ke2_acc = reinterpret_cast<double*>(reinterpret_cast<char*>(ctx->d_nhc) + sizeof(gamd_nhc_state));
  ctx->d_rng_ctr = reinterpret_cast<unsigned long long*>(ctx->d_ke2_acc + 1);
  return 0;
}

int thermo_ke2(gamd_ctx* ctx, const double* v, const double* mass, int64_t n, double* acc, cudaStream_t st) {
  GAMD_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  k_ke2<<<ceil_div(n, 256), 256, 0, st>>>(v, mass, n, acc);
  GAMD_LAUNCH_CHECK();
  return 0;
}

// chain from the accumulated sum (zeroed afterwards); the caller applies d_state->scale
int thermo_chain(gamd_ctx* ctx, gamd_nhc_state* d_state, double* ke2_acc, double dt, int bath, double* ke_out,
                 const int* ke_slot, cudaStream_t st) {
  k_nhc_chain<<<1, 32, 0, st>>>(d_state, ke2_acc, 1, dt, bath, ke_out, ke_slot);
  GAMD_LAUNCH_CHECK();
  return 0;
}

// chain from the cached post-scaling sum d_state->ke2 (valid when nothing touched v since the last chain + scaling)
int thermo_chain_cached(gamd_ctx* ctx, gamd_nhc_state* d_state, double dt, int bath, cudaStream_t st) {
  k_nhc_chain<<<1, 32, 0, st>>>(d_state, &d_state->ke2, 0, dt, bath, nullptr, nullptr);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int thermo_scale_v(gamd_ctx* ctx, const gamd_nhc_state* d_state, double* v, int64_t n, cudaStream_t st) {
  k_scale_v<<<ceil_div(3 * n, 256), 256, 0, st>>>(v, 3 * n, d_state);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int thermo_vv_first_scaled(gamd_ctx* ctx, double* x, double* v, const double* f, const double* mass, int64_t n, double dt,
                           cudaStream_t st) {
  k_vv_first_scaled<<<ceil_div(n, 256), 256, 0, st>>>(x, v, f, mass, n, dt, ctx->d_nhc);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int thermo_langevin_first(gamd_ctx* ctx, double* x, double* v, const double* f, const double* mass, int64_t n, double dt,
                          double kT, double friction, const double* gaussian, cudaStream_t st) {
  const double a = exp(-friction * dt), b = sqrt(1.0 - exp(-2.0 * friction * dt));
  k_langevin_first<<<ceil_div(n, 256), 256, 0, st>>>(x, v, f, mass, n, dt, kT, a, b, gaussian, ctx->md.seed, ctx->d_rng_ctr);
  GAMD_LAUNCH_CHECK();
  if (!gaussian) {
    k_inc_u64<<<1, 1, 0, st>>>(ctx->d_rng_ctr);
    GAMD_LAUNCH_CHECK();
  }
  return 0;
}

int thermo_andersen(gamd_ctx* ctx, double* v, const double* mass, int64_t n, double kT, double p_coll,
                    const double* uniform, const double* gaussian, cudaStream_t st) {
  k_andersen<<<ceil_div(n, 256), 256, 0, st>>>(v, mass, n, kT, p_coll, uniform, gaussian, ctx->md.seed, ctx->d_rng_ctr);
  GAMD_LAUNCH_CHECK();
  if (!(uniform && gaussian)) {
    k_inc_u64<<<1, 1, 0, st>>>(ctx->d_rng_ctr);
    GAMD_LAUNCH_CHECK();
  }
  return 0;
}

int thermo_settle_pos(gamd_ctx* ctx, const double* x0, double* x, double* v, const double* mass, int64_t n_mol,
                      double dt_corr, double d_oh, double d_hh, cudaStream_t st) {
  k_settle_pos<<<ceil_div(n_mol, 128), 128, 0, st>>>(x0, x, v, mass, n_mol, dt_corr, d_oh, d_hh);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int thermo_vv_first_rigid(gamd_ctx* ctx, double* x, double* v, const double* f, const double* mass, int64_t n_mol,
                          double dt, bool scaled, double d_oh, double d_hh, cudaStream_t st) {
  k_vv_first_rigid<<<ceil_div(n_mol, 128), 128, 0, st>>>(x, v, f, mass, n_mol, dt, scaled ? ctx->d_nhc : nullptr, d_oh, d_hh);
  GAMD_LAUNCH_CHECK();
  return 0;
}

int thermo_settle_vel(gamd_ctx* ctx, const double* x, double* v, const double* mass, int64_t n_mol, double* ke2_acc,
                      double* ke_out, const int* ke_slot, cudaStream_t st) {
  k_settle_vel<<<ceil_div(n_mol, 256), 256, 0, st>>>(x, v, mass, n_mol, ke2_acc, ke_out, ke_slot);
  GAMD_LAUNCH_CHECK();
  return 0;
}
