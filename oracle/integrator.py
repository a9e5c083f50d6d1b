"""Oracle integrator programs (numpy float64).  TEST INFRASTRUCTURE ONLY.

Restates the OpenMM ``CustomIntegrator`` step programs of ``code/hack_integrator.py`` in
OpenMM units (nm, ps, Da, kJ/mol; force kJ/mol/nm).  OpenMM itself is not installed here, so
parity with its interpreter is UNPINNED; tests validate these programs by analytic
properties (reversibility, harmonic-oscillator energy, equipartition).

  vv_first_half    HackNoseHooverIntegrator step body, hack_integrator.py:271-277
                   (``v+=0.5*dt*force_last/m; x+=dt*v``; with no constraints the
                   ``v+=(x-x1)/dt`` correction is zero), chain_length=0 -> plain VV (:206-207)
  vv_second_half   HackHalfVelocityIntegrator :171-178 / HackHalfNoseHooverIntegrator :419-425
  nhc_propagate    propagateNHC :289-316 (Yoshida-Suzuki n_ys in {1,3,5}, n_c multi-steps)
  langevin_first_half  HackLangevinIntegrator :141-165 (B, A/2, O, A/2)
  andersen_collide     HackAndersenVVIntegrator :66-68 (per-DOF collisions)
  bath_energies        computeEnergies :483-493
  run_nvt_nhc          the NVT driver loop of test_nosehoover.py:100-118 with both half-step programs
"""
import numpy as np

KB = 0.00831446261815324  # kJ/mol/K (openmmtools.constants.kB = BOLTZMANN * AVOGADRO)

YS_WEIGHTS = {
    1: [1.0],
    3: [0.8289815435887510, -0.6579630871775020, 0.8289815435887510],
    5: [0.2967324292201065, 0.2967324292201065, -0.1869297168804260, 0.2967324292201065, 0.2967324292201065],
}


def vv_first_half(x, v, f_last, m, dt):
    v = v + 0.5 * dt * f_last / m[:, None]
    x = x + dt * v
    return x, v


def vv_second_half(v, f_new, m, dt):
    return v + (dt / 2) * f_new / m[:, None]


def kinetic_energy(v, m):
    return 0.5 * float(np.sum(m[:, None] * v * v))


class NHCState:
    """Thermostat chain globals ``xi, vxi, G, Q`` (hack_integrator.py:249-261)."""

    def __init__(self, chain_length, kT, frequency, ndf):
        self.M = chain_length
        q = kT / frequency ** 2
        self.xi = np.zeros(chain_length)
        self.vxi = np.zeros(chain_length)
        self.G = np.full(chain_length, -frequency ** 2)
        self.Q = np.full(chain_length, q)
        if chain_length:
            self.Q[0] = ndf * q
        self.kT = kT
        self.ndf = ndf


def nhc_propagate(st, v, m, dt, n_c=5, n_ys=5):
    """``propagateNHC`` (hack_integrator.py:289-316); returns scaled velocities."""
    M = st.M
    if M == 0:
        return v
    w = YS_WEIGHTS[n_ys]
    scale = 1.0
    ke2 = float(np.sum(m[:, None] * v * v))
    st.G[0] = (ke2 - st.ndf * st.kT) / st.Q[0]
    for _ in range(n_c):
        for ys in range(n_ys):
            wdt = w[ys] * dt / n_c
            st.vxi[M - 1] = st.vxi[M - 1] + 0.25 * wdt * st.G[M - 1]
            for j in range(M - 2, -1, -1):
                aa = np.exp(-0.125 * wdt * st.vxi[j + 1])
                st.vxi[j] = aa * (aa * st.vxi[j] + 0.25 * wdt * st.G[j])
            aa = np.exp(-0.5 * wdt * st.vxi[0])
            scale = scale * aa
            for j in range(M):
                st.xi[j] = st.xi[j] + 0.5 * wdt * st.vxi[j]
            st.G[0] = (scale * scale * ke2 - st.ndf * st.kT) / st.Q[0]
            for j in range(M - 1):
                aa = np.exp(-0.125 * wdt * st.vxi[j + 1])
                st.vxi[j] = aa * (aa * st.vxi[j] + 0.25 * wdt * st.G[j])
                st.G[j + 1] = (st.Q[j] * st.vxi[j] * st.vxi[j] - st.kT) / st.Q[j + 1]
            st.vxi[M - 1] = st.vxi[M - 1] + 0.25 * wdt * st.G[M - 1]
    return scale * v


def langevin_first_half(x, v, f_last, m, dt, kT, gamma, gaussian):
    """HackLangevinIntegrator step body (hack_integrator.py:141-165), no constraints.
    ``gaussian`` is the [N,3] standard-normal draw OpenMM would make."""
    a = np.exp(-gamma * dt)
    b = np.sqrt(1 - np.exp(-2 * gamma * dt))
    sigma = np.sqrt(kT / m)[:, None]
    v = v + (dt / 2) * f_last / m[:, None]
    x = x + (dt / 2) * v
    v = a * v + b * sigma * gaussian
    x = x + (dt / 2) * v
    return x, v


def andersen_collide(v, m, kT, p_collision, uniform, gaussian):
    """``collision = step(p_collision - uniform); v = (1-collision)*v + collision*sigma_v*gaussian`` per DOF
    (hack_integrator.py:66-68; OpenMM ``step(x)`` is 0 for x < 0 and 1 otherwise)."""
    coll = (p_collision - uniform >= 0).astype(np.float64)
    sigma_v = np.sqrt(kT / m)[:, None]
    return (1 - coll) * v + coll * sigma_v * gaussian


def bath_energies(st):
    """computeEnergies (hack_integrator.py:483-493): (bathKE, bathPE)."""
    ke = float(np.sum(0.5 * st.Q * st.vxi ** 2))
    pe = st.kT * (st.ndf * st.xi[0] + float(np.sum(st.xi[1:]))) if st.M else 0.0
    return ke, pe


def run_nvt_nhc(force_fn, x, v, m, dt, n_steps, st, n_c=5, n_ys=5, constrain=None):
    """NVT driver loop (test_nosehoover.py:100-118): first half = propagateNHC, kick, drift [, constrain];
    forces; second half = kick [, constrain v], propagateNHC.  ``force_fn(x_nm) -> F``; ``constrain`` is an optional
    pair (positions(x0, x1) -> x1c, velocities(x, v) -> vc).  Returns x, v, f, KE trace."""
    x, v = np.array(x, dtype=np.float64), np.array(v, dtype=np.float64)
    f = force_fn(x)
    ke = []
    for _ in range(n_steps):
        v = nhc_propagate(st, v, m, dt, n_c, n_ys)
        x0 = x
        x, v = vv_first_half(x, v, f, m, dt)
        if constrain is not None:
            xc = constrain[0](x0, x)
            v = v + (xc - x) / dt
            x = xc
        f = force_fn(x)
        v = vv_second_half(v, f, m, dt)
        if constrain is not None:
            v = constrain[1](x, v)
        v = nhc_propagate(st, v, m, dt, n_c, n_ys)
        ke.append(kinetic_energy(v, m))
    return x, v, f, np.array(ke)
