"""Oracle force facade and MD driver (numpy f64 + torch CPU fp32).  TEST INFRASTRUCTURE ONLY.

Restates ``ParticleNetLightning.predict_forces`` (code/LJ/train_network_lj.py:133-157,
code/water/train_network_tip3p.py:142-159) and the two-half-step driver loop
(code/LJ/test_script/test_nosehoover.py:100-118) with the NVE program (chain_length=0).
"""
import numpy as np
import torch

from . import integrator as oint
from . import model as omodel
from . import neighbor as onb


class OracleForceField:
    """positions (Angstrom, any float dtype, [N,3]) -> forces (kJ/mol/nm, float64 [N,3])."""

    def __init__(self, sd, kind, box, cutoff, scaler_mean=0.0, scaler_var=1.0, bond=None, feat=None):
        self.sd = {k: torch.as_tensor(v) for k, v in sd.items()}
        self.kind, self.box, self.cutoff = kind, box, cutoff
        self.mean = np.asarray(scaler_mean, dtype=np.float64).reshape(-1)
        self.var = np.asarray(scaler_var, dtype=np.float64).reshape(-1)
        self.bond, self.feat = bond, feat

    def edges(self, pos):
        # search_for_neighbor: device_put (f64 -> f32), jnp.mod, strict predicate, self kept
        return onb.edges_jaxmd(np.asarray(pos), self.box, self.cutoff)

    def predict_forces(self, pos):
        pos = np.asarray(pos)
        edge = torch.from_numpy(self.edges(pos))
        p = torch.from_numpy(np.mod(pos, np.array(self.box))).float()        # lj:141-142
        pred = omodel.forward(self.sd, self.kind, [p], [edge], self.box, x=self.feat, bond=self.bond)
        pred = pred.numpy()
        return pred * np.sqrt(self.var) + self.mean                          # lj:128-131 (f64)


def maxwell_boltzmann(n, masses, temperature, seed):
    """v ~ N(0, sqrt(kB T / m)) nm/ps from a numpy PCG64 stream (stand-in for OpenMM's
    ``setVelocitiesToTemperature``, whose RNG is not reproducible outside OpenMM)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sigma = np.sqrt(oint.KB * temperature / masses)[:, None]
    return rng.standard_normal((n, 3)) * sigma


def run_nve(ff, x_nm, v, masses, dt, n_steps, report_every=1):
    """The driver loop of test_nosehoover.py:100-118 with the NVE program.

    x_nm float64 [N,3] in nm; forces are evaluated at ``x*10`` Angstrom.  Returns final
    (x, v, f) and the KE trace (total, COM-removed) sampled every ``report_every`` steps."""
    x = np.array(x_nm, dtype=np.float64)
    v = np.array(v, dtype=np.float64)
    m = np.asarray(masses, dtype=np.float64)
    f = ff.predict_forces(x * 10.0)
    trace = []
    for t in range(n_steps):
        x, v = oint.vv_first_half(x, v, f, m, dt)
        f = ff.predict_forces(x * 10.0)
        v = oint.vv_second_half(v, f, m, dt)
        if (t + 1) % report_every == 0:
            vcom = (m[:, None] * v).sum(0) / m.sum()
            trace.append((t + 1, oint.kinetic_energy(v, m), oint.kinetic_energy(v - vcom, m)))
    return x, v, f, np.array(trace)
