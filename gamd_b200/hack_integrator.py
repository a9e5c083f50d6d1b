"""Integrator hook of the reference (code/hack_integrator.py) without OpenMM.

The reference splits velocity Verlet into two OpenMM ``CustomIntegrator`` programs so that a Python-side
force can be injected between them (``setPerDofVariableByName('force_last' | 'gnn_force', F)`` then
``Simulation.step(1)``).  The classes below keep those names, constructor arguments, per-DOF variable names
and the ``copy_state_from_integrator`` protocol, and run the same per-DOF programs on a device-resident
fp64 state through the C ABI (``gamd_vv_first_half`` / ``gamd_vv_second_half``); a minimal
``System`` / ``Simulation`` / ``CompoundIntegrator`` / ``State`` runtime replaces the OpenMM objects the
driver scripts touch (code/LJ/test_script/test_nosehoover.py:41-57, 100-118).

Units are OpenMM's: nm, ps, Da, K, kJ/mol; quantities are plain floats (objects with ``value_in_unit_system``
are unwrapped when simtk/openmm is importable).  Constraints (SETTLE) are not built - SURVEY.md 8f rank 2.

  HackNoseHooverIntegrator      :182-330   propagateNHC; v+=0.5*dt*force_last/m; x+=dt*v
  HackHalfNoseHooverIntegrator  :334-493   v+=0.5*dt*gnn_force/m; propagateNHC; bath energies
  HackLangevinIntegrator        :90-169    B, A/2, O, A/2
  HackHalfVelocityIntegrator    :171-178   v+=(dt/2)*gnn_force/m
  HackAndersenVVIntegrator      :17-86     per-particle Andersen collisions + VV with test1/test2 forces
"""
import numpy as np
import torch

from . import _capi
from .md_module import neighbor_context

KB = 0.00831446261815324  # kJ/mol/K
YS_WEIGHTS = {
    1: [1.0],
    3: [0.8289815435887510, -0.6579630871775020, 0.8289815435887510],
    5: [0.2967324292201065, 0.2967324292201065, -0.1869297168804260, 0.2967324292201065, 0.2967324292201065],
}


def _val(q):
    """float from a plain number or an OpenMM Quantity (converted to the md unit system)."""
    if hasattr(q, "value_in_unit_system"):
        try:
            from openmm import unit as u
        except Exception:
            from simtk import unit as u
        return q.value_in_unit_system(u.md_unit_system)
    return float(q)


class System:
    """masses in Da; the only OpenMM ``System`` facts the hook needs."""

    def __init__(self, masses, n_constraints=0, has_cm_motion_remover=False):
        self.masses = np.asarray(masses, dtype=np.float64)
        self.n_constraints = n_constraints
        self.has_cm_motion_remover = has_cm_motion_remover

    def getNumParticles(self):
        return len(self.masses)

    def getNumConstraints(self):
        return self.n_constraints

    def ndf(self):
        dof = 3 * int((self.masses > 0).sum()) - self.n_constraints     # hack_integrator.py:233-239
        return dof - 3 if self.has_cm_motion_remover else dof


class State:
    def __init__(self, x=None, v=None, ke=None):
        self._x, self._v, self._ke = x, v, ke

    def getPositions(self, asNumpy=True):
        return self._x

    def getVelocities(self, asNumpy=True):
        return self._v

    def getKineticEnergy(self):
        return self._ke


class Context:
    """Device-resident positions (nm) / velocities (nm/ps), fp64."""

    def __init__(self, system, device="cuda:0"):
        if not torch.cuda.is_available():
            raise _capi.GamdError(_capi.ENOGPU, "the integrator hook needs a CUDA device (there is no CPU fallback)")
        self.system = system
        self.dev = torch.device(device)
        n = system.getNumParticles()
        self.x = torch.zeros((n, 3), dtype=torch.float64, device=self.dev)
        self.v = torch.zeros_like(self.x)
        self.mass = torch.as_tensor(system.masses, device=self.dev)
        self.box = None
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(0)
        self.lib = neighbor_context(self.dev.index or 0)

    def setPositions(self, pos_nm):
        self.x.copy_(torch.as_tensor(np.asarray(pos_nm, dtype=np.float64)))

    def setVelocities(self, v):
        self.v.copy_(torch.as_tensor(np.asarray(v, dtype=np.float64)))

    def setPeriodicBoxSize(self, box_nm):
        self.box = np.broadcast_to(np.asarray(box_nm, dtype=np.float64), (3,)).copy()

    def setVelocitiesToTemperature(self, temperature, seed=None):
        if seed is not None:
            self.gen.manual_seed(int(seed))
        sigma = torch.sqrt(KB * _val(temperature) / self.mass)[:, None]
        self.v.copy_(torch.randn(self.x.shape, dtype=torch.float64, device=self.dev, generator=self.gen) * sigma)

    def kinetic_energy(self):
        return float(0.5 * (self.mass[:, None] * self.v * self.v).sum().item())

    def getState(self, getPositions=False, getVelocities=False, getEnergy=False, enforcePeriodicBox=False, **kw):
        x = v = ke = None
        if getPositions:
            xs = self.x
            if enforcePeriodicBox and self.box is not None:
                b = torch.as_tensor(self.box, device=self.dev)
                xs = xs - torch.floor(xs / b) * b        # per-atom wrap (the LJ fluid has no molecules)
            x = xs.cpu().numpy()
        if getVelocities:
            v = self.v.cpu().numpy()
        if getEnergy:
            ke = self.kinetic_energy()
        return State(x, v, ke)


class _HackIntegrator:
    """Common per-DOF variable / global plumbing of the CustomIntegrator programs."""
    PER_DOF = ()

    def __init__(self, temperature, timestep):
        self.temperature = _val(temperature)
        self.kT = KB * self.temperature
        self.dt = _val(timestep)
        self._perdof = {}
        self._globals = {"kT": self.kT}
        self.context = None

    def getStepSize(self):
        return self.dt

    def bind(self, context):
        self.context = context
        for name in self.PER_DOF:
            self._perdof.setdefault(name, torch.zeros_like(context.x))

    def setPerDofVariableByName(self, name, values):
        if name not in self.PER_DOF:
            raise KeyError(name)
        t = values if isinstance(values, torch.Tensor) else torch.as_tensor(np.asarray(values, dtype=np.float64))
        self._perdof[name].copy_(t)

    def getPerDofVariableByName(self, name):
        return self._perdof[name].cpu().numpy()

    def getGlobalVariableByName(self, name):
        return self._globals[name]

    def setGlobalVariableByName(self, name, value):
        self._globals[name] = float(value)

    def step(self, n=1):
        for _ in range(n):
            self._step()


class _NHCMixin:
    YSWeights = YS_WEIGHTS

    def _init_nhc(self, system, collision_frequency, chain_length, num_mts, num_yoshidasuzuki):
        self.n_c, self.n_ys = num_mts, num_yoshidasuzuki
        if self.n_ys not in YS_WEIGHTS:
            raise Exception("Invalid Yoshida-Suzuki value. Allowed values are: %s" % ",".join(map(str, YS_WEIGHTS)))
        if chain_length < 0:
            raise Exception("Nose-Hoover chain length must be at least 0")
        self.weights = YS_WEIGHTS[self.n_ys]
        self.M = chain_length
        frequency = _val(collision_frequency)
        q = self.kT / frequency ** 2
        ndf = system.ndf() if system is not None else None
        g = self._globals
        g.update(ndf=ndf, bathKE=0.0, bathPE=0.0, KE2=0.0, Q=q, scale=1.0)
        for i in range(self.M):
            g[f"xi{i}"] = 0.0
            g[f"vxi{i}"] = 0.0
            g[f"G{i}"] = -frequency ** 2
            # Q1..Q{M-1} = Q always (the reference sets them inside the step program, hack_integrator.py:283-287);
            # Q0 = ndf * Q is fixed in propagateNHC when the system (ndf) is only known at bind time
            g[f"Q{i}"] = q if i else (ndf * q if ndf is not None else 0.0)

    def propagateNHC(self):
        """hack_integrator.py:289-316 / :454-481: chain scalars on the host, one KE reduction and one
        velocity scaling on the device."""
        M, g, ctx = self.M, self._globals, self.context
        if not M:
            return
        if g["ndf"] is None:
            g["ndf"] = 3 * ctx.system.getNumParticles()
            g["Q0"] = g["ndf"] * g["Q"]
        ke2 = float((ctx.mass[:, None] * ctx.v * ctx.v).sum().item())
        g["KE2"] = ke2
        kT, ndf, dt = self.kT, g["ndf"], self.dt
        xi = [g[f"xi{i}"] for i in range(M)]
        vxi = [g[f"vxi{i}"] for i in range(M)]
        G = [g[f"G{i}"] for i in range(M)]
        Q = [g[f"Q{i}"] for i in range(M)]
        scale = 1.0
        G[0] = (ke2 - ndf * kT) / Q[0]
        for _ in range(self.n_c):
            for w in self.weights:
                wdt = w * dt / self.n_c
                vxi[M - 1] += 0.25 * wdt * G[M - 1]
                for j in range(M - 2, -1, -1):
                    aa = np.exp(-0.125 * wdt * vxi[j + 1])
                    vxi[j] = aa * (aa * vxi[j] + 0.25 * wdt * G[j])
                aa = np.exp(-0.5 * wdt * vxi[0])
                scale *= aa
                for j in range(M):
                    xi[j] += 0.5 * wdt * vxi[j]
                G[0] = (scale * scale * ke2 - ndf * kT) / Q[0]
                for j in range(M - 1):
                    aa = np.exp(-0.125 * wdt * vxi[j + 1])
                    vxi[j] = aa * (aa * vxi[j] + 0.25 * wdt * G[j])
                    G[j + 1] = (Q[j] * vxi[j] * vxi[j] - kT) / Q[j + 1]
                vxi[M - 1] += 0.25 * wdt * G[M - 1]
        for i in range(M):
            g[f"xi{i}"], g[f"vxi{i}"], g[f"G{i}"] = xi[i], vxi[i], G[i]
        g["scale"] = scale
        ctx.v.mul_(scale)

    def copy_state_from_integrator(self, integrator):
        names = ["bathKE", "bathPE"] if isinstance(self, HackHalfNoseHooverIntegrator) else []
        for i in range(self.M):
            names += [f"xi{i}", f"vxi{i}", f"G{i}", f"Q{i}"]
        for nme in names:
            self.setGlobalVariableByName(nme, integrator.getGlobalVariableByName(nme))


class HackNoseHooverIntegrator(_HackIntegrator, _NHCMixin):
    """first half: propagateNHC; v += 0.5*dt*force_last/m; x += dt*v (hack_integrator.py:267-277)."""
    PER_DOF = ("force_last", "x1")

    def __init__(self, system=None, temperature=298.0, collision_frequency=50.0, timestep=0.001, chain_length=5,
                 num_mts=5, num_yoshidasuzuki=5):
        super().__init__(temperature, timestep)
        self._init_nhc(system, collision_frequency, chain_length, num_mts, num_yoshidasuzuki)

    def _step(self):
        c = self.context
        self.propagateNHC()
        c.lib.vv_first_half(c.x, c.v, self._perdof["force_last"], c.mass, self.dt)


class HackHalfNoseHooverIntegrator(_HackIntegrator, _NHCMixin):
    """second half: v += 0.5*dt*gnn_force/m; propagateNHC; bath energies (hack_integrator.py:419-425)."""
    PER_DOF = ("gnn_force", "x1")

    def __init__(self, system=None, temperature=298.0, collision_frequency=50.0, timestep=0.001, chain_length=5,
                 num_mts=5, num_yoshidasuzuki=5):
        super().__init__(temperature, timestep)
        self._init_nhc(system, collision_frequency, chain_length, num_mts, num_yoshidasuzuki)

    def _step(self):
        c, g = self.context, self._globals
        c.lib.vv_second_half(c.v, self._perdof["gnn_force"], c.mass, self.dt)
        self.propagateNHC()
        g["bathKE"] = sum(0.5 * g[f"Q{i}"] * g[f"vxi{i}"] ** 2 for i in range(self.M))
        if self.M:
            g["bathPE"] = self.kT * (g["ndf"] * g["xi0"] + sum(g[f"xi{i}"] for i in range(1, self.M)))


class HackHalfVelocityIntegrator(_HackIntegrator):
    """v += (dt/2)*gnn_force/m (hack_integrator.py:171-178)."""
    PER_DOF = ("gnn_force",)

    def __init__(self, timestep):
        super().__init__(0.0, timestep)

    def _step(self):
        c = self.context
        c.lib.vv_second_half(c.v, self._perdof["gnn_force"], c.mass, self.dt)


class HackLangevinIntegrator(_HackIntegrator):
    """B, A/2, O, A/2 with the injected force (hack_integrator.py:141-165), no constraints."""
    PER_DOF = ("force_last", "x1", "sigma")

    def __init__(self, temperature=298.0, collision_rate=1.0, timestep=0.001, constraint_tolerance=1e-8):
        super().__init__(temperature, timestep)
        self._gamma = _val(collision_rate)
        self._globals["a"] = float(np.exp(-self._gamma * self.dt))
        self._globals["b"] = float(np.sqrt(1 - np.exp(-2 * self._gamma * self.dt)))

    def _step(self):
        c, g = self.context, self._globals
        sigma = torch.sqrt(self.kT / c.mass)[:, None]
        c.v.add_(self._perdof["force_last"] / c.mass[:, None], alpha=self.dt / 2)
        c.x.add_(c.v, alpha=self.dt / 2)
        noise = torch.randn(c.x.shape, dtype=torch.float64, device=c.dev, generator=c.gen)
        c.v.mul_(g["a"]).add_(sigma * noise, alpha=g["b"])
        c.x.add_(c.v, alpha=self.dt / 2)


class HackAndersenVVIntegrator(_HackIntegrator):
    """per-particle Andersen collisions, then v+=0.5*dt*test1/m; x+=dt*v; v+=0.5*dt*test2/m
    (hack_integrator.py:66-86)."""
    PER_DOF = ("test1", "test2", "x1", "sigma_v", "collision")

    def __init__(self, temperature=298.0, collision_rate=91.0, timestep=0.001):
        super().__init__(temperature, timestep)
        self._globals["p_collision"] = self.dt * _val(collision_rate)

    def _step(self):
        c = self.context
        sigma_v = torch.sqrt(self.kT / c.mass)[:, None]
        u = torch.rand(c.x.shape, dtype=torch.float64, device=c.dev, generator=c.gen)
        gauss = torch.randn(c.x.shape, dtype=torch.float64, device=c.dev, generator=c.gen)
        coll = (self._globals["p_collision"] - u >= 0).to(torch.float64)      # step(p_collision - uniform)
        c.v.copy_((1 - coll) * c.v + coll * sigma_v * gauss)
        c.lib.vv_first_half(c.x, c.v, self._perdof["test1"], c.mass, self.dt)
        c.lib.vv_second_half(c.v, self._perdof["test2"], c.mass, self.dt)


class CompoundIntegrator:
    def __init__(self):
        self._ints = []
        self._cur = 0

    def addIntegrator(self, integrator):
        self._ints.append(integrator)
        return len(self._ints) - 1

    def setCurrentIntegrator(self, k):
        self._cur = k

    def getCurrentIntegrator(self):
        return self._cur

    def bind(self, context):
        for i in self._ints:
            i.bind(context)

    def step(self, n=1):
        self._ints[self._cur].step(n)


class Simulation:
    """``Simulation(topology, system, integrator)``: only ``context``, ``step`` and ``reporters``."""

    def __init__(self, topology, system, integrator, platform=None, device="cuda:0"):
        self.system, self.integrator = system, integrator
        self.context = Context(system, device)
        integrator.bind(self.context)
        self.reporters = []
        self.currentStep = 0

    def step(self, n=1):
        self.integrator.step(n)
        self.currentStep += n
        for r in self.reporters:
            r(self)

    def minimizeEnergy(self, *a, **k):
        """no-op: there is no classical potential in the GNN-driven loop"""
