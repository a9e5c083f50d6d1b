"""Integrator hook of the reference (code/hack_integrator.py) without OpenMM.

The reference splits velocity Verlet into two OpenMM ``CustomIntegrator`` programs so that a Python-side
force can be injected between them (``setPerDofVariableByName('force_last' | 'gnn_force', F)`` then
``Simulation.step(1)``).  The classes below keep those names, constructor arguments, per-DOF variable names
and the ``copy_state_from_integrator`` protocol, and run the same per-DOF programs on a device-resident
fp64 state through the C ABI (``gamd_vv_first_half`` / ``gamd_vv_second_half``); a minimal
``System`` / ``Simulation`` / ``CompoundIntegrator`` / ``State`` runtime replaces the OpenMM objects the
driver scripts touch (code/LJ/test_script/test_nosehoover.py:41-57, 100-118).

Units are OpenMM's: nm, ps, Da, K, kJ/mol; quantities are plain floats (objects with ``value_in_unit_system``
are unwrapped when simtk/openmm is importable).  ``System(..., rigid_water=True)`` holds [O,H,H] triplets rigid
with the analytic SETTLE kernels (``gamd_settle_positions`` / ``gamd_settle_velocities``), the stand-in for OpenMM's
ConstrainPositions / ConstrainVelocities on ``WaterBox(constrained=True)`` systems.

  HackNoseHooverIntegrator      :182-330   propagateNHC; v+=0.5*dt*force_last/m; x+=dt*v
  HackHalfNoseHooverIntegrator  :334-493   v+=0.5*dt*gnn_force/m; propagateNHC; bath energies
  HackLangevinIntegrator        :90-169    B, A/2, O, A/2
  HackHalfVelocityIntegrator    :171-178   v+=(dt/2)*gnn_force/m
  HackAndersenVVIntegrator      :17-86     per-particle Andersen collisions + VV with test1/test2 forces
  water-only (code/water/hack_integrator.py): HackIntegratorNHC :18-177, HackDummyIntegrator :180-188,
  HackDummyIntegratorNHC :193-347
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

    def __init__(self, masses, n_constraints=0, has_cm_motion_remover=False, rigid_water=False):
        self.masses = np.asarray(masses, dtype=np.float64)
        self.rigid_water = bool(rigid_water)
        if self.rigid_water and n_constraints == 0:
            n_constraints = len(self.masses)          # three distance constraints per [O,H,H] molecule
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
    def __init__(self, x=None, v=None, ke=None, time_ps=None):
        self._x, self._v, self._ke, self._t = x, v, ke, time_ps

    def getTime(self):
        return self._t

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
        # constrained=True water (code/water/test_script/test_nosehoover.py:33-37): [O,H,H] triplets held rigid by SETTLE
        self.rigid_water = bool(getattr(system, "rigid_water", False))
        self.x_ref = torch.zeros_like(self.x) if self.rigid_water else None

    # ---- the two kick programs every Hack* class is built from ----
    def first_half_kick_drift(self, force, dt):
        """v += 0.5 dt f/m; x += dt v; x1 = x; constrain positions; v += (x - x1)/dt  (hack_integrator.py:273-277)."""
        if self.rigid_water:
            self.x_ref.copy_(self.x)
        self.lib.vv_first_half(self.x, self.v, force, self.mass, dt)
        if self.rigid_water:
            self.lib.settle_positions(self.x_ref, self.x, self.mass, v=self.v, dt_corr=dt)

    def second_half_kick(self, force, dt):
        """v += 0.5 dt f/m; constrain velocities  (hack_integrator.py:421-422, :175-178)."""
        self.lib.vv_second_half(self.v, force, self.mass, dt)
        if self.rigid_water:
            self.lib.settle_velocities(self.x, self.v, self.mass)

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
        return State(x, v, ke, getattr(self, "time_ps", 0.0))


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
    """Nose-Hoover chain of one integrator object: a device-resident ``gamd_nhc_state`` (xi, vxi, G, Q, scale, KE2,
    bath energies) propagated by one CUDA thread from a device-side kinetic-energy reduction
    (``gamd_nhc_propagate`` = propagateNHC, hack_integrator.py:289-316) - no host synchronisation per half step.
    ``getGlobalVariableByName`` reads the state back on demand; ``copy_state_from_integrator`` is a device-to-device
    copy of the chain variables (hack_integrator.py:322-330, :440-452)."""
    YSWeights = YS_WEIGHTS
    _CHAIN_GLOBALS = ("xi", "vxi", "G", "Q")

    def _init_nhc(self, system, collision_frequency, chain_length, num_mts, num_yoshidasuzuki):
        self.n_c, self.n_ys = num_mts, num_yoshidasuzuki
        if self.n_ys not in YS_WEIGHTS:
            raise Exception("Invalid Yoshida-Suzuki value. Allowed values are: %s" % ",".join(map(str, YS_WEIGHTS)))
        if chain_length < 0:
            raise Exception("Nose-Hoover chain length must be at least 0")
        if chain_length > _capi.NHC_MAX:
            raise Exception("Nose-Hoover chain length above %d is not built" % _capi.NHC_MAX)
        self.weights = YS_WEIGHTS[self.n_ys]
        self.M = chain_length
        self._frequency = _val(collision_frequency)
        self._ndf = system.ndf() if system is not None else None
        self._globals.update(ndf=self._ndf, Q=self.kT / self._frequency ** 2)
        self._state = None            # device buffer, created at bind time (needs the context's library handle)
        self._pending = {}            # globals set before bind

    def bind(self, context):
        super().bind(context)
        if self._ndf is None:         # "system was not passed": ndf = number of DOFs (hack_integrator.py:225-231)
            self._ndf = 3 * context.system.getNumParticles()
            self._globals["ndf"] = self._ndf
        self._state = context.lib.nhc_new_state(self.M, self.n_c, self.n_ys, self.kT, self._frequency, self._ndf)
        for k, v in self._pending.items():
            self.setGlobalVariableByName(k, v)
        self._pending = {}

    def _host_state(self):
        return self.context.lib.nhc_get_state(self._state)

    def getGlobalVariableByName(self, name):
        for base in self._CHAIN_GLOBALS:
            if name.startswith(base) and name[len(base):].isdigit() and self._state is not None:
                return float(getattr(self._host_state(), base)[int(name[len(base):])])
        if name in ("bathKE", "bathPE", "scale") and self._state is not None:
            return float(getattr(self._host_state(), name))
        if name == "KE2" and self._state is not None:
            return float(self._host_state().ke2_in)
        return self._globals[name]

    def setGlobalVariableByName(self, name, value):
        for base in self._CHAIN_GLOBALS + ("bathKE", "bathPE"):
            idx = name[len(base):]
            if name.startswith(base) and (idx.isdigit() or (idx == "" and base.startswith("bath"))):
                if self._state is None:
                    self._pending[name] = float(value)
                    return
                h = self._host_state()
                if idx:
                    getattr(h, base)[int(idx)] = float(value)
                else:
                    setattr(h, base, float(value))
                self.context.lib.nhc_set_state(h, self._state)
                return
        self._globals[name] = float(value)

    def propagateNHC(self, bath=False):
        """hack_integrator.py:289-316 / :454-481 on the device (KE2 reduction, chain, v *= scale)."""
        if self.M:
            c = self.context
            c.lib.nhc_propagate(c.v, c.mass, self.dt, state=self._state, bath=bath)

    def copy_state_from_integrator(self, integrator):
        """chain variables xi, vxi, G, Q (and the bath energies for the second-half class) from the other
        integrator: a device-to-device copy of the state record, no host round trip."""
        if self._state is None or integrator._state is None:
            raise RuntimeError("copy_state_from_integrator needs both integrators bound to a Simulation")
        keep = self._host_state() if not isinstance(self, HackHalfNoseHooverIntegrator) else None
        self._state.copy_(integrator._state)
        if keep is not None:          # the first-half class does not copy bathKE / bathPE (hack_integrator.py:322-330)
            h = self._host_state()
            h.bathKE, h.bathPE = keep.bathKE, keep.bathPE
            self.context.lib.nhc_set_state(h, self._state)


class HackNoseHooverIntegrator(_NHCMixin, _HackIntegrator):
    """first half: propagateNHC; v += 0.5*dt*force_last/m; x += dt*v; constrain positions; v += (x-x1)/dt
    (hack_integrator.py:267-277)."""
    PER_DOF = ("force_last", "x1")

    def __init__(self, system=None, temperature=298.0, collision_frequency=50.0, timestep=0.001, chain_length=5,
                 num_mts=5, num_yoshidasuzuki=5):
        _HackIntegrator.__init__(self, temperature, timestep)
        self._init_nhc(system, collision_frequency, chain_length, num_mts, num_yoshidasuzuki)

    def _step(self):
        c = self.context
        self.propagateNHC()
        c.first_half_kick_drift(self._perdof["force_last"], self.dt)


class HackHalfNoseHooverIntegrator(_NHCMixin, _HackIntegrator):
    """second half: v += 0.5*dt*gnn_force/m; constrain velocities; propagateNHC; bath energies
    (hack_integrator.py:419-425)."""
    PER_DOF = ("gnn_force", "x1")

    def __init__(self, system=None, temperature=298.0, collision_frequency=50.0, timestep=0.001, chain_length=5,
                 num_mts=5, num_yoshidasuzuki=5):
        _HackIntegrator.__init__(self, temperature, timestep)
        self._init_nhc(system, collision_frequency, chain_length, num_mts, num_yoshidasuzuki)

    def _step(self):
        c = self.context
        c.second_half_kick(self._perdof["gnn_force"], self.dt)
        self.propagateNHC(bath=True)


class HackIntegratorNHC(_NHCMixin, _HackIntegrator):
    """water-only class (code/water/hack_integrator.py:18-177): the whole thermostatted velocity-Verlet step with both
    forces given: propagateNHC; v += 0.5 dt test1/m; x += dt v; constrain; v += 0.5 dt test2/m + (x-x1)/dt; constrain v;
    propagateNHC; bath energies."""
    PER_DOF = ("test1", "test2", "x1")

    def __init__(self, system=None, temperature=298.0, collision_frequency=50.0, timestep=0.001, chain_length=5,
                 num_mts=5, num_yoshidasuzuki=5):
        _HackIntegrator.__init__(self, temperature, timestep)
        self._init_nhc(system, collision_frequency, chain_length, num_mts, num_yoshidasuzuki)

    def _step(self):
        c = self.context
        self.propagateNHC()
        c.first_half_kick_drift(self._perdof["test1"], self.dt)
        c.second_half_kick(self._perdof["test2"], self.dt)
        self.propagateNHC(bath=True)


class HackDummyIntegrator(_HackIntegrator):
    """water-only class (code/water/hack_integrator.py:180-188): zero time step, constraints only."""

    def __init__(self):
        super().__init__(0.0, 0.0)

    def _step(self):
        c = self.context
        if c.rigid_water:
            c.lib.settle_positions(c.x_ref, c.x, c.mass)
            c.x_ref.copy_(c.x)
            c.lib.settle_velocities(c.x, c.v, c.mass)


class HackDummyIntegratorNHC(_NHCMixin, _HackIntegrator):
    """water-only class (code/water/hack_integrator.py:193-347): the chain propagation alone (thermostat scale
    without moving the atoms)."""

    def __init__(self, system=None, temperature=298.0, collision_frequency=50.0, timestep=0.001, chain_length=5,
                 num_mts=5, num_yoshidasuzuki=5):
        _HackIntegrator.__init__(self, temperature, timestep)
        self._init_nhc(system, collision_frequency, chain_length, num_mts, num_yoshidasuzuki)

    def _step(self):
        self.propagateNHC()


class HackHalfVelocityIntegrator(_HackIntegrator):
    """v += (dt/2)*gnn_force/m; constrain velocities (hack_integrator.py:171-178)."""
    PER_DOF = ("gnn_force",)

    def __init__(self, timestep):
        super().__init__(0.0, timestep)

    def _step(self):
        self.context.second_half_kick(self._perdof["gnn_force"], self.dt)


class HackLangevinIntegrator(_HackIntegrator):
    """B, A/2, O, A/2 with the injected force (hack_integrator.py:141-165).  One fused kernel without constraints
    (``gamd_langevin_first_half``, Philox4x32-10 noise); with rigid water every sub-step is followed by its
    constraint stage exactly as the per-DOF program lists them."""
    PER_DOF = ("force_last", "x1", "sigma")

    def __init__(self, temperature=298.0, collision_rate=1.0, timestep=0.001, constraint_tolerance=1e-8):
        super().__init__(temperature, timestep)
        self._gamma = _val(collision_rate)
        self._globals["a"] = float(np.exp(-self._gamma * self.dt))
        self._globals["b"] = float(np.sqrt(1 - np.exp(-2 * self._gamma * self.dt)))
        self.injected_gaussian = None        # tests: [N,3] standard normals instead of the Philox stream

    def _step(self):
        c, g = self.context, self._globals
        f = self._perdof["force_last"]
        if not c.rigid_water:
            c.lib.langevin_first_half(c.x, c.v, f, c.mass, self.dt, self.kT, self._gamma, gaussian=self.injected_gaussian)
            return
        h = self.dt / 2
        noise = self.injected_gaussian
        if noise is None:
            noise = torch.randn(c.x.shape, dtype=torch.float64, device=c.dev, generator=c.gen)
        c.lib.vv_second_half(c.v, f, c.mass, self.dt)                       # B: v += (dt/2) f/m
        c.lib.settle_velocities(c.x, c.v, c.mass)
        for stage in range(2):                                              # A, O, A
            c.x_ref.copy_(c.x)
            c.x.add_(c.v, alpha=h)
            c.lib.settle_positions(c.x_ref, c.x, c.mass, v=c.v, dt_corr=h)
            c.lib.settle_velocities(c.x, c.v, c.mass)
            if stage == 0:
                c.v.mul_(g["a"]).add_(torch.sqrt(self.kT / c.mass)[:, None] * noise, alpha=g["b"])
                c.lib.settle_velocities(c.x, c.v, c.mass)


class HackAndersenVVIntegrator(_HackIntegrator):
    """per-DOF Andersen collisions (``gamd_andersen_collide``), then v+=0.5*dt*test1/m; x+=dt*v; constrain;
    v+=0.5*dt*test2/m+(x-x1)/dt; constrain v (hack_integrator.py:66-86)."""
    PER_DOF = ("test1", "test2", "x1", "sigma_v", "collision")

    def __init__(self, temperature=298.0, collision_rate=91.0, timestep=0.001):
        super().__init__(temperature, timestep)
        self._globals["p_collision"] = self.dt * _val(collision_rate)
        self.injected_uniform = self.injected_gaussian = None

    def _step(self):
        c = self.context
        c.lib.andersen_collide(c.v, c.mass, self.kT, self._globals["p_collision"], uniform=self.injected_uniform,
                               gaussian=self.injected_gaussian)
        c.first_half_kick_drift(self._perdof["test1"], self.dt)
        c.second_half_kick(self._perdof["test2"], self.dt)


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
    """``Simulation(topology, system, integrator)``: ``context``, ``step``, ``currentStep`` and ``reporters``.

    Reporters follow OpenMM's protocol (``describeNextReport(simulation)`` -> steps until the next report and what the
    State must hold; ``report(simulation, state)``), e.g. ``gamd_b200.openmm_adapter.StateDataReporter``; a plain
    callable is called with the simulation after every ``step`` call.  As in OpenMM, the clock advances by the current
    integrator's step size per step - the reference's two half-step programs therefore make ``Step`` and ``Time`` run
    twice as fast as the MD step count (code/LJ/test_script/test_langevin.py:79-82)."""

    def __init__(self, topology, system, integrator, platform=None, device="cuda:0"):
        self.topology, self.system, self.integrator = topology, system, integrator
        self.context = Context(system, device)
        self.context.time_ps = 0.0
        integrator.bind(self.context)
        self.reporters = []
        self.currentStep = 0

    @property
    def time_ps(self):
        return self.context.time_ps

    def _dt(self):
        cur = self.integrator
        if isinstance(cur, CompoundIntegrator):
            cur = cur._ints[cur.getCurrentIntegrator()]
        return cur.getStepSize() if hasattr(cur, "getStepSize") else 0.0

    def step(self, n=1):
        remaining = int(n)
        while remaining > 0:
            nxt, due = remaining, []
            for r in self.reporters:
                if hasattr(r, "describeNextReport"):
                    d = r.describeNextReport(self)
                    if d[0] < nxt:
                        nxt, due = d[0], [(r, d)]
                    elif d[0] == nxt:
                        due.append((r, d))
            self.integrator.step(nxt)
            self.currentStep += nxt
            self.context.time_ps += nxt * self._dt()
            remaining -= nxt
            for r, d in due:
                st = self.context.getState(getPositions=bool(d[1]), getVelocities=bool(d[2]), getEnergy=bool(d[4]))
                r.report(self, st)
        for r in self.reporters:
            if not hasattr(r, "describeNextReport"):
                r(self)

    def minimizeEnergy(self, *a, **k):
        """no-op: there is no classical potential in the GNN-driven loop"""
