"""OpenMM-facing glue (gamd_b200/openmm_adapter.py): the StateDataReporter log format and reporter protocol on CPU with
a stand-in simulation, and the reference driver loop + reporter on the device-resident runtime (gpu)."""
import io

import numpy as np
import pytest

from gamd_b200.openmm_adapter import KB, StateDataReporter, positions_angstrom, run_gnn_md


class _Sys:
    def __init__(self, masses, n_constraints=0, cmm=False):
        self.masses, self._nc, self.has_cm_motion_remover = np.asarray(masses, dtype=float), n_constraints, cmm

    def getNumParticles(self):
        return len(self.masses)

    def getNumConstraints(self):
        return self._nc


class _State:
    def __init__(self, ke, t, x=None):
        self._ke, self._t, self._x = ke, t, x

    def getKineticEnergy(self):
        return self._ke

    def getTime(self):
        return self._t

    def getPositions(self, asNumpy=True):
        return self._x


class _Sim:
    def __init__(self, system):
        self.system, self.currentStep, self.time_ps = system, 0, 0.0


def test_state_data_reporter_format_and_schedule():
    buf = io.StringIO()
    sim = _Sim(_Sys([39.9] * 258))
    rep = StateDataReporter(buf, 100, totalSteps=1000, step=True, time=True, kineticEnergy=True, temperature=True,
                            separator="\t")
    assert rep.describeNextReport(sim) == (100, False, False, False, True)
    sim.currentStep = 130
    assert rep.describeNextReport(sim)[0] == 70
    sim.currentStep, sim.time_ps = 200, 0.4
    rep.report(sim, _State(321.5, 0.4))
    sim.currentStep = 300
    rep.report(sim, _State(330.25, 0.6))
    lines = buf.getvalue().splitlines()
    assert lines[0] == '#"Step"\t"Time (ps)"\t"Kinetic Energy (kJ/mole)"\t"Temperature (K)"'
    f = lines[1].split("\t")
    assert f[0] == "200" and float(f[1]) == 0.4 and float(f[2]) == 321.5
    assert abs(float(f[3]) - 2 * 321.5 / (3 * 258 * KB)) < 1e-9        # ndf = 3 N, no constraints, no CMMotionRemover
    assert lines[2].split("\t")[0] == "300" and len(lines) == 3
    # constraints and a CMMotionRemover lower the degrees of freedom as in OpenMM
    buf2 = io.StringIO()
    sim2 = _Sim(_Sys([15.999, 1.008, 1.008] * 10, n_constraints=30, cmm=True))
    rep2 = StateDataReporter(buf2, 10, temperature=True)
    rep2.report(sim2, _State(12.0, 0.0))
    assert abs(float(buf2.getvalue().splitlines()[1]) - 2 * 12.0 / ((90 - 30 - 3) * KB)) < 1e-9
    with pytest.raises(ValueError):
        StateDataReporter(buf, 10, potentialEnergy=True)
    with pytest.raises(ValueError):
        StateDataReporter(buf, 10, progress=True)


def test_positions_angstrom_from_plain_nm():
    st = _State(0.0, 0.0, x=np.array([[0.1, 0.2, 0.3]]))
    assert np.allclose(positions_angstrom(st), [[1.0, 2.0, 3.0]])


@pytest.mark.gpu
def test_reference_driver_loop_with_reporter(tmp_path, fixtures_dir):
    """code/LJ/test_script/test_nosehoover.py:41-118 through run_gnn_md on the device-resident runtime: the log has one
    row per 100 OpenMM steps (= 50 MD steps), Step / Time run at twice the MD step count, and the logged kinetic energy
    is the one of the fused device loop (gamd_md_run, Nose-Hoover chain) at the same steps."""
    import os
    import torch
    from types import SimpleNamespace
    from gamd_b200 import _capi
    from gamd_b200.engine import MDEngine, maxwell_boltzmann
    from gamd_b200.hack_integrator import (CompoundIntegrator, HackHalfNoseHooverIntegrator, HackNoseHooverIntegrator,
                                           Simulation, System)
    from gamd_b200.train_network_lj import ParticleNetLightning
    from gamd_b200.weights import random_state_dict
    pos = np.load(os.path.join(fixtures_dir, "lj_init_pos.npy")).astype(np.float64)
    m = np.full(258, 39.9)
    args = SimpleNamespace(use_layer_norm=True, encoding_size=128, hidden_dim=128, edge_embedding_dim=128, drop_edge=False,
                           conv_layer=4, rotate_aug=False, update_edge=False, use_part=False, data_dir="", loss="mae")
    model = ParticleNetLightning(args, precision=_capi.PREC_FP32)
    sd = random_state_dict(1, 5.2, 1.5, kind="lj")
    model.load_state_dict({"pnet_model." + k: v for k, v in sd.items()})
    model.load_training_stats(os.path.join(fixtures_dir, "scaler_lj.npz"))
    model.cuda().eval()
    system = System(m)
    comp = CompoundIntegrator()
    i1 = HackNoseHooverIntegrator(system, 100.0, collision_frequency=25.0, chain_length=10, timestep=0.002)
    i2 = HackHalfNoseHooverIntegrator(system, 100.0, collision_frequency=25.0, chain_length=10, timestep=0.002)
    comp.addIntegrator(i1)
    comp.addIntegrator(i2)
    sim = Simulation(None, system, comp)
    sim.context.setPositions(pos / 10.0)
    sim.context.setPeriodicBoxSize(2.727)
    v0 = maxwell_boltzmann(m, 100.0, 1234)
    sim.context.setVelocities(v0)
    log = tmp_path / "log_nvt_gnn_nosehoover.txt"
    rep = StateDataReporter(str(log), 100, totalSteps=400, step=True, time=True, kineticEnergy=True, temperature=True,
                            separator="\t")
    sim.reporters.append(rep)
    run_gnn_md(sim, comp, i1, i2, model.predict_forces, 100)
    rep.close()
    rows = [l.split("\t") for l in open(log).read().splitlines()]
    assert rows[0][0] == '#"Step"' and len(rows) == 3
    assert [int(r[0]) for r in rows[1:]] == [100, 200]
    assert abs(float(rows[1][1]) - 0.2) < 1e-12 and abs(float(rows[2][1]) - 0.4) < 1e-12     # 2 fs per OpenMM step
    # the fused device loop with the same thermostat reaches the same kinetic energy after 50 and 100 MD steps
    s = np.load(os.path.join(fixtures_dir, "scaler_lj.npz"))
    eng = MDEngine("lj", sd, 27.27, 7.5, m, s["mean"], s["var"], precision=_capi.PREC_FP32)
    eng.ctx.md_configure(thermostat=_capi.THERMO_NHC, kT=KB * 100.0, chain_length=10, num_mts=5, num_ys=5,
                         frequency=25.0, ndf=3 * 258)
    eng.set_state(pos / 10.0, v0)
    ke = torch.zeros(100, dtype=torch.float64, device="cuda:0")
    eng.step(100, 0.002, ke=ke)
    ke = ke.cpu().numpy()
    assert abs(float(rows[1][2]) - ke[49]) <= 1e-6 * ke[49]
    assert abs(float(rows[2][2]) - ke[99]) <= 1e-6 * ke[99]
    assert abs(float(rows[2][3]) - 2 * ke[99] / (3 * 258 * KB)) <= 1e-6 * float(rows[2][3])
    eng.close()


def test_run_gnn_md_call_sequence_matches_the_reference_loop():
    """the order of OpenMM calls of code/LJ/test_script/test_nosehoover.py:100-118, on recording stand-ins: integrator 0 is
    current for the first half (thermostat state copied from the second integrator except on the first step, force in
    ``force_last``), positions are read with ``enforcePeriodicBox`` in Angstrom for the model, integrator 1 for the
    second half (state copied from the first, force in ``gnn_force``); the Langevin pair copies nothing."""
    log = []

    class Integ:
        def __init__(self, name, nhc=True):
            self.name = name
            if nhc:
                self.copy_state_from_integrator = lambda other: log.append((name, "copy_from", other.name))

        def setPerDofVariableByName(self, var, f):
            log.append((self.name, "set", var, float(np.asarray(f).sum())))

    class Compound:
        def setCurrentIntegrator(self, k):
            log.append(("compound", "current", k))

    class Ctx:
        def __init__(self):
            self.n = 0

        def getState(self, getPositions=False, enforcePeriodicBox=False, **kw):
            log.append(("context", "getState", bool(getPositions), bool(enforcePeriodicBox)))
            self.n += 1
            return _State(0.0, 0.0, x=np.full((2, 3), 0.1 * self.n))      # nm

    class Sim:
        def __init__(self):
            self.context, self.currentStep = Ctx(), 0

        def step(self, n):
            self.currentStep += n
            log.append(("simulation", "step", n))

    def forces(pos_angstrom):
        log.append(("model", "predict_forces", float(pos_angstrom[0, 0])))
        return np.full((2, 3), pos_angstrom[0, 0])

    i1, i2 = Integ("first"), Integ("second")
    sim = Sim()
    f = run_gnn_md(sim, Compound(), i1, i2, forces, 2)
    want_step = lambda t, x_in, x_out: [                                     # noqa: E731
        ("compound", "current", 0)] + ([("first", "copy_from", "second")] if t else []) + [
        ("first", "set", "force_last", 6.0 * x_in), ("simulation", "step", 1),
        ("context", "getState", True, True), ("model", "predict_forces", x_out),
        ("compound", "current", 1), ("second", "copy_from", "first"), ("second", "set", "gnn_force", 6.0 * x_out),
        ("simulation", "step", 1)]
    expect = [("context", "getState", True, True), ("model", "predict_forces", 1.0)] + want_step(0, 1.0, 2.0) + want_step(1, 2.0, 3.0)
    norm = lambda seq: [tuple(round(v, 9) if isinstance(v, float) else v for v in e) for e in seq]   # noqa: E731
    assert norm(log) == norm(expect)
    assert sim.currentStep == 4 and np.allclose(f, 3.0)
    # Langevin / Andersen pairs (test_langevin.py:95-113): no thermostat state to copy, the first half reads 'gnn_force'
    log.clear()
    run_gnn_md(Sim(), Compound(), Integ("first", nhc=False), Integ("second", nhc=False), forces, 1, first_var="gnn_force",
               force=np.zeros((2, 3)))
    assert not any(e[1] == "copy_from" for e in log) and log[1] == ("first", "set", "gnn_force", 0.0)
