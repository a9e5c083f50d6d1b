"""Glue between the GNN force facade and an OpenMM-style simulation: the driver loop of the reference's test scripts
and the ``StateDataReporter`` log they write.

The reference drives OpenMM from Python (code/LJ/test_script/test_nosehoover.py:82-118, test_langevin.py:79-113,
code/water/test_script/test_nosehoover.py:88-128): a ``CompoundIntegrator`` of two half-step programs, the GNN force
injected between them with ``setPerDofVariableByName``, and a ``StateDataReporter`` that logs step, time, kinetic energy
and temperature.  Everything here talks to the simulation only through those OpenMM calls, so it runs

  * on the OpenMM-free runtime of ``gamd_b200.hack_integrator`` (``Simulation`` / ``CompoundIntegrator`` / ``Hack*``
    classes on device-resident state; OpenMM is not installed in this image), and
  * on real OpenMM objects when ``openmm`` / ``simtk.openmm`` is importable (positions arrive as ``Quantity`` and are
    unwrapped; forces are handed over as plain kJ/mol/nm arrays, which ``setPerDofVariableByName`` accepts).

``StateDataReporter`` follows OpenMM's reporter protocol (``describeNextReport`` / ``report``) and file format (quoted
column headers after ``#``, one row per report), restricted to the columns the reference asks for.
"""
import sys
import time as _time

import numpy as np

KB = 0.00831446261815324  # kJ/mol/K


def _unwrap(q, unit_name):
    """plain numpy array from an OpenMM Quantity (converted to ``unit_name``) or from a plain array (nm, ps, kJ/mol)."""
    if hasattr(q, "value_in_unit"):
        try:
            from openmm import unit as u
        except Exception:
            from simtk import unit as u
        return np.asarray(q.value_in_unit(getattr(u, unit_name)))
    return np.asarray(q)


def positions_angstrom(state):
    """``state.getPositions(asNumpy=True).value_in_unit(unit.angstrom)`` (test_nosehoover.py:112)"""
    p = state.getPositions(asNumpy=True)
    if hasattr(p, "value_in_unit"):
        return _unwrap(p, "angstrom")
    return np.asarray(p) * 10.0          # the OpenMM-free runtime returns nm


class StateDataReporter:
    """OpenMM's ``StateDataReporter`` restricted to what the reference logs (test_nosehoover.py:82-86):
    ``StateDataReporter(file, reportInterval, step=True, time=True, kineticEnergy=True, temperature=True,
    totalSteps=..., separator='\\t')``; optional ``progress`` / ``remainingTime`` / ``speed`` / ``elapsedTime`` columns.

    Temperature is ``2 KE / (ndf k_B)`` with ``ndf = 3 N(mass > 0) - constraints - 3 [CMMotionRemover]`` as OpenMM
    computes it.  Note the reference's clock: every MD step is two ``Simulation.step(1)`` calls, so ``Step`` and ``Time``
    run twice as fast as the MD step count (SURVEY.md section 8a note 8) - reproduced, not corrected."""

    def __init__(self, file, reportInterval, step=False, time=False, potentialEnergy=False, kineticEnergy=False,
                 totalEnergy=False, temperature=False, volume=False, density=False, progress=False,
                 remainingTime=False, speed=False, elapsedTime=False, separator=",", systemMass=None, totalSteps=None,
                 append=False):
        if potentialEnergy or totalEnergy:
            raise ValueError("the GNN predicts forces directly: there is no potential energy to report "
                             "(the reference logs kinetic energy and temperature only)")
        if volume or density:
            raise ValueError("volume / density columns are not supported")
        if (progress or remainingTime) and totalSteps is None:
            raise ValueError("Reporting progress or remaining time requires total steps to be specified")
        self._interval = int(reportInterval)
        self._own = isinstance(file, str)
        self._out = open(file, "a" if append else "w") if self._own else file
        self._append = append
        self._cols = dict(step=step, time=time, kineticEnergy=kineticEnergy, temperature=temperature, progress=progress,
                          remainingTime=remainingTime, speed=speed, elapsedTime=elapsedTime)
        self._sep = separator
        self._total = totalSteps
        self._has_init = False
        self._need_energy = kineticEnergy or temperature

    # ---- OpenMM reporter protocol ----
    def describeNextReport(self, simulation):
        steps = self._interval - simulation.currentStep % self._interval
        return (steps, False, False, False, self._need_energy)

    def report(self, simulation, state):
        if not self._has_init:
            self._init(simulation)
        vals = self._values(simulation, state)
        print(self._sep.join(str(v) for v in vals), file=self._out)
        try:
            self._out.flush()
        except AttributeError:
            pass

    # ---- internals ----
    def _init(self, simulation):
        system = simulation.system
        n = system.getNumParticles()
        if hasattr(system, "masses"):
            massive = int((np.asarray(system.masses) > 0).sum())
        else:
            massive = sum(1 for i in range(n) if _unwrap(system.getParticleMass(i), "dalton") > 0)
        dof = 3 * massive - system.getNumConstraints()
        if hasattr(system, "has_cm_motion_remover"):
            cmm = bool(system.has_cm_motion_remover)
        else:
            cmm = any(type(system.getForce(i)).__name__ == "CMMotionRemover" for i in range(system.getNumForces()))
        self._dof = dof - 3 if cmm else dof
        self._t0 = _time.time()
        self._step0 = simulation.currentStep
        self._time0 = self._sim_time(simulation, None)
        if not self._append:
            heads = []
            c = self._cols
            if c["progress"]: heads.append("Progress (%)")
            if c["step"]: heads.append("Step")
            if c["time"]: heads.append("Time (ps)")
            if c["kineticEnergy"]: heads.append("Kinetic Energy (kJ/mole)")
            if c["temperature"]: heads.append("Temperature (K)")
            if c["speed"]: heads.append("Speed (ns/day)")
            if c["elapsedTime"]: heads.append("Elapsed Time (s)")
            if c["remainingTime"]: heads.append("Time Remaining")
            print('#"%s"' % ('"' + self._sep + '"').join(heads), file=self._out)
        self._has_init = True

    @staticmethod
    def _sim_time(simulation, state):
        if state is not None and hasattr(state, "getTime"):
            t = state.getTime()
            if t is not None:
                return float(_unwrap(t, "picosecond"))
        return float(getattr(simulation, "time_ps", 0.0))

    def _values(self, simulation, state):
        c = self._cols
        vals = []
        now = _time.time()
        if c["progress"]:
            vals.append("%.1f%%" % (100.0 * simulation.currentStep / self._total))
        if c["step"]:
            vals.append(simulation.currentStep)
        t_ps = self._sim_time(simulation, state)
        if c["time"]:
            vals.append(t_ps)
        ke = None
        if self._need_energy:
            ke = float(_unwrap(state.getKineticEnergy(), "kilojoules_per_mole"))
        if c["kineticEnergy"]:
            vals.append(ke)
        if c["temperature"]:
            vals.append(2.0 * ke / (self._dof * KB))
        elapsed = now - self._t0
        if c["speed"]:
            ns_day = (t_ps - self._time0) / 1000.0 * 86400.0 / elapsed if elapsed > 0 else 0.0
            vals.append("%.3g" % ns_day if elapsed > 0 else "--")
        if c["elapsedTime"]:
            vals.append(elapsed)
        if c["remainingTime"]:
            done = simulation.currentStep - self._step0
            if done <= 0:
                vals.append("--")
            else:
                rem = int(elapsed * (self._total - simulation.currentStep) / done)
                d, rem = divmod(rem, 86400)
                h, rem = divmod(rem, 3600)
                m, s = divmod(rem, 60)
                vals.append(("%d:" % d if d else "") + ("%d:%02d:%02d" % (h, m, s) if d or h else "%d:%02d" % (m, s)))
        return vals

    def close(self):
        if self._own:
            self._out.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_gnn_md(simulation, compound, first, second, predict_forces, n_steps, first_var="force_last",
               second_var="gnn_force", force=None, copy_thermostat_state=None, progress_every=0, out=sys.stdout):
    """The reference driver loop (code/LJ/test_script/test_nosehoover.py:100-118; test_langevin.py:95-113 with
    ``first_var='gnn_force'``; water :110-128 with ``predict_forces = lambda pos: model.predict_forces(feat, pos)``):

        for t in range(n_steps):
            setCurrentIntegrator(0); [first.copy_state_from_integrator(second)]; first.<first_var> = F; step(1)
            pos = getState(getPositions, enforcePeriodicBox).getPositions() [Angstrom]; F = predict_forces(pos)
            setCurrentIntegrator(1); [second.copy_state_from_integrator(first)]; second.<second_var> = F; step(1)

    ``copy_thermostat_state``: None = do it when both integrators have ``copy_state_from_integrator`` (the Nose-Hoover
    pair).  ``force`` is F(x0) if the caller already has it.  Returns the last force array (kJ/mol/nm)."""
    if copy_thermostat_state is None:
        copy_thermostat_state = hasattr(first, "copy_state_from_integrator") and hasattr(second, "copy_state_from_integrator")
    if force is None:
        st = simulation.context.getState(getPositions=True, enforcePeriodicBox=True)
        force = predict_forces(positions_angstrom(st))
    for t in range(int(n_steps)):
        if progress_every and (t + 1) % progress_every == 0:
            print(f"Finished {t + 1} steps", file=out)
        compound.setCurrentIntegrator(0)
        if copy_thermostat_state and t != 0:
            first.copy_state_from_integrator(second)
        first.setPerDofVariableByName(first_var, force)
        simulation.step(1)
        st = simulation.context.getState(getPositions=True, enforcePeriodicBox=True)
        force = predict_forces(positions_angstrom(st))
        compound.setCurrentIntegrator(1)
        if copy_thermostat_state:
            second.copy_state_from_integrator(first)
        second.setPerDofVariableByName(second_var, force)
        simulation.step(1)
    return force
