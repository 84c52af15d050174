"""StateDataReporter with the rigid-body columns of the reference's Python layer (python/statedatareporter.py:3-88):
translational / rotational kinetic energy, the "refined" energy columns, and a temperature computed with the
rigid-body number of degrees of freedom (RigidBodySystem.getNumDOF(), minus 3 when a CMMotionRemover is present).

With OpenMM's application layer installed this subclasses app.StateDataReporter like the reference.  Without it
(this environment) a small self-contained reporter with the same constructor keywords and column headers is used;
it works with any `simulation` object exposing .integrator, .context (getState(getEnergy=True)), .system and
.currentStep.  Energies are plain kJ/mol floats here (the SWIG layer of the reference attaches units)."""
from __future__ import annotations

import sys

MOLAR_GAS_CONSTANT_R = 0.00831446261815324      # kJ/(mol K)

_EXTRA = [
    ("translationalEnergy", "Translational Energy (kJ/mole)"),
    ("rotationalEnergy", "Rotational Energy (kJ/mole)"),
    ("refinedPotentialEnergy", "Refined Potential Energy (kJ/mole)"),
    ("refinedKineticEnergy", "Refined Kinetic Energy (kJ/mole)"),
    ("refinedTotalEnergy", "Refined Total Energy (kJ/mole)"),
    ("refinedTemperature", "Refined Temperature (K)"),
    ("refinedTranslationalEnergy", "Refined Translational Energy (kJ/mole)"),
    ("refinedRotationalEnergy", "Refined Rotational Energy (kJ/mole)"),
]


def _is_rigid(simulation):
    return hasattr(simulation.integrator, "getRigidBodySystem")


def _value(x):
    return x.value_in_unit(x.unit) if hasattr(x, "value_in_unit") and hasattr(x, "unit") else float(x)


class _RigidBodyColumns:
    """The columns the reference adds, shared by both reporter variants."""

    def _pop_extra(self, kwargs):
        for key, _ in _EXTRA:
            setattr(self, "_" + key, kwargs.pop(key, False))

    def _rigid_dof(self, simulation):
        dof = simulation.integrator.getRigidBodySystem().getNumDOF()
        system = simulation.system
        forces = [system.getForce(i) for i in range(system.getNumForces())]
        if any(type(f).__name__ == "CMMotionRemover" for f in forces):
            dof -= 3
        return dof

    def _extra_headers(self):
        return [header for key, header in _EXTRA if getattr(self, "_" + key)]

    def _extra_values(self, simulation, potential, kinetic):
        values = []
        if self._translationalEnergy or self._rotationalEnergy:
            KE = simulation.integrator.getKineticEnergies() if _is_rigid(simulation) else [kinetic, 0.0]
            if self._translationalEnergy:
                values.append(_value(KE[0]))
            if self._rotationalEnergy:
                values.append(_value(KE[1]))
        U = potential
        if self._refinedPotentialEnergy or self._refinedTotalEnergy:
            if _is_rigid(simulation):
                U = U + _value(simulation.integrator.getPotentialEnergyRefinement())
            if self._refinedPotentialEnergy:
                values.append(U)
        if (self._refinedKineticEnergy or self._refinedTotalEnergy or self._refinedTemperature or
                self._refinedTranslationalEnergy or self._refinedRotationalEnergy):
            KE = simulation.integrator.getRefinedKineticEnergies() if _is_rigid(simulation) else [kinetic, 0.0]
            kt, kr = _value(KE[0]), _value(KE[1])
            if self._refinedKineticEnergy:
                values.append(kt + kr)
            if self._refinedTotalEnergy:
                values.append(kt + kr + U)
            if self._refinedTemperature:
                values.append(2 * (kt + kr) / (self._dof * MOLAR_GAS_CONSTANT_R))
            if self._refinedTranslationalEnergy:
                values.append(kt)
            if self._refinedRotationalEnergy:
                values.append(kr)
        return values


try:
    try:
        from openmm import app as _app, unit as _unit
    except ImportError:
        from simtk.openmm import app as _app
        from simtk import unit as _unit

    class StateDataReporter(_RigidBodyColumns, _app.StateDataReporter):
        def __init__(self, *args, **kwargs):
            self._pop_extra(kwargs)
            super().__init__(*args, **kwargs)

        def _initializeConstants(self, simulation):
            super()._initializeConstants(simulation)
            if (self._temperature or self._refinedTemperature) and _is_rigid(simulation):
                self._dof = self._rigid_dof(simulation)
            elif self._refinedTemperature:
                self._dof = 1

        def _constructHeaders(self):
            return super()._constructHeaders() + self._extra_headers()

        def _constructReportValues(self, simulation, state):
            values = super()._constructReportValues(simulation, state)
            pe = state.getPotentialEnergy().value_in_unit(_unit.kilojoules_per_mole)
            ke = state.getKineticEnergy().value_in_unit(_unit.kilojoules_per_mole)
            return values + self._extra_values(simulation, pe, ke)

except ImportError:

    class StateDataReporter(_RigidBodyColumns):
        """Self-contained reporter (no OpenMM): same keywords and headers as app.StateDataReporter for the columns
        it supports (step, time, potentialEnergy, kineticEnergy, totalEnergy, temperature) plus the rigid-body ones."""

        def __init__(self, file, reportInterval, step=False, time=False, potentialEnergy=False, kineticEnergy=False,
                     totalEnergy=False, temperature=False, separator=",", **kwargs):
            self._pop_extra(kwargs)
            if kwargs:
                raise TypeError("unexpected keyword arguments: %s" % ", ".join(kwargs))
            self._out = open(file, "w") if isinstance(file, str) else (file or sys.stdout)
            self._reportInterval = int(reportInterval)
            self._step, self._time, self._potentialEnergy = step, time, potentialEnergy
            self._kineticEnergy, self._totalEnergy, self._temperature = kineticEnergy, totalEnergy, temperature
            self._separator = separator
            self._hasInitialized = False
            self._dof = 1

        def describeNextReport(self, simulation):
            steps = self._reportInterval - simulation.currentStep % self._reportInterval
            return (steps, False, False, False, True)

        def _initializeConstants(self, simulation):
            if (self._temperature or self._refinedTemperature) and _is_rigid(simulation):
                self._dof = self._rigid_dof(simulation)       # DOF from the rigid-body system, not 3N
            elif self._temperature:
                system = simulation.system
                self._dof = 3 * system.getNumParticles() - system.getNumConstraints()
            elif self._refinedTemperature:
                self._dof = 1

        def _constructHeaders(self):
            headers = []
            if self._step:
                headers.append("Step")
            if self._time:
                headers.append("Time (ps)")
            if self._potentialEnergy:
                headers.append("Potential Energy (kJ/mole)")
            if self._kineticEnergy:
                headers.append("Kinetic Energy (kJ/mole)")
            if self._totalEnergy:
                headers.append("Total Energy (kJ/mole)")
            if self._temperature:
                headers.append("Temperature (K)")
            return headers + self._extra_headers()

        def _constructReportValues(self, simulation, state):
            pe, ke = _value(state.getPotentialEnergy()), _value(state.getKineticEnergy())
            values = []
            if self._step:
                values.append(simulation.currentStep)
            if self._time:
                values.append(_value(state.getTime()))
            if self._potentialEnergy:
                values.append(pe)
            if self._kineticEnergy:
                values.append(ke)
            if self._totalEnergy:
                values.append(pe + ke)
            if self._temperature:
                values.append(2 * ke / (self._dof * MOLAR_GAS_CONSTANT_R))
            return values + self._extra_values(simulation, pe, ke)

        def report(self, simulation, state=None):
            if not self._hasInitialized:
                self._initializeConstants(simulation)
                print('#"%s"' % ('"' + self._separator + '"').join(self._constructHeaders()), file=self._out)
                self._hasInitialized = True
            if state is None:
                state = simulation.context.getState(getEnergy=True)
            print(self._separator.join(str(v) for v in self._constructReportValues(simulation, state)), file=self._out)
            if hasattr(self._out, "flush"):
                self._out.flush()
