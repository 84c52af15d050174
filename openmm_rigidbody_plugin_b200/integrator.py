"""Host-side mirror of the reference's public API for the integrator path, in Python where the reference's
user-facing layer is Python (python/rigidbodyplugin.i:49-73 wraps exactly these classes and methods):

    RigidBodyIntegrator(stepSize, bodyIndices), setRotationMode, getRotationMode, setComputeRefinedEnergies,
    step, getBodyIndices, getRigidBodySystem, getKineticEnergies, getRefinedKineticEnergies,
    getPotentialEnergyRefinement                       (openmmapi/include/RigidBodyIntegrator.h:49-137)
    RigidBodySystem.getNumDOF / getNumFree / getNumBodies / ... / getKineticEnergy
                                                       (openmmapi/include/RigidBodySystem.h:27-46)

OpenMM itself is not available in this environment, so `System`, `Context` and `State` below are minimal
stand-ins with OpenMM's method names - just enough to drive the integrator the way OpenMM's Context does
(initialize -> stateChanged(Positions) -> stateChanged(Velocities) -> step).  Every time step runs on the
GPU through librbk's C ABI (rbk_execute_host); forces are host callables evaluated between Part 1 and
Part 2, exactly where ReferenceIntegrateRigidBodyStepKernel::execute calls calcForcesAndEnergy
(platforms/reference/src/ReferenceRigidBodyKernels.cpp:97-102).  Units: nm, ps, amu, kJ/mol (plain floats).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import OpenMMException
from .system import DeviceRigidBodySystem


class System:
    """Particles, constraints, virtual-site flags and host force objects (stand-in for OpenMM::System)."""

    def __init__(self):
        self._mass, self._virtual, self._constraints, self._forces = [], [], [], []

    def addParticle(self, mass):
        self._mass.append(float(mass))
        self._virtual.append(False)
        return len(self._mass) - 1

    def getNumParticles(self):
        return len(self._mass)

    def getParticleMass(self, i):
        return self._mass[i]

    def setVirtualSite(self, i, flag=True):
        self._virtual[i] = bool(flag)

    def isVirtualSite(self, i):
        return self._virtual[i]

    def addConstraint(self, a, b, distance):
        self._constraints.append((int(a), int(b), float(distance)))
        return len(self._constraints) - 1

    def getNumConstraints(self):
        return len(self._constraints)

    def getConstraintParameters(self, i):
        return self._constraints[i]

    def removeConstraint(self, i):
        del self._constraints[i]

    def addForce(self, force):
        """force: object with compute(positions[N,3]) -> (forces[N,3], potential_energy)."""
        self._forces.append(force)
        return len(self._forces) - 1

    def getNumForces(self):
        return len(self._forces)

    def getForce(self, i):
        return self._forces[i]


class HarmonicBondForce:
    """E = 1/2 k (r - r0)^2 per bond (the force used by the reference's testSingleBond)."""

    def __init__(self):
        self._bonds = []

    def addBond(self, a, b, length, k):
        self._bonds.append((int(a), int(b), float(length), float(k)))

    def compute(self, R):
        F = np.zeros_like(R)
        E = 0.0
        for a, b, r0, k in self._bonds:
            d = R[b] - R[a]
            r = float(np.sqrt(d @ d))
            E += 0.5 * k * (r - r0) ** 2
            f = k * (r - r0) * d / r
            F[a] += f
            F[b] -= f
        return F, E


class RigidBodySystem:
    """Read-only view of the integrator's body system (openmmapi/include/RigidBodySystem.h:37-46)."""

    def __init__(self, integrator):
        self._i = integrator

    def _c(self, key):
        if self._i._dev is None:
            return 0
        return self._i._dev.counts()[key]

    def getNumDOF(self):
        return self._c("numDOF")

    def getNumFree(self):
        return self._c("numFree")

    def getNumBodies(self):
        return self._c("numBodies")

    def getNumActualAtoms(self):
        return self._c("numActualAtoms")

    def getNumBodyAtoms(self):
        return self._c("numBodyAtoms")

    def getAtomIndex(self, i):
        return int(self._i._dev.atom_index()[i])

    def getTranslationalEnergy(self):
        return self._i._lastKE[0]

    def getRotationalEnergy(self):
        return self._i._lastKE[1]

    def getKineticEnergy(self):
        return self._i._lastKE[0] + self._i._lastKE[1]


class _DistanceConstraints:
    """Host stand-in for OpenMM's ReferenceConstraints (absent here): simultaneous SHAKE / RATTLE sweeps over the
    System's distance constraints until every one is within `tol` (relative), vectorised over constraints."""

    def __init__(self, system):
        cons = [system.getConstraintParameters(i) for i in range(system.getNumConstraints())]
        self.a = np.array([k[0] for k in cons], dtype=np.int64)
        self.b = np.array([k[1] for k in cons], dtype=np.int64)
        self.d2 = np.array([k[2] for k in cons], dtype=np.float64) ** 2
        mass = np.array([system.getParticleMass(i) for i in range(system.getNumParticles())], dtype=np.float64)
        self.invm = np.where(mass == 0.0, 0.0, 1.0 / np.where(mass == 0.0, 1.0, mass))
        # atoms shared between constraints: damp the simultaneous update by the number of constraints per atom
        count = np.bincount(np.concatenate([self.a, self.b]), minlength=len(mass))
        self.scale = 1.0 / np.maximum(count[self.a], count[self.b])

    def apply(self, old, new, tol):
        a, b, wa, wb = self.a, self.b, self.invm[self.a], self.invm[self.b]
        r0 = old[a] - old[b]
        for _ in range(1000):
            r = new[a] - new[b]
            diff = self.d2 - np.einsum("ij,ij->i", r, r)
            if np.all(np.abs(diff) <= 2.0 * tol * self.d2):
                return
            g = self.scale * diff / (2.0 * (wa + wb) * np.einsum("ij,ij->i", r0, r))
            np.add.at(new, a, (g * wa)[:, None] * r0)
            np.add.at(new, b, -(g * wb)[:, None] * r0)
        raise OpenMMException("constraint solver (SHAKE) did not converge")

    def applyToVelocities(self, R, V, tol):
        a, b, wa, wb = self.a, self.b, self.invm[self.a], self.invm[self.b]
        r = R[a] - R[b]
        for _ in range(1000):
            rv = np.einsum("ij,ij->i", r, V[a] - V[b])
            if np.all(np.abs(rv) <= tol * self.d2):
                return
            g = self.scale * rv / ((wa + wb) * self.d2)
            np.add.at(V, a, -(g * wa)[:, None] * r)
            np.add.at(V, b, (g * wb)[:, None] * r)
        raise OpenMMException("constraint solver (RATTLE) did not converge")


class RigidBodyIntegrator:
    """openmmapi/src/RigidBodyIntegrator.cpp, same names, argument meaning and error behaviour."""

    def __init__(self, stepSize, bodyIndices):
        self._stepSize = float(stepSize)
        self._constraintTolerance = 1e-5                       # RigidBodyIntegrator.cpp:20
        self._bodyIndices = [int(b) for b in bodyIndices]
        self._rotationMode = 0
        self._computeRefinedEnergies = False
        self._context = None
        self._dev = None
        self._lastKE = (0.0, 0.0)

    # -- parameters -----------------------------------------------------------------------------
    def getStepSize(self):
        return self._stepSize

    def setStepSize(self, size):
        self._stepSize = float(size)

    def getConstraintTolerance(self):
        return self._constraintTolerance

    def setConstraintTolerance(self, tol):
        self._constraintTolerance = float(tol)

    def setRotationMode(self, mode):
        if mode < 0:
            raise OpenMMException("Rotation mode cannot be negative")                       # :27-28
        if self._context is not None:
            raise OpenMMException("Cannot set rotation mode: integrator already bound to a context")   # :29-30
        self._rotationMode = int(mode)

    def getRotationMode(self):
        return self._rotationMode

    def setComputeRefinedEnergies(self, compute):
        if self._context is not None:
            raise OpenMMException("Cannot set refined energy computation: integrator already bound to a context")
        self._computeRefinedEnergies = bool(compute)

    def getComputeRefinedEnergies(self):
        return self._computeRefinedEnergies

    def getBodyIndices(self):
        return list(self._bodyIndices)

    def getRigidBodySystem(self):
        return RigidBodySystem(self)

    # -- called by Context ------------------------------------------------------------------------
    def _initialize(self, context):
        if self._context is not None and self._context is not context:
            raise OpenMMException("This Integrator is already bound to a context")          # :41-42
        system = context.getSystem()
        if system.getNumParticles() != len(self._bodyIndices):
            raise OpenMMException("Number of body indices differs from that of atoms in Context")   # :46-47
        n = system.getNumParticles()
        masses = [system.getParticleMass(i) for i in range(n)]
        virt = [system.isVirtualSite(i) for i in range(n)]
        cons = [system.getConstraintParameters(i)[:2] for i in range(system.getNumConstraints())]
        self._dev = DeviceRigidBodySystem(self._bodyIndices, masses, self._rotationMode,
                                          isVirtual=virt if any(virt) else None, constraints=cons or None)
        self._context = context

    def _stateChanged(self, positions_changed):
        """RigidBodyIntegrator::stateChanged (:63-74)."""
        c = self._context
        if positions_changed:
            c._computeForces()
            self._dev.update(c._R, c._V, c._F, True, True)
        else:
            self._dev.update(V=c._V, geometry=False, velocities=True)
        self._dev.upload()
        if self._computeRefinedEnergies:
            self._dev.set_refined_energies(1)

    # -- stepping ---------------------------------------------------------------------------------
    def step(self, steps):
        if self._context is None:
            raise OpenMMException("This Integrator is not bound to a context!")             # :97-98
        c = self._context
        cb = hp = hv = None
        n = c._R.shape[0]
        if c.getSystem().getNumForces() > 0:
            def cb(Rp, Fp, count, user):          # host force evaluation between Part 1 and Part 2
                c._computeForces()
        if c.getSystem().getNumConstraints() > 0:
            # constraints among free atoms: the two calls the reference kernel makes around the force evaluation
            # (ReferenceRigidBodyKernels.cpp:98-104), here through rbk_execute_host_hooks
            solver = _DistanceConstraints(c.getSystem())
            tol = self._constraintTolerance

            def hp(oldp, newp, count, user):
                old = np.ctypeslib.as_array(C.cast(oldp, C.POINTER(C.c_double)), shape=(n, 3))
                solver.apply(old, c._R, tol)
                return 1

            def hv(Rp, Vp, count, user):
                solver.applyToVelocities(c._R, c._V, tol)
                return 1
        self._dev.execute_host(self._stepSize, int(steps), c._R, c._V, c._F, forces=cb, constrain_positions=hp,
                               constrain_velocities=hv)
        c._time += self._stepSize * int(steps)
        c._stepCount += int(steps)

    def getKineticEnergies(self):
        ke = self._dev.kinetic_host(self._context._V)
        self._lastKE = (float(ke[0]), float(ke[1]))
        return [self._lastKE[0], self._lastKE[1]]

    def getRefinedKineticEnergies(self):
        # CudaRigidBodyKernels.cpp:469-476: the refined estimate when it was switched on, else the plain energies
        # (the reference's Reference platform always returns the plain ones, ReferenceRigidBodyKernels.cpp:123-128)
        if not self._computeRefinedEnergies or self._context._stepCount == 0:
            return self.getKineticEnergies()
        ke = self._dev.refined_kinetic_host(self._stepSize, self._context._V)
        return [float(ke[0]), float(ke[1])]

    def getPotentialEnergyRefinement(self):
        # CudaRigidBodyKernels.cpp:481-494: -(dt^2/24) sum of squared forces over masses / torques over inertia
        if not self._computeRefinedEnergies or self._context._stepCount == 0:
            return 0.0
        return self._dev.potential_refinement_host(self._stepSize, self._context._F)

    def _computeKineticEnergy(self):
        return sum(self.getKineticEnergies())


class State:
    def __init__(self, time, R=None, V=None, F=None, ke=None, pe=None):
        self._time, self._R, self._V, self._F, self._ke, self._pe = time, R, V, F, ke, pe

    def getTime(self):
        return self._time

    def getPositions(self):
        return self._R

    def getVelocities(self):
        return self._V

    def getForces(self):
        return self._F

    def getKineticEnergy(self):
        return self._ke

    def getPotentialEnergy(self):
        return self._pe


class Context:
    """Minimal stand-in for OpenMM::Context: owns host R/V/F (float64 [N,3], the Reference platform's data)
    and forwards state changes to the integrator the way ContextImpl does."""

    def __init__(self, system, integrator):
        self._system = system
        n = system.getNumParticles()
        self._R = np.zeros((n, 3))
        self._V = np.zeros((n, 3))
        self._F = np.zeros((n, 3))
        self._pe = 0.0
        self._time = 0.0
        self._stepCount = 0
        self._integrator = integrator
        integrator._initialize(self)

    def getSystem(self):
        return self._system

    def getIntegrator(self):
        return self._integrator

    def _computeForces(self):
        self._F[:] = 0.0
        self._pe = 0.0
        for k in range(self._system.getNumForces()):
            f, e = self._system.getForce(k).compute(self._R)
            self._F += f
            self._pe += e

    def setPositions(self, positions):
        self._R[:] = np.asarray(positions, dtype=np.float64)
        self._integrator._stateChanged(True)

    def setVelocities(self, velocities):
        self._V[:] = np.asarray(velocities, dtype=np.float64)
        self._integrator._stateChanged(False)

    def getState(self, getPositions=False, getVelocities=False, getForces=False, getEnergy=False):
        ke = self._integrator._computeKineticEnergy() if getEnergy else None
        return State(self._time, self._R.copy() if getPositions else None, self._V.copy() if getVelocities else None,
                     self._F.copy() if getForces else None, ke, self._pe if getEnergy else None)

    def getTime(self):
        return self._time
