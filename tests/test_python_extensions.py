"""CPU tests of the Python layer around the hot path: ForceField body resolution (python/forcefield.py of the
reference), StateDataReporter columns/DOF (python/statedatareporter.py) and the integrator's XML proxy
(serialization/tests/TestSerializeRigidBodyIntegrator.cpp:48-63).  OpenMM is absent, so topologies are fakes."""
import io

import pytest

from openmm_rigidbody_plugin_b200 import RigidBodyIntegrator, serialization
from openmm_rigidbody_plugin_b200.forcefield import ForceField
from openmm_rigidbody_plugin_b200.statedatareporter import MOLAR_GAS_CONSTANT_R, StateDataReporter


class Atom:
    def __init__(self, name, index):
        self.name, self.index = name, index


class Residue:
    def __init__(self, name, atoms):
        self.name, self._atoms = name, atoms

    def atoms(self):
        return iter(self._atoms)


class Topology:
    def __init__(self, residues):
        self._res = []
        k = 0
        for name, atom_names in residues:
            atoms = []
            for a in atom_names:
                atoms.append(Atom(a, k))
                k += 1
            self._res.append(Residue(name, atoms))
        self._n = k

    def residues(self):
        return iter(self._res)

    def getNumAtoms(self):
        return self._n


class TemplateAtom:
    def __init__(self, name):
        self.name = name


class VSite:
    def __init__(self, index):
        self.index = index


class Template:
    def __init__(self, names, virtual=()):
        self.atoms = [TemplateAtom(n) for n in names]
        self.virtualSites = [VSite(names.index(v)) for v in virtual]


class FakeNonbondedForce:
    isNonbonded = True

    def __init__(self):
        self.exceptions = {}

    def addException(self, i, j, chargeProd, sigma, epsilon, replace=False):
        self.exceptions[(min(i, j), max(i, j))] = (chargeProd, sigma, epsilon)


class FakeSystem:
    def __init__(self, n, constraints=(), forces=()):
        self.n, self.constraints, self.forces = n, list(constraints), list(forces)

    def getNumParticles(self):
        return self.n

    def getNumConstraints(self):
        return len(self.constraints)

    def getConstraintParameters(self, i):
        return self.constraints[i]

    def removeConstraint(self, i):
        del self.constraints[i]

    def getNumForces(self):
        return len(self.forces)

    def getForce(self, i):
        return self.forces[i]


def make_ff(builder=None):
    templates = {"HOH": Template(["O", "H1", "H2", "M"], virtual=["M"]), "ALA": Template(["N", "CA", "CB", "C", "O"]),
                 "NA": Template(["NA"])}
    return ForceField(templates, system_builder=builder)


TOPOLOGY = [("HOH", ["O", "H1", "H2", "M"]), ("ALA", ["N", "CA", "CB", "C", "O"]), ("NA", ["NA"]), ("HOH", ["O", "H1", "H2", "M"])]


def test_register_and_resolve_bodies():
    ff = make_ff()
    ff.registerBodyTemplate("water", "HOH")
    ff.registerBodyTemplate("backbone", "ALA", pattern="(N|CA|C)$")
    with pytest.raises(ValueError, match="already been registered"):
        ff.registerBodyTemplate("water", "HOH")
    with pytest.raises(ValueError, match="Unknown residue"):
        ff.registerBodyTemplate("x", "GLY")
    top = Topology(TOPOLOGY)
    # bodies are numbered per (residue, template) in topology order; virtual site M and unmatched atoms stay free
    assert ff.resolveBodies(top) == [1, 1, 1, 0, 2, 2, 0, 2, 0, 0, 3, 3, 3, 0]
    assert str(ff.getBodyTemplate("water")).startswith("HOH")


def test_merge_list_semantics():
    ff = make_ff()
    ff.registerBodyTemplate("water", "HOH")
    ff.registerBodyTemplate("ala", "ALA")
    top = Topology(TOPOLOGY)
    assert ff.resolveBodies(top) == [1, 1, 1, 0, 2, 2, 2, 2, 2, 0, 3, 3, 3, 0]
    # flat list: one group; the fused body takes the smallest index, leaving a gap (index 3 unused)
    assert ff.resolveBodies(top, merge=[3, 2]) == [1, 1, 1, 0, 2, 2, 2, 2, 2, 0, 2, 2, 2, 0]
    # nested, overlapping groups are unioned transitively
    assert ff.resolveBodies(top, merge=[[1, 2], [2, 3]]) == [1, 1, 1, 0, 1, 1, 1, 1, 1, 0, 1, 1, 1, 0]
    assert sorted(map(sorted, ff._disjointSets([{1, 2}, {5, 6}, {2, 3}, {7}, {6, 7}]))) == [[1, 2, 3], [5, 6, 7]]
    with pytest.raises(ValueError, match="not a sequence"):
        ff.resolveBodies(top, merge=3)
    ff2 = make_ff()
    ff2.registerBodyTemplate("odd", "ALA", pattern="XYZ")
    with pytest.raises(ValueError, match="no atom in residue ALA matches pattern XYZ"):
        ff2.resolveBodies(top)


def test_create_system_removes_constraints_and_forces():
    nb = FakeNonbondedForce()

    def builder(topology, **kwargs):
        assert kwargs == {"nonbondedMethod": "PME"}                       # extra keywords are forwarded untouched
        return FakeSystem(14, constraints=[(0, 1, 0.0957), (0, 2, 0.0957), (4, 5, 0.15), (8, 9, 0.2)], forces=[nb, object()])

    ff = make_ff(builder)
    ff.registerBodyTemplate("water", "HOH")
    system, bodies = ff.createSystem(Topology(TOPOLOGY), nonbondedMethod="PME", removeConstraints=True, removeForces=True)
    assert bodies == [1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 0]
    assert system.constraints == [(4, 5, 0.15), (8, 9, 0.2)]               # only constraints touching body atoms go
    assert sorted(nb.exceptions) == [(0, 1), (0, 2), (1, 2), (10, 11), (10, 12), (11, 12)]
    assert all(v == (0, 1, 0) for v in nb.exceptions.values())
    assert sorted(ff._intraBodyPairs([1, 0, 1, 1])) == [(2, 0), (3, 0), (3, 2)]


def test_bodies_from_forcefield_feed_the_integrator():
    """createSystem's bodyIndices (merged labels with gaps) go straight into the integrator's index cleaning."""
    from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem
    ff = make_ff(lambda topology, **kw: FakeSystem(14))
    ff.registerBodyTemplate("water", "HOH")
    ff.registerBodyTemplate("ala", "ALA")
    _, bodies = ff.createSystem(Topology(TOPOLOGY), mergeList=[1, 3])
    assert bodies == [1, 1, 1, 0, 2, 2, 2, 2, 2, 0, 1, 1, 1, 0]
    masses = [16.0, 1.0, 1.0, 0.0, 14.0, 12.0, 12.0, 12.0, 16.0, 23.0, 16.0, 1.0, 1.0, 0.0]
    virt = [0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1]
    s = DeviceRigidBodySystem(bodies, masses, 0, isVirtual=virt)
    c = s.counts()
    assert (c["numBodies"], c["numFree"], c["numActualAtoms"], c["numBodyAtoms"]) == (2, 1, 12, 11)
    assert s.atom_index().tolist() == [9, 0, 1, 2, 10, 11, 12, 4, 5, 6, 7, 8]


def test_serialization_round_trip_like_the_reference_test():
    integ = RigidBodyIntegrator(0.00342, [5, 4, 3, 2, 1])
    integ.setRotationMode(7)
    xml = serialization.serialize(integ)
    assert 'version="1"' in xml and xml.count("<bodyIndex ") == 5
    back = serialization.deserialize(xml)
    assert back.getConstraintTolerance() == integ.getConstraintTolerance()
    assert back.getStepSize() == integ.getStepSize()
    assert back.getBodyIndices() == [5, 4, 3, 2, 1]
    assert back.getRotationMode() == 0                       # not persisted, exactly like the reference proxy
    with pytest.raises(Exception, match="Unsupported version number"):
        serialization.deserialize(xml.replace('version="1"', 'version="2"'))


class FakeRBS:
    def getNumDOF(self):
        return 12


class FakeIntegrator:
    def getRigidBodySystem(self):
        return FakeRBS()

    def getKineticEnergies(self):
        return [30.0, 12.0]

    def getRefinedKineticEnergies(self):
        return [30.5, 12.5]

    def getPotentialEnergyRefinement(self):
        return -0.25


class CMMotionRemover:
    pass


class FakeState:
    def getTime(self):
        return 1.5

    def getPotentialEnergy(self):
        return -100.0

    def getKineticEnergy(self):
        return 42.0


class FakeContext:
    def getState(self, getEnergy=False):
        return FakeState()


class FakeSimulation:
    def __init__(self, forces=()):
        self.integrator, self.context, self.currentStep = FakeIntegrator(), FakeContext(), 10
        self.system = FakeSystem(6, forces=forces)


def test_state_data_reporter_columns_and_dof():
    out = io.StringIO()
    rep = StateDataReporter(out, 5, step=True, temperature=True, totalEnergy=True, translationalEnergy=True,
                            rotationalEnergy=True, refinedPotentialEnergy=True, refinedKineticEnergy=True,
                            refinedTotalEnergy=True, refinedTemperature=True, refinedTranslationalEnergy=True,
                            refinedRotationalEnergy=True)
    sim = FakeSimulation(forces=[CMMotionRemover()])
    rep.report(sim)
    header, row = out.getvalue().strip().split("\n")
    cols = [c.strip('"') for c in header.lstrip("#").split('","')]
    assert cols == ["Step", "Total Energy (kJ/mole)", "Temperature (K)", "Translational Energy (kJ/mole)",
                    "Rotational Energy (kJ/mole)", "Refined Potential Energy (kJ/mole)", "Refined Kinetic Energy (kJ/mole)",
                    "Refined Total Energy (kJ/mole)", "Refined Temperature (K)", "Refined Translational Energy (kJ/mole)",
                    "Refined Rotational Energy (kJ/mole)"]
    v = [float(x) for x in row.split(",")]
    dof = 12 - 3                                            # rigid-body DOF minus 3 for the CMMotionRemover
    assert v[0] == 10 and v[1] == -58.0
    assert v[2] == pytest.approx(2 * 42.0 / (dof * MOLAR_GAS_CONSTANT_R))
    assert v[3:5] == [30.0, 12.0]
    assert v[5] == -100.25 and v[6] == 43.0 and v[7] == pytest.approx(43.0 - 100.25)
    assert v[8] == pytest.approx(2 * 43.0 / (dof * MOLAR_GAS_CONSTANT_R)) and v[9:] == [30.5, 12.5]
    assert rep.describeNextReport(sim)[0] == 5


def test_host_constraint_solver_chain():
    """SHAKE / RATTLE stand-in used by the Python Context for constraints among free atoms (coupled constraints)."""
    import numpy as np
    from openmm_rigidbody_plugin_b200.integrator import System, _DistanceConstraints
    system = System()
    for m in (12.0, 14.0, 16.0, 1.0):
        system.addParticle(m)
    system.addConstraint(0, 1, 0.11)
    system.addConstraint(1, 2, 0.13)
    solver = _DistanceConstraints(system)
    old = np.array([[0.0, 0.0, 0.0], [0.11, 0.0, 0.0], [0.11, 0.13, 0.0], [1.0, 1.0, 1.0]])
    rng = np.random.default_rng(1)
    new = old + 0.004 * rng.normal(size=old.shape)
    free = new[3].copy()
    com = (np.array([12.0, 14.0, 16.0])[:, None] * new[:3]).sum(0)
    solver.apply(old, new, 1e-12)
    assert abs(np.linalg.norm(new[0] - new[1]) - 0.11) < 1e-12 and abs(np.linalg.norm(new[1] - new[2]) - 0.13) < 1e-12
    assert np.allclose((np.array([12.0, 14.0, 16.0])[:, None] * new[:3]).sum(0), com, atol=1e-13)     # momentum-conserving
    assert np.array_equal(new[3], free)
    V = rng.normal(size=old.shape)
    p = (np.array([12.0, 14.0, 16.0])[:, None] * V[:3]).sum(0)
    solver.applyToVelocities(new, V, 1e-12)
    for a, b in ((0, 1), (1, 2)):
        assert abs(np.dot(new[a] - new[b], V[a] - V[b])) < 1e-12
    assert np.allclose((np.array([12.0, 14.0, 16.0])[:, None] * V[:3]).sum(0), p, atol=1e-13)
