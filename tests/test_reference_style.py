"""The reference's own integration test, restated against this repo's mirror of its API
(tests/TestRigidBodyIntegrator.h:49-83, testSingleBond), plus the API error behaviour of
openmmapi/src/RigidBodyIntegrator.cpp.  The stepping tests need the GPU; the API-contract tests do not."""
import numpy as np
import pytest

from openmm_rigidbody_plugin_b200 import (Context, HarmonicBondForce, OpenMMException, RigidBodyIntegrator, System)


def test_api_contract_without_device():
    integ = RigidBodyIntegrator(0.004, [1, 1, 1, 0])
    assert integ.getStepSize() == 0.004 and integ.getConstraintTolerance() == 1e-5      # RigidBodyIntegrator.cpp:18-24
    assert integ.getRotationMode() == 0 and integ.getComputeRefinedEnergies() is False
    assert integ.getBodyIndices() == [1, 1, 1, 0]
    with pytest.raises(OpenMMException, match="Rotation mode cannot be negative"):
        integ.setRotationMode(-1)
    with pytest.raises(OpenMMException, match="not bound to a context"):
        integ.step(1)
    system = System()
    for m in (16.0, 1.0, 1.0):
        system.addParticle(m)
    with pytest.raises(OpenMMException, match="Number of body indices differs"):
        Context(system, integ)
    system.addParticle(12.0)
    ctx = Context(system, integ)
    with pytest.raises(OpenMMException, match="already bound to a context"):
        integ.setRotationMode(3)
    with pytest.raises(OpenMMException, match="already bound to a context"):
        Context(system, integ)
    assert ctx.getIntegrator() is integ


@pytest.mark.gpu
def test_single_bond_like_the_reference():
    """Two free atoms (bodyIndices all 0), harmonic bond, dt 0.01, 1000 steps vs the analytic solution."""
    system = System()
    system.addParticle(2.0)
    system.addParticle(2.0)
    integrator = RigidBodyIntegrator(0.01, [0, 0])
    bond = HarmonicBondForce()
    bond.addBond(0, 1, 1.5, 1)
    system.addForce(bond)
    context = Context(system, integrator)
    context.setPositions([[-1, 0, 0], [1, 0, 0]])
    freq = 1.0
    state = context.getState(getEnergy=True)
    initial = state.getKineticEnergy() + state.getPotentialEnergy()
    for _ in range(1000):
        state = context.getState(getPositions=True, getVelocities=True, getEnergy=True)
        t = state.getTime()
        dist = 1.5 + 0.5 * np.cos(freq * t)
        assert np.allclose(state.getPositions()[0], [-0.5 * dist, 0, 0], atol=0.02)
        assert np.allclose(state.getPositions()[1], [0.5 * dist, 0, 0], atol=0.02)
        speed = -0.5 * freq * np.sin(freq * t)
        assert np.allclose(state.getVelocities()[0], [-0.5 * speed, 0, 0], atol=0.02)
        assert np.allclose(state.getVelocities()[1], [0.5 * speed, 0, 0], atol=0.02)
        energy = state.getKineticEnergy() + state.getPotentialEnergy()
        assert abs(energy - initial) <= 0.01 * abs(initial)
        integrator.step(1)
    assert abs(context.getState().getTime() - 10.0) < 1e-5


@pytest.mark.gpu
def test_readme_style_rigid_water_run():
    """README.md:191-211 flow on a rigid-water box: createSystem-like body indices -> integrator -> setPositions /
    setVelocities -> step; DOF and kinetic-energy split as the StateDataReporter extension reads them."""
    import common
    sysd = common.synth.water_box(512, seed=3)
    system = System()
    for m in sysd["masses"]:
        system.addParticle(m)
    integrator = RigidBodyIntegrator(0.001, sysd["bodyIndices"].tolist())
    integrator.setRotationMode(0)
    context = Context(system, integrator)
    context.setPositions(sysd["R"])
    context.setVelocities(sysd["V"])
    assert integrator.getRigidBodySystem().getNumDOF() == 6 * 512
    ke0 = integrator.getKineticEnergies()
    integrator.step(50)                                  # no forces: free flight + free rotation conserve both terms
    ke1 = integrator.getKineticEnergies()
    assert abs(ke1[0] - ke0[0]) <= 1e-10 * ke0[0] and abs(ke1[1] - ke0[1]) <= 1e-10 * ke0[1]
    T = 2 * sum(ke1) / (integrator.getRigidBodySystem().getNumDOF() * 0.00831446261815324)
    assert 200.0 < T < 400.0
