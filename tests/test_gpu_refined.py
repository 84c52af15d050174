"""Refined ("shadow") energy diagnostics: RigidBodyIntegrator::setComputeRefinedEnergies /
getRefinedKineticEnergies / getPotentialEnergyRefinement.

The reference computes them only on its CUDA platform (COMPMOD paths of
platforms/cuda/src/kernels/rigidbodyintegrator.cu, host side CudaRigidBodyKernels.cpp:118-194,405-438,481-494), which
cannot run here: the CPU checker is this repo's fp64 restatement of those paths (oracle/rb_oracle.c, PARITY
UNPINNED), and the second test pins the semantics physically instead - the refined total energy of rigid bodies is
a shadow Hamiltonian, conserved one order better than the plain one."""
import numpy as np
import pytest

import common
from common import GpuStepper
from oracle.checkers import CpuStepper

pytestmark = pytest.mark.gpu


def charges_for(sysd, seed=2):
    rng = np.random.Generator(np.random.Philox(key=seed))
    return rng.uniform(-0.5, 0.5, len(sysd["masses"]))


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("fused", [False, True])
def test_refined_energies_match_cpu_restatement(mode, fused):
    sysd = common.synth.mixed_system(150, 120, seed=21, max_atoms=14)
    ch = charges_for(sysd)
    dt, steps = 0.002, 6
    steppers = []
    for make in (lambda: CpuStepper("oracle", sysd["bodyIndices"], sysd["masses"], mode),
                 lambda: GpuStepper(sysd["bodyIndices"], sysd["masses"], mode)):
        st = make()
        st.set_tether(800.0, (30.0, -50.0, 80.0), ch, sysd["R"] * 0.98)
        st.set_state(sysd["R"], sysd["V"], sysd["F"])
        st.compute_forces()
        st.update(True, True)
        steppers.append(st)
    o, g = steppers
    g.fused = fused
    o.set_refined(True)
    g.sys.set_refined_energies(1)
    # before any step the accumulators are empty
    assert np.all(g.sys.refined_kinetic(dt, g.dV) == 0.0)
    for st in (o, g):
        st.step(dt, steps)
    ko, kg = o.refined_kinetic(dt), g.sys.refined_kinetic(dt, g.dV)
    uo, ug = o.potential_refinement(dt), g.sys.potential_refinement(dt, g.dF)
    assert common.rel_inf(kg, ko) < 1e-9, (kg, ko)
    assert abs(ug - uo) <= 1e-11 * abs(uo), (ug, uo)
    # refined estimates are close to, but not the same as, the plain energies
    plain = g.kinetic()
    assert 1e-7 < abs(kg[1] - plain[1]) / plain[1] < 0.2
    # the diagnostics do not disturb the trajectory
    Ro, Vo, _ = o.get_state()
    Rg, Vg, _ = g.get_state()
    assert common.rel_inf(Rg, Ro) < 1e-10 and common.rel_inf(Vg, Vo) < 1e-9
    h = GpuStepper(sysd["bodyIndices"], sysd["masses"], mode)
    h.set_tether(800.0, (30.0, -50.0, 80.0), ch, sysd["R"] * 0.98)
    h.set_state(sysd["R"], sysd["V"], sysd["F"])
    h.compute_forces()
    h.update(True, True)
    h.step(dt, steps)
    Rh, Vh, _ = h.get_state()
    assert np.array_equal(Rh, Rg) and np.array_equal(Vh, Vg)
    # layouts / host entry points agree with the device-array call
    import torch
    assert np.array_equal(g.sys.refined_kinetic(dt, g.dV.t().contiguous()), kg)


@pytest.mark.parametrize("mode", [0, 4])
def test_refined_total_energy_is_a_shadow_hamiltonian(mode):
    """Rigid waters in the tether + field potential: E_refined = KE_refined + U + dU fluctuates an order of
    magnitude less than E = KE + U, and its fluctuation shrinks faster with dt (the point of the diagnostic)."""
    sysd = common.synth.water_box(216, seed=6)
    n = len(sysd["masses"])
    ch = np.tile([-0.834, 0.417, 0.417], n // 3)
    stds = {}
    for dt in (0.001, 0.002):
        g = GpuStepper(sysd["bodyIndices"], sysd["masses"], mode)
        g.set_tether(5000.0, (300.0, -500.0, 800.0), ch, sysd["R"])
        g.set_state(sysd["R"], sysd["V"], sysd["F"])
        g.compute_forces()
        g.update(True, True)
        g.sys.set_refined_energies(1)
        plain, refined = [], []
        for _ in range(300):
            g.part1(dt)
            U = g.compute_forces()
            g.part2(dt)
            plain.append(g.kinetic().sum() + U)
            refined.append(g.sys.refined_kinetic(dt, g.dV).sum() + U + g.sys.potential_refinement(dt, g.dF))
        stds[dt] = (np.std(plain), np.std(refined))
        assert stds[dt][1] < 0.2 * stds[dt][0], stds
    assert stds[0.002][0] / stds[0.001][0] < stds[0.002][1] / stds[0.001][1]


def test_free_atom_term_reproduces_the_reference_factors():
    """Free atoms in uniform motion (no forces): the reference's stencil factors (-1, 5, 2) give posDot = 8 v dt, i.e.
    4/3 of the kinetic energy - a reference quirk (the consistent stencil needs +1; see oracle/rb_oracle.c) that is
    reproduced, and that a caller can avoid by driving rbk_free_dot_openmm with its own factors."""
    import torch
    from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem
    dev = torch.device("cuda:0")
    rng = np.random.Generator(np.random.Philox(key=9))
    n, dt = 64, 0.002
    masses = rng.uniform(1.0, 16.0, n)
    V = rng.normal(size=(n, 3))
    s = DeviceRigidBodySystem(np.zeros(n, np.int32), masses, 0)
    s.update(rng.normal(size=(n, 3)), V, np.zeros((n, 3)), True, True)
    s.upload()
    s.set_refined_energies(1)
    R, Vd, F = (torch.from_numpy(rng.normal(size=(n, 3))).to(dev), torch.from_numpy(V).to(dev),
                torch.zeros(n, 3, dtype=torch.float64, device=dev))
    s.part1(dt, R, Vd, F)
    s.part2(dt, R, Vd, F)
    ke = 0.5 * float(np.sum(masses[:, None] * V * V))
    assert abs(s.refined_kinetic(dt, Vd)[0] - 4.0 / 3.0 * ke) < 1e-12 * ke
    assert s.potential_refinement(dt, F) == 0.0
    # the caller-driven accumulation with the consistent factor (+1) returns the kinetic energy itself
    s.set_refined_energies(2)
    velm = torch.zeros(n, 4, dtype=torch.float64, device=dev)
    velm[:, :3] = Vd
    velm[:, 3] = torch.from_numpy(1.0 / masses).to(dev)
    force = torch.zeros(3, n, dtype=torch.int64, device=dev)
    delta = torch.zeros(n, 4, dtype=torch.float64, device=dev)
    s.free_delta_openmm(-dt, velm, force, n, 2, delta)
    s.free_dot_openmm(delta, 2, 1.0, True)
    s.free_delta_openmm(dt, velm, force, n, 2, delta)
    s.free_dot_openmm(delta, 2, 5.0, False)
    s.free_dot_openmm(delta, 2, 2.0, False)
    assert abs(s.refined_kinetic_openmm(dt, velm, 2)[0] - ke) < 1e-12 * ke


def test_errors_and_python_api():
    from openmm_rigidbody_plugin_b200 import Context, DeviceRigidBodySystem, HarmonicBondForce, RbkError, RigidBodyIntegrator, System
    import torch
    s = DeviceRigidBodySystem([1, 1, 1, 0], [1.0, 2.0, 3.0, 4.0], 0)
    with pytest.raises(RbkError, match="rbk_upload first"):
        s.set_refined_energies(1)
    rng = np.random.Generator(np.random.Philox(key=1))
    s.update(rng.normal(size=(4, 3)), rng.normal(size=(4, 3)), np.zeros((4, 3)), True, True)
    s.upload()
    x = torch.zeros(4, 3, dtype=torch.float64, device="cuda")
    with pytest.raises(RbkError, match="not enabled"):
        s.refined_kinetic(0.001, x)
    with pytest.raises(RbkError, match="unknown mode"):
        s.set_refined_energies(7)

    sysd = common.synth.water_box(27, seed=4)
    system = System()
    for m in sysd["masses"]:
        system.addParticle(float(m))
    bonds = HarmonicBondForce()
    for m in range(26):
        bonds.addBond(3 * m, 3 * (m + 1), 0.31, 1000.0)
    system.addForce(bonds)
    integ = RigidBodyIntegrator(0.002, list(sysd["bodyIndices"]))
    integ.setComputeRefinedEnergies(True)
    ctx = Context(system, integ)
    ctx.setPositions(sysd["R"])
    ctx.setVelocities(sysd["V"])
    assert integ.getRefinedKineticEnergies() == integ.getKineticEnergies()      # nothing accumulated yet
    assert integ.getPotentialEnergyRefinement() == 0.0
    plain, refined = [], []
    for _ in range(100):
        integ.step(1)
        st = ctx.getState(getEnergy=True)
        plain.append(st.getKineticEnergy() + st.getPotentialEnergy())
        refined.append(sum(integ.getRefinedKineticEnergies()) + st.getPotentialEnergy() + integ.getPotentialEnergyRefinement())
    assert integ.getPotentialEnergyRefinement() < 0.0
    assert np.std(refined) < 0.3 * np.std(plain)
