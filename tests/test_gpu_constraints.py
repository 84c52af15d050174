"""Free-atom constraint / virtual-site hooks.

Host flow: rbk_execute_host_hooks = ReferenceIntegrateRigidBodyStepKernel::execute with constraints.apply after
Part 1 and constraints.applyToVelocities after Part 2 (platforms/reference/src/ReferenceRigidBodyKernels.cpp:92-104).
The checker runs the SAME numpy SHAKE/RATTLE between the oracle's part1 / part2.
Device flow: rbk_free_delta_openmm -> (constraint solver on posDelta) -> rbk_part1_delta_openmm, the CUDA
platform's sequence (platforms/cuda/src/CudaRigidBodyKernels.cpp:405-421)."""
import ctypes as C

import numpy as np
import pytest

import common
from common import GpuStepper
from oracle.checkers import CpuStepper

pytestmark = pytest.mark.gpu


def pair_system(seed=11, n_bodies=150):
    """Rigid bodies + free atoms, the first 200 free atoms bonded pairwise by distance constraints.  n_bodies = 150: few atom
    tiles, the free atoms get their own launch; 600: the free atoms ride along in the large-body Part 2 kernel."""
    sysd = common.synth.mixed_system(n_bodies, 500, seed=seed, max_atoms=20)
    free = np.flatnonzero(np.asarray(sysd["bodyIndices"]) <= 0)
    pairs = free[:200].reshape(-1, 2).astype(np.int32)
    R = sysd["R"]
    # put each partner 0.1 nm from its mate so that the constraint length is sensible
    rng = np.random.Generator(np.random.Philox(key=seed))
    u = rng.normal(size=(len(pairs), 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    R[pairs[:, 1]] = R[pairs[:, 0]] + 0.1 * u
    d0 = np.full(len(pairs), 0.1)
    return sysd, pairs, d0


def shake(old, new, pairs, d0, invm, iters=30):
    a, b = pairs[:, 0], pairs[:, 1]
    r0 = old[a] - old[b]
    wa, wb = invm[a][:, None], invm[b][:, None]
    for _ in range(iters):
        r = new[a] - new[b]
        diff = d0 ** 2 - np.einsum("ij,ij->i", r, r)
        g = diff / (2.0 * (wa + wb)[:, 0] * np.einsum("ij,ij->i", r0, r))
        new[a] += (g[:, None] * wa) * r0
        new[b] -= (g[:, None] * wb) * r0


def rattle(R, V, pairs, d0, invm):
    a, b = pairs[:, 0], pairs[:, 1]
    r = R[a] - R[b]
    g = np.einsum("ij,ij->i", r, V[a] - V[b]) / ((invm[a] + invm[b]) * d0 ** 2)
    V[a] -= (g * invm[a])[:, None] * r
    V[b] += (g * invm[b])[:, None] * r


def tether_forces(R, R0, k=2000.0):
    return -k * (R - R0)


@pytest.mark.parametrize("n_bodies", [150, 600])
@pytest.mark.parametrize("mode", [0, 3])
def test_host_hooks_match_oracle_with_same_constraint_solver(mode, n_bodies):
    import torch
    sysd, pairs, d0 = pair_system(n_bodies=n_bodies)
    n = len(sysd["masses"])
    invm = 1.0 / np.asarray(sysd["masses"])
    R0 = sysd["R"].copy()
    # project the initial velocities onto the constraint manifold
    rattle(sysd["R"], sysd["V"], pairs, d0, invm)
    sysd["F"] = tether_forces(sysd["R"], R0)

    o = CpuStepper("oracle", sysd["bodyIndices"], sysd["masses"], mode, constraints=pairs)
    g = GpuStepper(sysd["bodyIndices"], sysd["masses"], mode, constraints=pairs)
    assert g.sys.counts()["numDOF"] == o.counts()["numDOF"]
    for st in (o, g):
        common.init_like_reference(st, sysd)

    dt, steps = 0.002, 5
    for _ in range(steps):
        Rold, _, _ = o.get_state()
        o.part1(dt)
        R, V, _ = o.get_state()
        shake(Rold, R, pairs, d0, invm)
        o.set_state(R=R, F=tether_forces(R, R0))
        o.part2(dt)
        R, V, _ = o.get_state()
        rattle(R, V, pairs, d0, invm)
        o.set_state(V=V)
    Ro, Vo, _ = o.get_state()

    R = torch.from_numpy(sysd["R"].copy()).pin_memory()
    V = torch.from_numpy(sysd["V"].copy()).pin_memory()
    F = torch.from_numpy(sysd["F"].copy()).pin_memory()
    seen = {"pos": 0, "vel": 0}

    def view(ptr):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n, 3))

    def forces(Rp, Fp, count, user):
        view(Fp)[:] = tether_forces(view(Rp), R0)

    def constrain_positions(oldp, newp, count, user):
        seen["pos"] += 1
        shake(view(oldp), view(newp), pairs, d0, invm)
        return 1

    def constrain_velocities(Rp, Vp, count, user):
        seen["vel"] += 1
        rattle(view(Rp), view(Vp), pairs, d0, invm)
        return 1

    g.sys.execute_host(dt, steps, R, V, F, forces=forces, constrain_positions=constrain_positions,
                       constrain_velocities=constrain_velocities)
    assert seen == {"pos": steps, "vel": steps}
    Rg, Vg = R.numpy(), V.numpy()
    assert common.rel_inf(Rg, Ro) < 1e-10 and common.rel_inf(Vg, Vo) < 1e-9
    # the constraints hold on the device result, and the constrained pairs did move
    d = np.linalg.norm(Rg[pairs[:, 0]] - Rg[pairs[:, 1]], axis=1)
    assert np.max(np.abs(d - d0)) < 1e-12
    assert np.max(np.abs(Rg[pairs[:, 0]] - sysd["R"][pairs[:, 0]])) > 1e-4
    rv = np.einsum("ij,ij->i", Rg[pairs[:, 0]] - Rg[pairs[:, 1]], Vg[pairs[:, 0]] - Vg[pairs[:, 1]])
    assert np.max(np.abs(rv)) < 1e-12
    # kinetic energies through the host entry point agree with the oracle's
    assert common.rel_inf(g.sys.kinetic_host(V), o.kinetic()) < 1e-10


def test_hooks_that_change_nothing_equal_the_plain_call():
    import torch
    sysd = common.synth.mixed_system(200, 300, seed=3, max_atoms=12)
    out = []
    for hooks in (False, True):
        g = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
        common.init_like_reference(g, sysd)
        R = torch.from_numpy(sysd["R"].copy()).pin_memory()
        V = torch.from_numpy(sysd["V"].copy()).pin_memory()
        F = torch.from_numpy(sysd["F"].copy()).pin_memory()
        kw = dict(constrain_positions=lambda a, b, n, u: 0, constrain_velocities=lambda a, b, n, u: 0) if hooks else {}
        g.sys.execute_host(0.001, 3, R, V, F, **kw)
        out.append((R.numpy().copy(), V.numpy().copy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("n_bodies", [120, 500])       # 500: enough atom tiles for the free atoms to ride along in Part 2
@pytest.mark.parametrize("precision", [0, 1, 2])
def test_device_delta_hooks_openmm_formats(precision, n_bodies):
    """delta pre-pass + part1_delta == plain part1 when the solver leaves posDelta alone, and a modified posDelta
    moves the free atoms by exactly that displacement; what the solver changed reaches the velocities in part 2 as
    (x - savedPos)/dt, the Reference platform's arithmetic (RigidBodySystem.cpp:196-197)."""
    import torch
    from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem
    dev = torch.device("cuda:0")
    sysd = common.synth.mixed_system(n_bodies, 260, seed=8, max_atoms=16)
    n = len(sysd["masses"])
    padded = (n + 31) // 32 * 32
    free = np.flatnonzero(np.asarray(sysd["bodyIndices"]) <= 0)
    real = torch.float32 if precision == 0 else torch.float64
    mixed = torch.float32 if precision == 0 else torch.float64

    def fresh():
        s = DeviceRigidBodySystem(sysd["bodyIndices"], sysd["masses"], 0)
        s.update(sysd["R"], sysd["V"], sysd["F"], True, True)
        s.upload()
        R = torch.from_numpy(sysd["R"]).to(dev)
        if precision == 1:
            posq = torch.zeros(padded, 4, dtype=torch.float32, device=dev)
            corr = torch.zeros(padded, 4, dtype=torch.float32, device=dev)
            posq[:n, :3] = R.float()
            corr[:n, :3] = (R - posq[:n, :3].double()).float()
        else:
            posq = torch.zeros(padded, 4, dtype=real, device=dev)
            posq[:n, :3] = R.to(real)
            corr = None
        velm = torch.zeros(padded, 4, dtype=mixed, device=dev)
        velm[:n, :3] = torch.from_numpy(sysd["V"]).to(dev).to(mixed)
        velm[:n, 3] = torch.from_numpy(1.0 / np.asarray(sysd["masses"])).to(dev).to(mixed)
        force = torch.zeros(3, padded, dtype=torch.int64, device=dev)
        force[:, :n] = torch.round(torch.from_numpy(sysd["F"]).to(dev).t() * 4294967296.0).to(torch.int64)
        return s, posq, corr, velm, force

    def positions(posq, corr):
        x = posq[:n, :3].double()
        return (x + corr[:n, :3].double()) if corr is not None else x

    dt = 0.001
    a, posqA, corrA, velmA, forceA = fresh()
    a.part1_openmm(dt, posqA, corrA, velmA, forceA, padded, precision)

    b, posqB, corrB, velmB, forceB = fresh()
    delta = torch.full((padded, 4), 7.0, dtype=mixed, device=dev)
    b.free_delta_openmm(dt, velmB, forceB, padded, precision, delta)
    torch.cuda.synchronize()
    body_rows = np.setdiff1d(np.arange(padded), free)
    assert torch.all(delta[torch.from_numpy(body_rows).to(dev)] == 7.0)           # only free atoms are written
    assert torch.all(delta[torch.from_numpy(free).to(dev), 3] == 7.0)             # .w untouched
    b.part1_delta_openmm(dt, posqB, corrB, velmB, forceB, padded, precision, delta)
    torch.cuda.synchronize()
    tol = 2e-7 if precision == 0 else (1e-14 if precision == 2 else 2e-14)
    assert common.rel_inf(positions(posqB, corrB).cpu().numpy(), positions(posqA, corrA).cpu().numpy()) <= tol
    assert torch.equal(velmB, velmA)

    # a solver that shifts every free displacement: positions follow, and part 2 turns the shift into velocity
    c, posqC, corrC, velmC, forceC = fresh()
    c.free_delta_openmm(dt, velmC, forceC, padded, precision, delta)
    fr = torch.from_numpy(free).to(dev)
    shift = torch.tensor([1e-3, -2e-3, 5e-4], dtype=mixed, device=dev)
    delta[fr, :3] += shift
    x0 = positions(posqC, corrC).clone()
    c.part1_delta_openmm(dt, posqC, corrC, velmC, forceC, padded, precision, delta)
    x1 = positions(posqC, corrC)
    moved = (x1 - x0)[fr].cpu().numpy()
    # (x1 - x0 carries the rounding of both positions in the arrays' own precision: float32 | float32 + float32 residual | float64)
    xmax = float(np.abs(sysd["R"]).max()) + 0.01
    ulp32 = float(np.spacing(np.float32(xmax)))
    atol = 1.3*ulp32 if precision == 0 else (4.0*float(np.spacing(np.float32(0.5*ulp32))) if precision == 1 else 4.0*float(np.spacing(xmax)))
    assert np.allclose(moved, delta[fr, :3].double().cpu().numpy(), rtol=0, atol=atol)
    vBefore = velmC.clone()
    forceC.zero_()
    c.part2_openmm(dt, posqC, corrC, velmC, forceC, padded, precision)
    torch.cuda.synchronize()
    dv = (velmC[fr, :3] - vBefore[fr, :3]).double().cpu().numpy()                 # zero force: only (x - savedPos)/dt
    assert np.allclose(dv, (shift.double()/dt).cpu().numpy()[None, :], rtol=0, atol=max(1e-3, 2.0*ulp32/dt) if precision == 0 else 1e-9)
    assert torch.equal(velmC[fr, 3], vBefore[fr, 3])
    # ... and without a solver (b above) part 2 adds exactly nothing
    vB = velmB.clone()
    forceB.zero_()
    b.part2_openmm(dt, posqB, corrB, velmB, forceB, padded, precision)
    if precision != 0:
        assert torch.equal(velmB[fr], vB[fr])


def test_python_context_with_constrained_free_atoms():
    """The Python mirror of the plugin API: rigid waters + a constrained triatomic chain of free atoms."""
    from openmm_rigidbody_plugin_b200 import Context, HarmonicBondForce, RigidBodyIntegrator, System
    sysd = common.synth.water_box(27, seed=4)
    system = System()
    for m in sysd["masses"]:
        system.addParticle(float(m))
    base = len(sysd["masses"])
    chain = [system.addParticle(12.0), system.addParticle(14.0), system.addParticle(16.0)]
    system.addConstraint(chain[0], chain[1], 0.11)
    system.addConstraint(chain[1], chain[2], 0.13)
    bonds = HarmonicBondForce()
    bonds.addBond(chain[0], 0, 0.5, 500.0)
    bonds.addBond(chain[2], 3, 0.5, 500.0)
    for m in range(26):
        bonds.addBond(3 * m, 3 * (m + 1), 0.31, 1000.0)
    system.addForce(bonds)
    integ = RigidBodyIntegrator(0.001, list(sysd["bodyIndices"]) + [0, 0, 0])
    integ.setConstraintTolerance(1e-10)
    ctx = Context(system, integ)
    R = np.vstack([sysd["R"], sysd["R"][0] + [[0.3, 0.3, 0.3], [0.41, 0.3, 0.3], [0.41, 0.43, 0.3]]])
    V = np.vstack([sysd["V"], [[0.1, 0.0, 0.2], [0.1, 0.0, 0.2], [0.1, 0.0, 0.2]]])
    ctx.setPositions(R)
    ctx.setVelocities(V)
    assert integ.getRigidBodySystem().getNumDOF() == 3 - 2 + 6 * 27
    st = ctx.getState(getEnergy=True)
    e0 = st.getKineticEnergy() + st.getPotentialEnergy()
    for _ in range(5):
        integ.step(100)
        st = ctx.getState(getPositions=True, getVelocities=True, getEnergy=True)
        P, W = st.getPositions(), st.getVelocities()
        for (a, b, d) in ((chain[0], chain[1], 0.11), (chain[1], chain[2], 0.13)):
            r = P[a] - P[b]
            assert abs(np.linalg.norm(r) - d) < 1e-9
            assert abs(np.dot(r, W[a] - W[b])) < 1e-9
        assert abs(st.getKineticEnergy() + st.getPotentialEnergy() - e0) < 2e-3 * max(1.0, abs(e0))
    assert np.linalg.norm(P[base] - R[base]) > 1e-3
