"""GPU-side body build (rbk_update_device) against the CPU oracle (RigidBodySystem::update
restated, oracle/rb_oracle.c, bit-identical to the reference sources) and against the host model."""
import numpy as np
import pytest

import common
from common import quat_rel, rel_inf
from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem, RbkError

pytestmark = pytest.mark.gpu


def host_built(sysd, mode):
    n = len(sysd["masses"])
    s = DeviceRigidBodySystem(sysd["bodyIndices"], sysd["masses"], mode)
    s.update(sysd["R"], np.zeros((n, 3)), sysd["F"], True, True)
    s.update(V=sysd["V"], geometry=False, velocities=True)
    s.upload()
    return s


@pytest.mark.parametrize("case", ["water", "mixed", "edge"])
def test_device_build_matches_host_build(case):
    import torch
    dev = torch.device("cuda:0")
    if case == "water":
        sysd = common.synth.water_box(5000, seed=71)
    elif case == "mixed":
        sysd = common.synth.mixed_system(1500, 2000, seed=72)
    else:   # linear bodies (dof 5), symmetric top, labels with gaps
        parts = [common.edge_cases()[k][0] for k in ("linear_bodies", "symmetric_top", "gaps_unordered")]
        off, body = 0, []
        for p in parts:
            b = p["bodyIndices"].copy()
            b[b > 0] += off
            off = max(off, int(b.max()))
            body.append(b)
        sysd = {k: np.concatenate([p[k] for p in parts]) for k in ("masses", "R", "V", "F", "charges")}
        sysd["bodyIndices"] = np.concatenate(body).astype(np.int32)
    mode = 2 if case == "edge" else 0
    a = host_built(sysd, mode)
    ha = a.host_bodies()
    da = a.download_bodies()

    b = DeviceRigidBodySystem(sysd["bodyIndices"], sysd["masses"], mode)
    R = torch.from_numpy(sysd["R"]).to(dev)
    V = torch.from_numpy(sysd["V"]).to(dev)
    F = torch.from_numpy(sysd["F"]).to(dev)
    Z = torch.zeros_like(V)
    b.update_device(R, Z, F, geometry=True, velocities=True)          # setPositions (velocities still zero)
    b.update_device(vel=V, geometry=False, velocities=True)           # setVelocities
    assert b.counts() == a.counts()                                   # incl. numDOF (5 for linear bodies)
    db = b.download_bodies()
    tol = 1e-12
    for k in ("rcm", "pcm", "pi", "force", "torque"):
        assert rel_inf(db[k], da[k]) <= tol, (k, rel_inf(db[k], da[k]))
    assert quat_rel(db["q"], da["q"]) <= tol
    with pytest.raises(RbkError, match="built on the device"):
        b.host_bodies()
    # the stepped trajectories agree as well (I, 1/I and the body-frame coordinates enter here)
    Ra, Va, Rb, Vb = R.clone(), V.clone(), R.clone(), V.clone()
    for _ in range(3):
        a.part1(0.001, Ra, Va, F); a.part2(0.001, Ra, Va, F)
        b.part1(0.001, Rb, Vb, F); b.part2(0.001, Rb, Vb, F)
    assert float((Ra - Rb).abs().max()) <= 1e-11 * float(Ra.abs().max())
    assert float((Va - Vb).abs().max()) <= 1e-10 * float(Va.abs().max())
    assert np.allclose(a.kinetic(Va), b.kinetic(Vb), rtol=1e-11)
    assert ha["dof"].sum() + a.counts()["numFree"] == a.counts()["numDOF"]
    # ... and directly against the oracle: body state right after the build, trajectory after three steps
    from oracle.checkers import CpuStepper
    o = CpuStepper("oracle", sysd["bodyIndices"], sysd["masses"], mode)
    common.init_like_reference(o, sysd)
    ob = o.bodies()
    for k in ("rcm", "pcm", "pi", "force", "torque"):
        assert rel_inf(db[k], ob[k]) <= tol, ("oracle", k, rel_inf(db[k], ob[k]))
    assert quat_rel(db["q"], ob["q"]) <= tol
    assert b.counts()["numDOF"] == o.counts()["numDOF"]
    o.step(0.001, 3)
    Ro, Vo, _ = o.get_state()
    assert rel_inf(Rb.cpu().numpy(), Ro) <= 2e-10 and rel_inf(Vb.cpu().numpy(), Vo) <= 2e-10
    assert rel_inf(b.kinetic(Vb), o.kinetic()) <= 2e-10


def test_device_build_1M_waters_is_fast():
    import time
    import torch
    dev = torch.device("cuda:0")
    sysd = common.synth.water_box(1_000_000, seed=73)
    s = DeviceRigidBodySystem(sysd["bodyIndices"], sysd["masses"], 0)
    R = torch.from_numpy(sysd["R"]).to(dev)
    V = torch.from_numpy(sysd["V"]).to(dev)
    F = torch.from_numpy(sysd["F"]).to(dev)
    s.update_device(R, V, F)                       # first call allocates
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.update_device(R, V, F)
    torch.cuda.synchronize()
    dt_dev = time.perf_counter() - t0
    assert s.counts()["numDOF"] == 6_000_000
    assert dt_dev < 0.05, dt_dev                   # the host rebuild + upload of the same system takes seconds
    Rm = R.clone(); Vm = V.clone()
    s.part1(0.001, Rm, Vm, F); s.part2(0.001, Rm, Vm, F)
    assert bool(torch.isfinite(Rm).all()) and bool(torch.isfinite(Vm).all())
    print(f"device body build of 1M waters: {1e3*dt_dev:.2f} ms")
