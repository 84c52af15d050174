"""State changes around stepping (RigidBodyIntegrator::stateChanged, openmmapi/src/RigidBodyIntegrator.cpp:63-74) and the
host-buffer step call's contract (include/rbk.h, rbk_execute_host)."""
import numpy as np
import pytest

import common
from common import GpuStepper, rel_inf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["water", "mixed"])
@pytest.mark.parametrize("fused", [False, True])
def test_set_velocities_after_stepping_keeps_the_configuration(case, fused):
    """setVelocities(getVelocities()) in the middle of a run must be a no-op up to rounding: the velocities-only rebuild
    has to start from the CURRENT orientation and centre of mass, not from those of the last setPositions."""
    sysd = common.synth.water_box(3000, seed=61) if case == "water" else common.synth.mixed_system(600, 500, seed=62, max_atoms=40)
    a = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
    b = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
    for s in (a, b):
        s.fused = fused
        common.init_like_reference(s, sysd)
    a.step(0.001, 9)
    b.step(0.001, 5)
    R5, V5, _ = b.get_state()
    b.set_state(V=V5)
    b.update(False, True)                                  # stateChanged(Velocities): rbk_update(velocities) + rbk_upload
    R5b, _, _ = b.get_state()
    assert np.array_equal(R5, R5b)
    bb = b.bodies()
    b.step(0.001, 4)
    Ra, Va, _ = a.get_state()
    Rb, Vb, _ = b.get_state()
    assert rel_inf(Rb, Ra) <= 1e-11 and rel_inf(Vb, Va) <= 1e-10, (rel_inf(Rb, Ra), rel_inf(Vb, Va))
    assert rel_inf(b.kinetic(), a.kinetic()) <= 1e-11
    assert np.isfinite(bb["q"]).all()


def test_velocity_rescaling_after_device_build():
    """The same through the GPU-side build: rbk_update_device(geometry) ... steps ... host-side velocities-only rebuild."""
    import torch
    from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem
    sysd = common.synth.water_box(2000, seed=63)
    dev = torch.device("cuda:0")
    s = DeviceRigidBodySystem(sysd["bodyIndices"], sysd["masses"], 0)
    R, V, F = (torch.from_numpy(sysd[k].copy()).to(dev) for k in ("R", "V", "F"))
    s.update_device(R, V, F)
    for _ in range(4):
        s.part1(0.001, R, V, F); s.part2(0.001, R, V, F)
    ke = s.kinetic(V)
    before = s.download_bodies()
    Vh = V.cpu().numpy() * 0.5                             # velocity rescaling: KE must drop to a quarter, nothing else moves
    s.update(V=Vh, geometry=False, velocities=True)
    s.upload()
    after = s.download_bodies()
    assert np.array_equal(after["rcm"], before["rcm"]) and np.array_equal(after["q"], before["q"])
    V.copy_(torch.from_numpy(Vh).to(dev))
    assert rel_inf(s.kinetic(V), 0.25 * ke) <= 1e-12
    assert rel_inf(after["pcm"], 0.5 * before["pcm"]) <= 1e-12 and rel_inf(after["pi"], 0.5 * before["pi"]) <= 1e-11


def _pinned(sysd):
    import torch
    return tuple(torch.from_numpy(sysd[k].copy()).pin_memory() for k in ("R", "V", "F"))


@pytest.mark.parametrize("case", ["water", "mixed"])
def test_execute_host_multi_step_and_lazy_velocities(case):
    """rbk_execute_host: (1) n steps in one call with a host force callback (fused with rbk_part2_part1 where the system
    allows it) = n one-step calls = the device path; (2) V = NULL leaves the velocities on the device, a later call with
    V returns them; (3) kinetic_host after such a call uses the device copy, not the stale host array."""
    sysd = common.synth.water_box(1500, seed=64) if case == "water" else common.synth.mixed_system(400, 600, seed=65)
    sysd = dict(sysd, R=np.vstack([sysd["R"]]), V=sysd["V"])
    n = len(sysd["masses"])
    ref = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
    common.init_like_reference(ref, sysd)
    ref.step(0.001, 6)
    Rr, Vr, _ = ref.get_state()

    calls = []

    def forces(Rp, Fp, count, user):
        calls.append(count)

    a = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
    common.init_like_reference(a, sysd)
    R, V, F = _pinned(sysd)
    a.sys.execute_host(0.001, 4, R, V, F, forces=forces)
    a.sys.execute_host(0.001, 2, R, V, F, forces=forces)
    assert calls == [n] * 6
    assert rel_inf(R.numpy(), Rr) <= 1e-12 and rel_inf(V.numpy(), Vr) <= 1e-11

    b = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
    common.init_like_reference(b, sysd)
    R, V, F = _pinned(sysd)
    V0 = V.numpy().copy()
    b.sys.execute_host(0.001, 1, R, V, F)
    for _ in range(4):
        b.sys.execute_host(0.001, 1, R, None, F)
    V1 = V.numpy().copy()                                  # still the velocities of the first call
    ke_dev = b.sys.kinetic_host(V)                         # must come from the device copy (free atoms!)
    b.sys.execute_host(0.001, 1, R, V, F)
    assert rel_inf(R.numpy(), Rr) <= 1e-12 and rel_inf(V.numpy(), Vr) <= 1e-11
    assert not np.array_equal(V1, V.numpy()) and not np.array_equal(V0, V1)
    c = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
    common.init_like_reference(c, sysd)
    c.step(0.001, 5)
    assert rel_inf(ke_dev, c.kinetic()) <= 1e-11

    # forces == NULL, several steps in one call: F serves every step, R and V come back once
    d = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
    common.init_like_reference(d, sysd)
    R, V, F = _pinned(sysd)
    d.sys.execute_host(0.001, 6, R, V, F)
    assert rel_inf(R.numpy(), Rr) <= 1e-12 and rel_inf(V.numpy(), Vr) <= 1e-11


def test_execute_host_reads_the_callers_forces_every_call():
    """A host-side change of F between two calls (updateParametersInContext + getState in OpenMM terms) must reach
    Part 1 of the next call: the reference reads data.forces every step (ReferenceRigidBodyKernels.cpp:96)."""
    sysd = common.synth.mixed_system(300, 400, seed=66, max_atoms=20)
    F2 = sysd["F"] * -0.5

    def forces_keep(Rp, Fp, count, user):
        pass

    a = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
    common.init_like_reference(a, sysd)
    R, V, F = _pinned(sysd)
    a.sys.execute_host(0.001, 2, R, V, F, forces=forces_keep)
    F.numpy()[:] = F2                                      # the caller re-evaluated the forces
    a.sys.execute_host(0.001, 2, R, V, F, forces=forces_keep)

    b = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
    common.init_like_reference(b, sysd)
    b.step(0.001, 2)
    b.set_state(F=F2)
    b.step(0.001, 2)
    Rb, Vb, _ = b.get_state()
    assert rel_inf(R.numpy(), Rb) <= 1e-12 and rel_inf(V.numpy(), Vb) <= 1e-11
