"""GPU tests of the large-body bucket (mean body size > 8 atoms): part2LargeKernel (persistent cp.async pipeline, arms
staged in shared memory, per-body sums in atom order) and the unrolled atomPositionKernel, against the CPU oracle.

Cases the mixed-system tests elsewhere do not reach: no free atoms at all (identity atom map: the kernels compute slots
instead of reading atomLoc), runs of tiny bodies next to big ones (many bodies per 256-atom tile: several passes of the
8-lanes-per-body reduction and a large per-tile body stage), bodies exactly at / just over the tile capacity (the
CTA-wide single-body path), SoA planes, permuted atoms, and both stepping protocols (part1/part2 and part2_part1).
Reference behaviour: RigidBody::forceAndTorque / updateAtomicVelocities / updateAtomicPositions
(openmmapi/src/RigidBody.cpp:148-183) driven as in ReferenceRigidBodyKernels.cpp:82-108.
"""
import numpy as np
import pytest

import common
from common import GpuStepper
from oracle.checkers import CpuStepper
from test_gpu_parity import compare_state

pytestmark = pytest.mark.gpu


def _system(sizes, n_free, seed):
    """Bodies of the given sizes (Gaussian clouds) followed by n_free free atoms; labels 1..nb."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    sizes = np.asarray(sizes)
    body = np.concatenate([np.repeat(np.arange(1, len(sizes) + 1), sizes), np.zeros(n_free, np.int64)]).astype(np.int32)
    n = body.shape[0]
    centres = rng.uniform(0, 30, (len(sizes) + 1, 3))
    R = centres[np.where(body > 0, body - 1, len(sizes))] + rng.standard_normal((n, 3)) * 0.2
    if n_free:
        R[body == 0] = rng.uniform(0, 30, (n_free, 3))
    return {"bodyIndices": body, "masses": rng.uniform(1.0, 16.0, n), "R": R, "V": rng.standard_normal((n, 3)) * 0.3,
            "F": rng.standard_normal((n, 3)) * 100.0, "charges": rng.uniform(-0.5, 0.5, n)}


def _run_pair(sysd, mode, layout, shuffle, fused, steps=3, dt=0.001, tol=2e-9):
    o = CpuStepper("oracle", sysd["bodyIndices"], sysd["masses"], mode)
    s = GpuStepper(sysd["bodyIndices"], sysd["masses"], mode, layout=layout, shuffle=shuffle)
    s.fused = fused
    for st in (o, s):
        common.init_like_reference(st, sysd, tether=True)
        st.step(dt, steps)
    R, V, _ = o.get_state()
    compare_state(("large", mode, layout, shuffle, fused), s, R, V, o.kinetic(), o.bodies(), tol=tol)
    out = s.get_state()
    s.close()
    return out


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("layout", ["vec3", "soa"])
def test_identity_atom_map_no_free_atoms(mode, layout):
    rng = np.random.Generator(np.random.Philox(key=301))
    sysd = _system(rng.integers(9, 61, size=500), 0, seed=302)
    _run_pair(sysd, mode, layout, False, fused=True)
    _run_pair(sysd, mode, layout, False, fused=False)


@pytest.mark.parametrize("layout", ["vec3", "soa"])
def test_identity_atom_map_free_atoms_first(layout):
    """Free atoms in front of the bodies in the caller's arrays: plugin order = caller order, no atomLoc table at all - the free
    atoms that ride along in part2LargeKernel take their slots from their list index."""
    rng = np.random.Generator(np.random.Philox(key=331))
    sysd = _system(rng.integers(9, 61, size=400), 600, seed=332)
    order = np.concatenate([np.flatnonzero(sysd["bodyIndices"] <= 0), np.flatnonzero(sysd["bodyIndices"] > 0)])
    sysd = {k: np.ascontiguousarray(v[order]) for k, v in sysd.items()}
    _run_pair(sysd, 0, layout, False, fused=True)
    _run_pair(sysd, 3, layout, False, fused=False)


@pytest.mark.parametrize("shuffle", [False, True])
def test_many_tiny_bodies_between_big_ones(shuffle):
    """Mean size > 8 but long runs of 3-atom bodies: up to 85 bodies in one 256-atom tile."""
    sizes = ([3] * 170 + [200] * 12 + [4, 5, 3, 60, 3, 3, 97]) * 3
    sysd = _system(sizes, 40, seed=303)
    a = _run_pair(sysd, 0, "vec3", shuffle, fused=True)
    b = _run_pair(sysd, 0, "vec3", shuffle, fused=False)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])      # same kernels either way: bit-identical


@pytest.mark.parametrize("mode", [0, 2])
def test_bodies_at_and_over_the_tile_capacity(mode):
    """255/256/257-atom bodies (the last is alone in its tile and reduced by the whole CTA), a 513- and a 3000-atom body."""
    sizes = [255, 256, 257, 9, 513, 31, 3000, 12, 256, 10]
    sysd = _system(sizes, 5, seed=304)
    _run_pair(sysd, mode, "vec3", True, fused=True, dt=0.0005)
    _run_pair(sysd, mode, "soa", False, fused=False, dt=0.0005)


@pytest.mark.parametrize("bulk", [True, False])
@pytest.mark.parametrize("layout,n_free", [("vec3", 400), ("soa", 400), ("soa", 401)])
def test_bodies_as_runs_of_slots_whole_molecules_permuted(layout, n_free, bulk, monkeypatch):
    """Whole bodies permuted (every body stays a run of consecutive slots, in any order, starting at even and odd slots):
    part2LargeRunsKernel - TMA bulk copies per body, serial sums - and, with RBK_NO_BULK_PART2=1, the per-atom request
    kernel on the same arrays.  SoA planes with an even stride take the three-copies-per-body path, with an odd stride
    (misaligned planes) the per-atom kernel whatever the switch says."""
    monkeypatch.setenv("RBK_NO_BULK_PART2", "0" if bulk else "1")
    rng = np.random.Generator(np.random.Philox(key=311))
    sysd = _system(rng.integers(3, 61, size=700), n_free, seed=312)
    if (len(sysd["masses"]) % 2 == 1) != (n_free == 401):
        sysd = _system(np.concatenate([rng.integers(3, 61, size=700), [4]]), n_free, seed=312)
    for mode, fused in ((0, True), (3, False)):
        _run_pair(sysd, mode, layout, "molecules", fused=fused, tol=5e-10)


@pytest.mark.parametrize("last", [3, 4, 60])
def test_last_run_ends_with_the_array(last):
    """No free atoms, identity atom map, the last body's forces end exactly where the Vec3 array ends: when that is at an odd
    8-byte word, the bulk copy stops 16 bytes early and the last word comes by cp.async - nothing is read past the array
    (compute-sanitizer run: tools/gpu_sanitize.sh)."""
    rng = np.random.Generator(np.random.Philox(key=313))
    for extra in (0, 1):
        sizes = np.concatenate([rng.integers(9, 61, size=300), [3 + extra, last]])
        sysd = _system(sizes, 0, seed=314)
        _run_pair(sysd, 0, "vec3", False, fused=True, tol=5e-10)


def test_sums_follow_the_atom_order():
    """The per-body force / torque sums add the atoms in the reference's order (RigidBody::forceAndTorque,
    openmmapi/src/RigidBody.cpp:174-183) up to the association into chunks of consecutive atoms: the body force agrees with
    the oracle's to a few ulps of the largest term, whatever the layout the forces came in."""
    rng = np.random.Generator(np.random.Philox(key=315))
    sysd = _system(rng.integers(9, 61, size=400), 100, seed=316)
    o = CpuStepper("oracle", sysd["bodyIndices"], sysd["masses"], 0)
    common.init_like_reference(o, sysd)
    o.step(0.001, 1)
    ob = o.bodies()
    ref = None
    for layout, shuffle in (("vec3", False), ("soa", True), ("vec3", "molecules")):
        s = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0, layout=layout, shuffle=shuffle)
        common.init_like_reference(s, sysd)
        s.step(0.001, 1)
        b = s.bodies()
        assert common.rel_inf(b["force"], ob["force"]) <= 1e-14
        assert common.rel_inf(b["torque"], ob["torque"]) <= 2e-12
        if ref is None:
            ref = b
        else:
            assert np.array_equal(b["force"], ref["force"]) and np.array_equal(b["torque"], ref["torque"])
        s.close()


@pytest.mark.parametrize("layout,shuffle", [("vec3", False), ("vec3", "molecules"), ("soa", True)])
def test_free_atoms_ride_along_in_the_body_kernel(layout, shuffle, monkeypatch):
    """Large-body systems with few enough free atoms per tile take them along in part2LargeKernel (Part 2, and Part 2 + Part 1
    in fused steps) instead of launching freeAtomsKernel: same arithmetic, bit-identical arrays, both against the oracle."""
    rng = np.random.Generator(np.random.Philox(key=321))
    sysd = common.synth.mixed_system(900, 2300, seed=322)
    outs = []
    for ride in ("0", "1"):
        monkeypatch.setenv("RBK_NO_FREE_RIDE", ride)
        for fused in (True, False):
            outs.append(_run_pair(sysd, 0, layout, shuffle, fused=fused, steps=4))
    for o in outs[1:]:
        assert np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1])


def test_bit_reproducible_large_bodies():
    rng = np.random.Generator(np.random.Philox(key=305))
    sysd = _system(rng.integers(3, 61, size=800), 300, seed=306)
    a = _run_pair(sysd, 0, "vec3", True, fused=True)
    b = _run_pair(sysd, 0, "vec3", True, fused=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def _raw_system(seed, n_bodies=600, n_free=500):
    """DeviceRigidBodySystem + device arrays of a large-body system with interleaved free atoms (non-identity atom map)."""
    import torch
    from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem
    sysd = common.synth.mixed_system(n_bodies, n_free, seed=seed)
    dev = torch.device("cuda:0")
    system = DeviceRigidBodySystem(sysd["bodyIndices"], sysd["masses"], 0)
    system.update(sysd["R"], sysd["V"], sysd["F"], True, True)
    system.upload()
    pos, vel = (torch.from_numpy(sysd[k].copy()).to(dev) for k in ("R", "V"))
    forces = (torch.from_numpy(sysd["F"].copy()).to(dev), torch.from_numpy(-sysd["F"]).to(dev))
    return system, pos, vel, forces


def test_side_stream_keeps_stream_order_and_can_be_captured():
    """The free atoms of a large-body step run on a second stream owned by the handle, forked from / joined to the
    caller's stream with events (include/rbk.h, rbk_part2_part1).  (1) On a non-default stream, work queued right
    behind the call must see the call's complete result.  (2) The fork/join pattern must be legal under stream
    capture: the same steps replayed from a CUDA graph give bit-identical arrays."""
    import torch
    dt, steps = 0.001, 6

    def plain(stream):
        system, pos, vel, F = _raw_system(411)
        snap = []
        with torch.cuda.stream(stream):
            system.part1(dt, pos, vel, F[0])
            for i in range(steps):
                system.part2_part1(dt, pos, vel, F[(i + 1) & 1])
                snap.append((pos.clone(), vel.clone()))          # queued on the same stream, no host sync in between
            system.part2(dt, pos, vel, F[(steps + 1) & 1])
        stream.synchronize()
        out = [(p.cpu().numpy(), v.cpu().numpy()) for p, v in snap] + [(pos.cpu().numpy(), vel.cpu().numpy())]
        system.close()
        return out

    ref = plain(torch.cuda.current_stream())
    other = plain(torch.cuda.Stream())
    for (p0, v0), (p1, v1) in zip(ref, other):
        assert np.array_equal(p0, p1) and np.array_equal(v0, v1)

    # CUDA graph: two plain fused steps first (they create the handle's side stream and set kernel attributes - not
    # things to do while capturing), then capture two fused steps and replay them twice: six fused steps in all
    system, pos, vel, F = _raw_system(411)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        system.part1(dt, pos, vel, F[0])
        system.part2_part1(dt, pos, vel, F[1])
        system.part2_part1(dt, pos, vel, F[0])
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        system.part2_part1(dt, pos, vel, F[1])
        system.part2_part1(dt, pos, vel, F[0])
    for _ in range((steps - 2) // 2):
        g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(pos.cpu().numpy(), ref[steps - 1][0]) and np.array_equal(vel.cpu().numpy(), ref[steps - 1][1])
    system.close()
