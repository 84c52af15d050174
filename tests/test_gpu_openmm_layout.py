"""GPU tests of the OpenMM-CUDA boundary formats: posq real4 (+ posqCorrection in mixed
precision), velm mixed4 with the inverse mass in .w, int64 fixed-point force planes with stride paddedNumAtoms, atoms in
a reordered order with padding.  OpenMM is not installed, so the buffers are built here exactly as
platforms/cuda/src/kernels/rigidbodyintegrator.cu:30-64,270-279 of the reference reads them, and the result is compared
with the fp64 Vec3 path fed with the SAME (quantised) forces."""
import numpy as np
import pytest

import common
from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem
from openmm_rigidbody_plugin_b200._lib import RBK_OPENMM_DOUBLE, RBK_OPENMM_MIXED, RBK_OPENMM_SINGLE

pytestmark = pytest.mark.gpu


def build(sysd, mode):
    n = len(sysd["masses"])
    s = DeviceRigidBodySystem(sysd["bodyIndices"], sysd["masses"], mode)
    s.update(sysd["R"], np.zeros((n, 3)), sysd["F"], True, True)
    s.update(V=sysd["V"], geometry=False, velocities=True)
    s.upload()
    return s


@pytest.mark.parametrize("precision", [RBK_OPENMM_MIXED, RBK_OPENMM_DOUBLE, RBK_OPENMM_SINGLE])
@pytest.mark.parametrize("mode", [0, 2])
def test_openmm_formats_match_vec3_path(precision, mode):
    import torch
    dev = torch.device("cuda:0")
    sysd = common.synth.mixed_system(700, 900, seed=41, max_atoms=30)
    n = len(sysd["masses"])
    padded = ((n + 31) // 32) * 32 + 64
    rng = np.random.Generator(np.random.Philox(key=8))
    order = rng.permutation(n)                       # original atom i lives at slot order[i] (OpenMM reorders atoms)
    # forces quantised to OpenMM's fixed point, used by BOTH paths
    Fq = np.round(sysd["F"] * 4294967296.0).astype(np.int64)
    sysd = dict(sysd, F=Fq.astype(np.float64) / 4294967296.0)
    charges = sysd["charges"]
    invm = 1.0 / sysd["masses"]

    # ---- reference: fp64 Vec3 layout, same permutation
    a = build(sysd, mode)
    loc = order[a.atom_index()].astype(np.int32)
    a.set_atom_location(loc)

    def permuted(x):
        out = np.zeros((padded, 3))
        out[order] = x
        return torch.from_numpy(out).to(dev)
    R, V, F = permuted(sysd["R"]), permuted(sysd["V"]), permuted(sysd["F"])
    for _ in range(3):
        a.part1(0.001, R, V, F)
        a.part2(0.001, R, V, F)
    keA = a.kinetic(V)

    # ---- OpenMM formats
    b = build(sysd, mode)
    b.set_atom_location(loc)
    real = torch.float32 if precision != RBK_OPENMM_DOUBLE else torch.float64
    vreal = torch.float32 if precision == RBK_OPENMM_SINGLE else torch.float64
    pos0 = np.zeros((padded, 4)); pos0[order, :3] = sysd["R"]; pos0[order, 3] = charges
    vel0 = np.zeros((padded, 4)); vel0[order, :3] = sysd["V"]; vel0[order, 3] = invm
    posq = torch.from_numpy(pos0).to(dev).to(real).contiguous()
    corr = None
    if precision == RBK_OPENMM_MIXED:
        corr = (torch.from_numpy(pos0).to(dev) - posq.double()).float().contiguous()
        corr[:, 3] = 7.0                                         # must never be touched
    velm = torch.from_numpy(vel0).to(dev).to(vreal).contiguous()
    fplanes = np.zeros((3, padded), np.int64)
    fplanes[:, order] = Fq.T
    force = torch.from_numpy(fplanes).to(dev).contiguous()
    for _ in range(3):
        b.part1_openmm(0.001, posq, corr, velm, force, padded, precision)
        b.part2_openmm(0.001, posq, corr, velm, force, padded, precision)
    keB = b.kinetic_openmm(velm, precision)

    Rb = posq.double()[:, :3] + (corr.double()[:, :3] if corr is not None else 0.0)
    Vb = velm.double()[:, :3]
    tolR = {RBK_OPENMM_MIXED: 1e-12, RBK_OPENMM_DOUBLE: 1e-15, RBK_OPENMM_SINGLE: 2e-6}[precision]
    tolV = 2e-6 if precision == RBK_OPENMM_SINGLE else 1e-14
    scaleR, scaleV = float(R.abs().max()), float(V.abs().max())
    assert float((Rb - R).abs().max()) <= tolR * scaleR
    assert float((Vb - V).abs().max()) <= tolV * scaleV
    assert np.allclose(keA, keB, rtol=1e-5 if precision == RBK_OPENMM_SINGLE else 1e-14)
    # .w components (charge, inverse mass, correction.w) are preserved bit for bit; padding is untouched
    assert torch.equal(posq[:, 3], torch.from_numpy(pos0[:, 3]).to(dev).to(real))
    assert torch.equal(velm[:, 3], torch.from_numpy(vel0[:, 3]).to(dev).to(vreal))
    if corr is not None:
        assert bool((corr[:, 3] == 7.0).all())
    pad = np.setdiff1d(np.arange(padded), order)
    assert float(posq[pad].abs().max()) == 0.0 and float(velm[pad].abs().max()) == 0.0


class OpenMMArrays:
    """posq / posqCorrection / velm / force planes as OpenMM's CUDA platform holds them, atoms stored at slot order[i]."""

    def __init__(self, sysd, order, padded, precision, Fq):
        import torch
        self.torch, self.dev, self.precision, self.padded = torch, torch.device("cuda:0"), precision, padded
        n = len(sysd["masses"])
        pos0 = np.zeros((padded, 4)); pos0[order, :3] = sysd["R"]; pos0[order, 3] = sysd["charges"]
        vel0 = np.zeros((padded, 4)); vel0[order, :3] = sysd["V"]; vel0[order, 3] = 1.0 / sysd["masses"]
        p64 = torch.from_numpy(pos0).to(self.dev)
        if precision == RBK_OPENMM_MIXED:
            self.posq = p64.float().contiguous()
            self.corr = (p64 - self.posq.double()).float().contiguous()
        else:
            self.posq, self.corr = p64.contiguous(), None
        self.velm = torch.from_numpy(vel0).to(self.dev).contiguous()
        planes = np.zeros((3, padded), np.int64)
        planes[:, order] = Fq.T
        self.force = torch.from_numpy(planes).to(self.dev).contiguous()
        self.order = order.copy()

    def args(self):
        return self.posq, self.corr, self.velm, self.force, self.padded, self.precision

    def reorder(self, new_order):
        """what cu.reorderAtoms() does to posq / posqCorrection / velm (the forces are the listener's job)"""
        t = self.torch
        src = t.from_numpy(self.order).to(self.dev)
        dst = t.from_numpy(new_order).to(self.dev)
        for name in ("posq", "corr", "velm"):
            a = getattr(self, name)
            if a is None:
                continue
            b = t.zeros_like(a)
            b[dst] = a[src]
            setattr(self, name, b)
        self.order = new_order.copy()

    def host(self):
        R = self.posq.double()[:, :3]
        if self.corr is not None:
            R = R + self.corr.double()[:, :3]
        return R.cpu().numpy()[self.order], self.velm[:, :3].cpu().numpy()[self.order]


@pytest.mark.parametrize("precision", [RBK_OPENMM_MIXED, RBK_OPENMM_DOUBLE])
@pytest.mark.parametrize("case,mode", [("water", 0), ("water", 3), ("small_mixed", 0), ("medium_mixed", 0), ("large_mixed", 0),
                                       ("large_mixed_molecules", 0), ("large_mixed_molecules", 2)])
def test_openmm_fused_stepping_and_reorder_vs_oracle(case, mode, precision):
    """The CUDA-platform flow on the OpenMM formats against the CPU ORACLE (not against this repo's own fp64 path):
    part1, [rbk_part2_part1_openmm, atoms reordered every other step through rbk_reorder_openmm] x n, part2 - with the
    reordering at the only point a fused step leaves for it, after the one-pass call (state: Part 1 of the next step done)."""
    import torch
    from oracle.checkers import CpuStepper
    if case == "water":
        sysd = common.synth.water_box(4100, seed=51)
    elif case == "small_mixed":
        sysd = common.synth.mixed_system(2500, 3000, seed=52, max_atoms=7)
    elif case == "medium_mixed":
        sysd = common.synth.mixed_system(2500, 1000, seed=53, max_atoms=12)
    else:
        sysd = common.synth.mixed_system(500, 600, seed=54, max_atoms=60)
    n = len(sysd["masses"])
    Fq = np.round(sysd["F"] * 4294967296.0).astype(np.int64)
    sysd = dict(sysd, F=Fq.astype(np.float64) / 4294967296.0)
    steps, dt = 6, 0.001
    o = CpuStepper("oracle", sysd["bodyIndices"], sysd["masses"], mode)
    common.init_like_reference(o, sysd)
    o.step(dt, steps)
    Ro, Vo, _ = o.get_state()

    rng = np.random.Generator(np.random.Philox(key=55))
    padded = ((n + 31) // 32) * 32 + 32
    s = build(sysd, mode)
    index = s.atom_index()
    # water: whole molecules are permuted, as OpenMM's reorderAtoms does (the handle re-sorts its bodies each time);
    # large_mixed_molecules: the same for bodies of 3-60 atoms - every body stays a run of slots and large-body Part 2 moves
    # whole runs of the fixed-point force planes with bulk copies; the other cases: every atom on its own, the general gather path
    permutation = (lambda: common.molecule_permutation(sysd["bodyIndices"], int(rng.integers(1 << 30)))) if case in ("water", "large_mixed_molecules") else (lambda: rng.permutation(n))
    A = OpenMMArrays(sysd, permutation(), padded, precision, Fq)
    s.set_atom_location(A.order[index].astype(np.int32))
    s.part1_openmm(dt, *A.args())
    for k in range(steps - 1):
        s.part2_part1_openmm(dt, *A.args())
        if k % 2 == 1:
            new_order = permutation()
            expect = torch.zeros_like(A.force)
            expect[:, torch.from_numpy(new_order).to(A.dev)] = A.force[:, torch.from_numpy(A.order).to(A.dev)]
            A.reorder(new_order)
            s.reorder_openmm(new_order[index].astype(np.int32), A.force, padded)
            assert torch.equal(A.force, expect)               # the listener moved every owned atom's force to its new slot
    s.part2_openmm(dt, *A.args())
    Rg, Vg = A.host()
    eR, eV = common.rel_inf(Rg, Ro), common.rel_inf(Vg, Vo)
    assert eR <= 2e-10 and eV <= 2e-10, (case, mode, precision, eR, eV)
    ke = s.kinetic_openmm(A.velm, precision)
    assert common.rel_inf(ke, o.kinetic()) <= 2e-10
    b, ob = s.download_bodies(), o.bodies()
    for k in ("rcm", "pcm", "pi", "force", "torque"):
        assert common.rel_inf(b[k], ob[k]) <= 1e-8, k
    assert common.quat_rel(b["q"], ob["q"]) <= 2e-10
