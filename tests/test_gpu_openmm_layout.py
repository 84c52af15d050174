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
