"""CPU tests pinning the oracle (oracle/rb_oracle.c):
  1. against the committed golden fixtures generated from the TRUE reference (tests/golden/make_golden.py);
  2. against the true reference itself when oracle/_ref/librb_ref.so is present (bit-for-bit);
  3. its elliptic functions against independent mpmath known answers.
"""
import glob
import os

import numpy as np
import pytest

import common
from common import GOLDEN_DIR, rel_inf
from oracle import checkers
from oracle.checkers import CpuStepper

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*_mode*.npz"))
               if not os.path.basename(p).startswith(("drift_", "gravity_")))


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def sysd_of(g):
    return {k: g[k] for k in ("bodyIndices", "masses", "R", "V", "F", "charges")}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(oracle_lib, name):
    g = load(name)
    s = CpuStepper("oracle", g["bodyIndices"], g["masses"], int(g["mode"]))
    common.init_like_reference(s, sysd_of(g), tether=bool(g["tether"]))
    c = s.counts()
    # integer work: bit exact
    assert [c[k] for k in ("numBodies", "numFree", "numActualAtoms", "numBodyAtoms", "numDOF")] == g["counts"].tolist()
    assert np.array_equal(s.body_index(), g["cleanIndex"])
    assert np.array_equal(s.atom_index(), g["atomIndex"])
    b = s.bodies()
    for k in ("N", "dof", "loc"):
        assert np.array_equal(b[k], g["b0_" + k])
    # floating point: the restatement keeps the operation order, so demand 1e-13 relative
    tol = 1e-13
    for k in ("mass", "I", "invI"):
        assert rel_inf(b[k], g["b0_" + k]) <= tol
    assert rel_inf(s.body_fixed(), g["b0_d"]) <= tol
    done = 0
    for cp in [0] + g["checkpoints"].tolist():
        s.step(float(g["dt"]), cp - done)
        done = cp
        R, V, _ = s.get_state()
        assert rel_inf(R, g[f"s{cp}_R"]) <= tol, (name, cp)
        assert rel_inf(V, g[f"s{cp}_V"]) <= tol, (name, cp)
        assert rel_inf(s.kinetic(), g[f"s{cp}_KE"]) <= tol
        if f"s{cp}_q" in g:
            bb = s.bodies()
            for k in ("rcm", "pcm", "q", "pi", "force", "torque"):
                assert rel_inf(bb[k], g[f"s{cp}_{k}"]) <= tol, (name, cp, k)


@pytest.mark.skipif(not checkers.available("reference"), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("mode", [0, 1, 10])
def test_oracle_bitwise_vs_true_reference(oracle_lib, mode):
    sysd = common.synth.mixed_system(60, 80, seed=31 + mode, max_atoms=25)
    out = {}
    for kind in ("reference", "oracle"):
        s = CpuStepper(kind, sysd["bodyIndices"], sysd["masses"], mode)
        common.init_like_reference(s, sysd, tether=True)
        s.step(0.002, 20)
        out[kind] = (s.get_state()[:2], s.kinetic(), s.bodies(), s.atom_index(), s.counts())
    (Ra, Va), ka, ba, ia, ca = out["reference"]
    (Rb, Vb), kb, bb, ib, cb = out["oracle"]
    assert ca == cb and np.array_equal(ia, ib)
    assert np.array_equal(Ra, Rb) and np.array_equal(Va, Vb) and np.array_equal(ka, kb)
    for k in ba:
        assert np.array_equal(ba[k], bb[k]), k


def test_reference_error_behaviour(oracle_lib):
    # RigidBodySystem.cpp:107-113: constraints touching body atoms are rejected
    with pytest.raises(RuntimeError, match="Constraints involving rigid-body atoms"):
        CpuStepper("oracle", [1, 1, 1, 0], [1.0] * 4, 0, constraints=[[0, 3]])
    # constraints between free atoms are fine and lower the DOF count (RigidBodySystem.cpp:130)
    s = CpuStepper("oracle", [1, 1, 1, 0, 0], [1.0] * 5, 0, constraints=[[3, 4]])
    R = np.array([[0, 0, 0], [0.1, 0, 0], [0, 0.1, 0], [1, 1, 1], [1.1, 1, 1.0]])
    s.set_state(R, np.zeros((5, 3)), np.zeros((5, 3)))
    s.update(True, True)
    assert s.counts()["numDOF"] == 2 - 1 + 6


def test_virtual_sites_and_massless_are_skipped(oracle_lib):
    # RigidBodySystem.cpp:63-81: virtual sites are not actual atoms; massless atoms are neither free nor counted in N
    body = [1, 1, 1, 0, 0, 0]
    masses = [16.0, 1.0, 1.0, 0.0, 12.0, 12.0]
    virt = [0, 0, 0, 1, 0, 0]
    s = CpuStepper("oracle", body, masses, 1, isVirtual=virt)
    c = s.counts()
    assert (c["numBodies"], c["numFree"], c["numActualAtoms"]) == (1, 2, 5)
    assert s.atom_index()[:2].tolist() == [4, 5]


def test_special_functions_vs_mpmath(oracle_lib):
    import ctypes as C
    lib = checkers._lib("oracle")
    k = np.load(os.path.join(GOLDEN_DIR, "special_functions_mpmath.npz"))
    sn, cn, dn = C.c_double(), C.c_double(), C.c_double()
    for u, m, esn, ecn, edn in k["jacobi"]:
        lib.orc_jacobi(u, m, C.byref(sn), C.byref(cn), C.byref(dn))
        assert abs(sn.value - esn) < 5e-13 and abs(cn.value - ecn) < 5e-13 and abs(dn.value - edn) < 5e-13, (u, m)
    for x, y, z, e in k["rf"]:
        assert abs(lib.orc_carlson_rf(x, y, z) / e - 1) < 1e-13
    for x, y, z, p, e in k["rj"]:
        assert abs(lib.orc_carlson_rj(x, y, z, p) / e - 1) < 1e-12
    for x, y, e in k["rc"]:
        assert abs(lib.orc_carlson_rc(x, y) / e - 1) < 1e-13


def test_exact_vs_nosquish_converge(oracle_lib):
    """Mode 0 (exact) and NO-SQUISH with many sub-steps must agree: an internal consistency check of
    the elliptic-function path that does not involve the reference at all."""
    import ctypes as C
    lib = checkers._lib("oracle")
    rng = np.random.Generator(np.random.Philox(key=3))
    for _ in range(50):
        I = np.sort(rng.uniform(0.5, 3.0, 3))[::-1].copy()
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        Lb = rng.standard_normal(3) * 5
        pi = 2 * np.array([-q[1]*Lb[0] - q[2]*Lb[1] - q[3]*Lb[2], q[0]*Lb[0] - q[3]*Lb[1] + q[2]*Lb[2],
                           q[3]*Lb[0] + q[0]*Lb[1] - q[1]*Lb[2], -q[2]*Lb[0] + q[1]*Lb[1] + q[0]*Lb[2]])
        q1, p1, q2, p2 = q.copy(), pi.copy(), q.copy(), pi.copy()
        invI = 1.0 / I
        lib.orc_exact_rotation(0.05, checkers._d(I), checkers._d(q1), checkers._d(p1))
        lib.orc_nosquish_rotation(0.05, 2000, 6, checkers._d(invI), checkers._d(q2), checkers._d(p2))
        assert np.max(np.abs(q1 - q2)) < 1e-7 and np.max(np.abs(p1 - p2)) < 1e-6 * np.max(np.abs(p2))


def test_refined_energy_restatement_matches_the_reference_cuda_kernels():
    """The reference computes refined kinetic energies / the potential-energy refinement only in its CUDA kernels
    (rigidbodyintegrator.cu:433-469).  Those kernels, compiled in place and run on a B200 (baseline/ref_cuda,
    tests/golden/make_golden_refcuda.py), produced tests/golden/refcuda_refined_*.npz: inputs, trajectory end point and
    the three energies.  oracle/rb_oracle.c's restatement must reproduce them - this pins the oracle's last section.
    (double precision + exact rotation: OpenMM defines USE_DOUBLE_PRECISION instead of USE_MIXED_PRECISION there and the
    reference's elliptic.cu then runs with its single-precision tolerances - its own rotation is only ~1e-7 accurate.)"""
    import glob
    paths = sorted(glob.glob(os.path.join(common.GOLDEN_DIR, "refcuda_refined_*.npz")))
    assert len(paths) >= 6
    for path in paths:
        g = dict(np.load(path))
        sysd = {k: g[k] for k in ("bodyIndices", "masses", "R", "V", "F", "charges")}
        sysd["F"] = np.round(sysd["F"] * 4294967296.0) / 4294967296.0            # OpenMM's fixed-point forces
        mode, steps, dt = int(g["mode"]), int(g["steps"]), float(g["dt"])
        o = CpuStepper("oracle", sysd["bodyIndices"], sysd["masses"], mode)
        common.init_like_reference(o, sysd)
        o.set_refined(True)
        o.step(dt, steps)
        ours = np.concatenate([o.refined_kinetic(dt), [o.potential_refinement(dt)]])
        R, V, _ = o.get_state()
        loose = str(g["precision"]) == "double" and mode == 0
        tol = np.array([1e-9, 5e-6, 1e-6]) if loose else 1e-11
        assert np.all(np.abs(ours - g["reference"]) <= tol * np.abs(g["reference"])), (path, ours, g["reference"])
        assert common.rel_inf(R, g["R_end"]) <= (2e-6 if loose else 1e-12), path
        assert common.rel_inf(V, g["V_end"]) <= (2e-5 if loose else 1e-10), path


def test_refined_energy_restatement_is_a_shadow_energy():
    """oracle/rb_oracle.c restates the reference's CUDA-only refined-energy diagnostics (pinned above against the
    reference's own kernels).  The physics behind them: for rigid bodies KE_refined + U + dU is conserved an order of
    magnitude better than KE + U, and better still at a smaller step (that is the purpose of the diagnostic); for free
    atoms the reference's factors (-1, 5, 2) give 4/3 of the kinetic energy in uniform motion (documented quirk)."""
    from oracle.checkers import CpuStepper
    from openmm_rigidbody_plugin_b200 import synth
    sysd = synth.water_box(64, seed=3)
    n = len(sysd["masses"])
    ch = np.tile([-0.834, 0.417, 0.417], n // 3)
    ratio = {}
    for mode in (0, 4):
        for dt in (0.001, 0.002):
            o = CpuStepper("oracle", sysd["bodyIndices"], sysd["masses"], mode)
            o.set_tether(5000.0, (300.0, -500.0, 800.0), ch, sysd["R"])
            o.set_state(sysd["R"], sysd["V"], sysd["F"])
            o.compute_forces()
            o.update(True, True)
            o.set_refined(True)
            plain, refined = [], []
            for _ in range(200):
                o.part1(dt)
                U = o.compute_forces()
                o.part2(dt)
                plain.append(o.kinetic().sum() + U)
                refined.append(o.refined_kinetic(dt).sum() + U + o.potential_refinement(dt))
            ratio[(mode, dt)] = np.std(refined) / np.std(plain)
            assert ratio[(mode, dt)] < 0.2, ratio
        assert ratio[(mode, 0.001)] < ratio[(mode, 0.002)]
    rng = np.random.default_rng(0)
    nf = 50
    masses, V = rng.uniform(1, 16, nf), rng.normal(size=(nf, 3))
    o = CpuStepper("oracle", np.zeros(nf, np.int32), masses, 0)
    o.set_state(rng.normal(size=(nf, 3)), V, np.zeros((nf, 3)))
    o.update(True, True)
    o.set_refined(True)
    o.step(0.002, 1)
    ke = 0.5 * np.sum(masses[:, None] * V * V)
    assert abs(o.refined_kinetic(0.002)[0] - 4.0 / 3.0 * ke) < 1e-12 * ke
    assert o.potential_refinement(0.002) == 0.0
