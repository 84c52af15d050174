"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol include/rbk.h
declares, and its host model (index mapping, DOF, body build) reproduces the true reference's golden
values bit for bit.  No compute entry point is called here (there is no GPU and no CPU fallback)."""
import glob
import os
import re

import numpy as np
import pytest

import common
from common import GOLDEN_DIR

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*_mode*.npz"))
               if not os.path.basename(p).startswith(("drift_", "gravity_")))


@pytest.fixture(scope="module")
def rbk():
    import __graft_entry__ as g
    g.build()
    import openmm_rigidbody_plugin_b200 as pkg
    return pkg


def test_library_exports_every_declared_symbol(rbk):
    from openmm_rigidbody_plugin_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(common.ROOT, "include", "rbk.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rbk_[a-z0-9_]+)\s*\(", header)) - {"rbk_force_fn", "rbk_positions_fn", "rbk_velocities_fn"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rbk_version() == 1


@pytest.mark.parametrize("name", CASES)
def test_host_model_matches_true_reference(rbk, name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    s = rbk.DeviceRigidBodySystem(g["bodyIndices"], g["masses"], int(g["mode"]))
    n = len(g["masses"])
    F0 = g["F"]
    if int(g["tether"]):
        F0 = g["charges"][:, None] * common.TETHER_E[None, :] + 0.0 * g["R"]      # dx = 0 at the initial positions
        F0 = (g["R"] - g["R"]) * (-common.TETHER_K) + common.TETHER_E[None, :] * g["charges"][:, None]
    s.update(g["R"], np.zeros((n, 3)), F0, True, True)
    s.update(V=g["V"], geometry=False, velocities=True)
    c = s.counts()
    assert [c[k] for k in ("numBodies", "numFree", "numActualAtoms", "numBodyAtoms", "numDOF")] == g["counts"].tolist()
    assert np.array_equal(s.body_index(), g["cleanIndex"])
    assert np.array_equal(s.atom_index(), g["atomIndex"])
    b = s.host_bodies()
    for k in ("N", "dof", "loc", "mass", "I", "invI"):
        assert np.array_equal(b[k], g["b0_" + k]), k
    assert np.array_equal(s.body_fixed(), g["b0_d"])
    if "s0_q" in g:
        for k in ("rcm", "pcm", "q", "pi", "force", "torque"):
            assert np.array_equal(b[k], g["s0_" + k]), k


def test_error_behaviour(rbk):
    from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem, RbkError
    with pytest.raises(RbkError, match="Rotation mode cannot be negative"):          # RigidBodyIntegrator.cpp:27-28
        DeviceRigidBodySystem([1, 1, 1], [1.0, 1.0, 1.0], -1)
    with pytest.raises(RbkError, match="Constraints involving rigid-body atoms are not allowed"):   # RigidBodySystem.cpp:111-112
        DeviceRigidBodySystem([1, 1, 1, 0], [1.0] * 4, 0, constraints=[[0, 3]])
    with pytest.raises(RbkError, match="virtual sites"):
        DeviceRigidBodySystem([1, 1, 1, 1], [1.0] * 4, 0, isVirtual=[0, 0, 0, 1])
    s = DeviceRigidBodySystem([1, 1, 1, 0, 0], [1.0] * 5, 0, constraints=[[3, 4]])
    R = np.array([[0, 0, 0], [0.1, 0, 0], [0, 0.1, 0], [1, 1, 1], [1.1, 1, 1.0]])
    s.update(R, np.zeros((5, 3)), np.zeros((5, 3)), True, True)
    assert s.counts()["numDOF"] == 2 - 1 + 6          # free atoms count ONE dof each (RigidBodySystem.cpp:130)
    with pytest.raises(RbkError, match="CUDA device|not uploaded"):
        import ctypes
        s.part1(0.001, ctypes.c_void_p(0).value or 0, 0, 0, layout=(0, 0))


def test_no_device_means_loud_failure(rbk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem, RbkError
    s = DeviceRigidBodySystem([1, 1, 1], [16.0, 1.0, 1.0], 0)
    s.update(np.array([[0, 0, 0], [0.1, 0, 0], [0, 0.1, 0.0]]), np.zeros((3, 3)), np.zeros((3, 3)), True, True)
    with pytest.raises(RbkError, match="no CPU fallback"):
        s.upload()


def test_velocities_set_twice_give_the_same_momentum():
    """setVelocities(V) twice must leave the same body momenta as once (the reference's buildDynamics adds into its
    previous pcm, RigidBody.cpp:126-130, which doubles p and quadruples 2Kt; the product SETS p = sum m v, like its
    GPU-side build)."""
    from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem
    sysd = common.synth.mixed_system(50, 30, seed=3, max_atoms=12)
    n = len(sysd["masses"])
    s = DeviceRigidBodySystem(sysd["bodyIndices"], sysd["masses"], 0)
    s.update(sysd["R"], np.zeros((n, 3)), sysd["F"], True, True)
    s.update(V=sysd["V"], geometry=False, velocities=True)
    once = s.host_bodies()
    s.update(V=sysd["V"], geometry=False, velocities=True)
    twice = s.host_bodies()
    for k in ("pcm", "pi", "twoK"):
        assert np.array_equal(once[k], twice[k]), k
    s.update(sysd["R"], sysd["V"], sysd["F"], True, True)
    third = s.host_bodies()
    assert np.array_equal(once["pcm"], third["pcm"]) and np.allclose(once["twoK"], third["twoK"], rtol=1e-13)
