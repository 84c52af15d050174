"""CPU tests of the PRODUCT's device arithmetic (csrc/rbk_math.cuh, rbk_step.cuh) compiled for the host
(lib/librbk_hostmath.so), checked against the oracle.  The kernels' orchestration (tiles, shuffles,
shared memory) can only be tested on a GPU (tests/test_gpu_parity.py); the math can be tested here."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import common
from oracle import checkers

LIB = os.path.join(common.ROOT, "openmm_rigidbody_plugin_b200", "lib", "librbk_hostmath.so")
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def hm(oracle_lib):
    subprocess.run(["make", "-s", "-C", os.path.join(common.ROOT, "openmm_rigidbody_plugin_b200", "csrc"),
                    "../lib/librbk_hostmath.so"], check=True)
    lib = C.CDLL(LIB)
    lib.rbkh_jacobi.argtypes = [C.c_double, C.c_double, _dp, _dp, _dp]
    for nm, n in (("rbkh_rc", 2), ("rbkh_rf", 3), ("rbkh_rj", 4)):
        getattr(lib, nm).argtypes = [C.c_double] * n
        getattr(lib, nm).restype = C.c_double
    lib.rbkh_exact_rotation.argtypes = [C.c_double, _dp, _dp, _dp, C.c_int]
    lib.rbkh_nosquish.argtypes = [C.c_double, C.c_int, _dp, _dp, _dp]
    lib.rbkh_exact_series.argtypes = [C.c_int, C.c_double, _dp, _dp, _dp]
    lib.rbkh_exact_series.restype = C.c_int
    lib.rbkh_series_excess.argtypes = [C.c_double, _dp, _dp, _dp]
    lib.rbkh_series_excess.restype = C.c_double
    return lib


def _d(a):
    return a.ctypes.data_as(_dp)


def random_rotor(rng, scale=5.0, water=False):
    if water:
        I = np.array([0.0176, 0.0115, 0.0061]) * rng.uniform(0.9, 1.1)     # amu nm^2, TIP3P-like
        Lb = rng.standard_normal(3) * np.sqrt(2.494 * I)
    else:
        I = np.sort(rng.uniform(0.5, 3.0, 3))[::-1].copy()
        Lb = rng.standard_normal(3) * scale
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    pi = 2 * np.array([-q[1]*Lb[0] - q[2]*Lb[1] - q[3]*Lb[2], q[0]*Lb[0] - q[3]*Lb[1] + q[2]*Lb[2],
                       q[3]*Lb[0] + q[0]*Lb[1] - q[1]*Lb[2], -q[2]*Lb[0] + q[1]*Lb[1] + q[0]*Lb[2]])
    return I, q, pi


def test_special_functions_match_mpmath(hm):
    k = np.load(os.path.join(common.GOLDEN_DIR, "special_functions_mpmath.npz"))
    sn, cn, dn = C.c_double(), C.c_double(), C.c_double()
    for u, m, esn, ecn, edn in k["jacobi"]:
        hm.rbkh_jacobi(u, m, C.byref(sn), C.byref(cn), C.byref(dn))
        assert abs(sn.value - esn) < 5e-13 and abs(cn.value - ecn) < 5e-13 and abs(dn.value - edn) < 5e-13
    for x, y, z, e in k["rf"]:
        assert abs(hm.rbkh_rf(x, y, z) / e - 1) < 1e-13
    for x, y, z, p, e in k["rj"]:
        assert abs(hm.rbkh_rj(x, y, z, p) / e - 1) < 1e-12
    for x, y, e in k["rc"]:
        assert abs(hm.rbkh_rc(x, y) / e - 1) < 1e-13


@pytest.mark.parametrize("elliptic_only", [1, 0])
@pytest.mark.parametrize("water,dt", [(False, 0.05), (False, 0.002), (True, 0.001), (True, 0.004), (False, 1.5)])
def test_exact_rotation_matches_oracle(hm, elliptic_only, water, dt):
    olib = checkers._lib("oracle")
    rng = np.random.Generator(np.random.Philox(key=17))
    worst = 0.0
    for _ in range(400):
        I, q, pi = random_rotor(rng, water=water)
        q1, p1, q2, p2 = q.copy(), pi.copy(), q.copy(), pi.copy()
        olib.orc_exact_rotation(dt, _d(I), _d(q1), _d(p1))
        hm.rbkh_exact_rotation(dt, _d(I), _d(q2), _d(p2), elliptic_only)
        worst = max(worst, np.max(np.abs(q1 - q2)), np.max(np.abs(p1 - p2)) / np.max(np.abs(p1)))
    assert worst < 2e-12, worst


def pi_from(q, Lb):
    return 2 * np.array([-q[1]*Lb[0] - q[2]*Lb[1] - q[3]*Lb[2], q[0]*Lb[0] - q[3]*Lb[1] + q[2]*Lb[2],
                         q[3]*Lb[0] + q[0]*Lb[1] - q[1]*Lb[2], -q[2]*Lb[0] + q[1]*Lb[1] + q[0]*Lb[2]])


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_series_reduction_axis_choice(hm, axis, sign):
    """Angular momentum almost along +-each principal axis: the series picks the reduction axis per body
    (rbk_math.cuh, exactRotationSeries), so every case converges at the production order in ONE step, agrees with the
    elliptic oracle (which itself loses digits for rotation close to a principal axis - measured up to 3e-9 next to
    +axis 1 and 3e-11 next to the others - hence the 1e-8 / 1e-9 bars) and with itself taken as two half steps
    (1e-14): the flow property, independent of any oracle."""
    olib = checkers._lib("oracle")
    rng = np.random.Generator(np.random.Philox(key=100 + 3 * axis + int(sign > 0)))
    dt = 0.002
    for _ in range(200):
        I = np.sort(rng.uniform(0.02, 3.0, 3))[::-1].copy()
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        Lb = 0.02 * rng.standard_normal(3)
        Lb[axis] = sign * rng.uniform(0.5, 3.0)
        Lb *= min(1.0, 0.05 * np.min(I) / (dt * np.max(np.abs(Lb))))            # keep |omega| dt moderate
        pi = pi_from(q, Lb)
        assert hm.rbkh_series_excess(dt, _d(I), _d(q), _d(pi)) == 0.0
        q1, p1, q2, p2, q3, p3 = q.copy(), pi.copy(), q.copy(), pi.copy(), q.copy(), pi.copy()
        assert hm.rbkh_exact_series(12, dt, _d(I), _d(q1), _d(p1)) == 1
        for _half in range(2):
            assert hm.rbkh_exact_series(12, dt / 2, _d(I), _d(q2), _d(p2)) == 1
        assert np.max(np.abs(q1 - q2)) < 1e-14 and np.max(np.abs(p1 - p2)) < 1e-14 * max(1.0, np.max(np.abs(p1)))
        olib.orc_exact_rotation(dt, _d(I), _d(q3), _d(p3))
        tol = 1e-8 if (axis == 0 and sign > 0) else 1e-9
        if np.all(np.isfinite(q3)):
            assert np.max(np.abs(q1 - q3)) < tol and np.max(np.abs(p1 - p3)) < tol * max(1.0, np.max(np.abs(p3)))


def test_nosquish_matches_oracle(hm):
    olib = checkers._lib("oracle")
    rng = np.random.Generator(np.random.Philox(key=18))
    for n in (1, 3, 10):
        for _ in range(100):
            I, q, pi = random_rotor(rng)
            invI = 1.0 / I
            q1, p1, q2, p2 = q.copy(), pi.copy(), q.copy(), pi.copy()
            olib.orc_nosquish_rotation(0.01, n, 6, _d(invI), _d(q1), _d(p1))
            hm.rbkh_nosquish(0.01, n, _d(invI), _d(q2), _d(p2))
            assert np.max(np.abs(q1 - q2)) < 1e-14 and np.max(np.abs(p1 - p2)) < 1e-13 * np.max(np.abs(p1))
