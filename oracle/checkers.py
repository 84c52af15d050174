"""ctypes front-end for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

  kind="oracle"    -> oracle/_build/librb_oracle.so  (this repo's C restatement, rb_oracle.c)
  kind="reference" -> oracle/_ref/librb_ref.so       (the unmodified reference sources + shim)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing in the product package openmm_rigidbody_plugin_b200 does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "oracle": os.path.join(_HERE, "_build", "librb_oracle.so"),
    "reference": os.path.join(_HERE, "_ref", "librb_ref.so"),
}
_PREFIX = {"oracle": "orc_", "reference": "ref_"}
_libs: dict = {}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(kind: str = "oracle") -> None:
    """(Re)build a checker with oracle/Makefile.  'reference' needs /root/reference."""
    target = "oracle" if kind == "oracle" else "ref"
    subprocess.run(["make", "-s", "-C", _HERE, target], check=True)


def available(kind: str) -> bool:
    return os.path.exists(_PATHS[kind])


def _lib(kind: str):
    if kind in _libs:
        return _libs[kind]
    if kind == "oracle" and not available(kind):
        build("oracle")
    if not available(kind):
        raise FileNotFoundError(f"CPU checker '{kind}' not built: {_PATHS[kind]}")
    lib = C.CDLL(_PATHS[kind])
    p = _PREFIX[kind]
    f = getattr(lib, p + "create")
    f.restype = C.c_void_p
    f.argtypes = [C.c_int, _ip, _dp, C.c_char_p, C.c_int, _ip, C.c_int]
    getattr(lib, p + "last_error").restype = C.c_char_p
    getattr(lib, p + "compute_forces").restype = C.c_double
    getattr(lib, p + "compute_forces").argtypes = [C.c_void_p]
    for name, args in [
        ("destroy", [C.c_void_p]),
        ("counts", [C.c_void_p, _ip]),
        ("body_index", [C.c_void_p, _ip]),
        ("atom_index", [C.c_void_p, _ip]),
        ("set_state", [C.c_void_p, _dp, _dp, _dp]),
        ("get_state", [C.c_void_p, _dp, _dp, _dp]),
        ("set_tether", [C.c_void_p, C.c_double, _dp, _dp, _dp]),
        ("update", [C.c_void_p, C.c_int, C.c_int]),
        ("part1", [C.c_void_p, C.c_double]),
        ("part2", [C.c_void_p, C.c_double]),
        ("step", [C.c_void_p, C.c_double, C.c_int]),
        ("set_alternate", [C.c_void_p, C.c_int]),
        ("kinetic", [C.c_void_p, _dp]),
        ("get_bodies", [C.c_void_p, _ip, _ip, _ip] + [_dp] * 10),
        ("get_body_fixed", [C.c_void_p, _dp]),
    ]:
        fn = getattr(lib, p + name)
        fn.restype = None
        fn.argtypes = args
    if kind == "oracle":
        lib.orc_jacobi.argtypes = [C.c_double, C.c_double, _dp, _dp, _dp]
        lib.orc_jacobi.restype = None
        for nm, n in (("orc_carlson_rc", 2), ("orc_carlson_rf", 3), ("orc_carlson_rj", 4)):
            getattr(lib, nm).argtypes = [C.c_double] * n
            getattr(lib, nm).restype = C.c_double
        lib.orc_exact_rotation.argtypes = [C.c_double, _dp, _dp, _dp]
        lib.orc_exact_rotation.restype = None
        lib.orc_nosquish_rotation.argtypes = [C.c_double, C.c_int, C.c_int, _dp, _dp, _dp]
        lib.orc_nosquish_rotation.restype = None
        lib.orc_set_refined.argtypes = [C.c_void_p, C.c_int]
        lib.orc_set_refined.restype = None
        lib.orc_refined_kinetic.argtypes = [C.c_void_p, C.c_double, _dp]
        lib.orc_refined_kinetic.restype = None
        lib.orc_potential_refinement.argtypes = [C.c_void_p, C.c_double]
        lib.orc_potential_refinement.restype = C.c_double
    _libs[kind] = lib
    return lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


class CpuStepper:
    """One rigid-body system stepped on the CPU by the oracle or by the true reference."""

    def __init__(self, kind, bodyIndices, masses, mode=0, isVirtual=None, constraints=None):
        self.kind = kind
        self.lib = _lib(kind)
        self.p = _PREFIX[kind]
        bi = np.ascontiguousarray(bodyIndices, dtype=np.int32)
        ms = np.ascontiguousarray(masses, dtype=np.float64)
        self.n = int(bi.shape[0])
        iv = None if isVirtual is None else np.ascontiguousarray(isVirtual, dtype=np.uint8).tobytes()
        cons = np.zeros((0, 2), np.int32) if constraints is None else np.ascontiguousarray(constraints, dtype=np.int32)
        self.h = getattr(self.lib, self.p + "create")(self.n, _i(bi), _d(ms), iv, int(cons.shape[0]), _i(cons), int(mode))
        if not self.h:
            raise RuntimeError(getattr(self.lib, self.p + "last_error")().decode())

    def _f(self, name):
        return getattr(self.lib, self.p + name)

    def close(self):
        if getattr(self, "h", None):
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counts(self):
        out = np.zeros(6, np.int32)
        self._f("counts")(self.h, _i(out))
        return dict(zip(["numBodies", "numFree", "numActualAtoms", "numBodyAtoms", "numDOF", "numAtoms"], out.tolist()))

    def body_index(self):
        out = np.zeros(self.n, np.int32)
        self._f("body_index")(self.h, _i(out))
        return out

    def atom_index(self):
        out = np.zeros(self.counts()["numActualAtoms"], np.int32)
        self._f("atom_index")(self.h, _i(out))
        return out

    def set_state(self, R=None, V=None, F=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (R, V, F)]
        self._f("set_state")(self.h, *[_d(a) for a in arrs])

    def get_state(self):
        R, V, F = (np.zeros((self.n, 3)) for _ in range(3))
        self._f("get_state")(self.h, _d(R), _d(V), _d(F))
        return R, V, F

    # refined ("shadow") energies: oracle only - the reference has them on its CUDA platform alone
    def set_refined(self, flag=True):
        self.lib.orc_set_refined(self.h, int(bool(flag)))

    def refined_kinetic(self, dt):
        out = np.zeros(2)
        self.lib.orc_refined_kinetic(self.h, float(dt), _d(out))
        return out

    def potential_refinement(self, dt):
        return float(self.lib.orc_potential_refinement(self.h, float(dt)))

    def set_tether(self, k, E, charges, x0):
        E = np.ascontiguousarray(E, dtype=np.float64)
        ch = np.ascontiguousarray(charges, dtype=np.float64)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self._f("set_tether")(self.h, float(k), _d(E), _d(ch), _d(x0))

    def compute_forces(self):
        return float(self._f("compute_forces")(self.h))

    def update(self, geometry=True, velocities=True):
        self._f("update")(self.h, int(geometry), int(velocities))

    def part1(self, dt):
        self._f("part1")(self.h, float(dt))

    def part2(self, dt):
        self._f("part2")(self.h, float(dt))

    def step(self, dt, steps=1):
        self._f("step")(self.h, float(dt), int(steps))

    def set_alternate(self, flag=True):
        self._f("set_alternate")(self.h, int(flag))

    def kinetic(self):
        out = np.zeros(2)
        self._f("kinetic")(self.h, _d(out))
        return out

    def bodies(self):
        nb = self.counts()["numBodies"]
        o = {
            "N": np.zeros(nb, np.int32), "dof": np.zeros(nb, np.int32), "loc": np.zeros(nb, np.int32),
            "mass": np.zeros(nb), "I": np.zeros((nb, 3)), "invI": np.zeros((nb, 3)), "rcm": np.zeros((nb, 3)),
            "pcm": np.zeros((nb, 3)), "q": np.zeros((nb, 4)), "pi": np.zeros((nb, 4)), "force": np.zeros((nb, 3)),
            "torque": np.zeros((nb, 4)), "twoK": np.zeros((nb, 2)),
        }
        self._f("get_bodies")(self.h, _i(o["N"]), _i(o["dof"]), _i(o["loc"]),
                              *[_d(o[k]) for k in ("mass", "I", "invI", "rcm", "pcm", "q", "pi", "force", "torque", "twoK")])
        return o

    def body_fixed(self):
        d = np.zeros((self.counts()["numBodyAtoms"], 3))
        self._f("get_body_fixed")(self.h, _d(d))
        return d
