"""ctypes binding of librbk.so (include/rbk.h).  Fails loudly when the library is missing or broken:
there is no Python/CPU fallback for the integrator step."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RBK_LIB_PATH: load another build of the same library (A/B experiments with different compile-time settings)
LIB_PATH = os.environ.get("RBK_LIB_PATH") or os.path.join(_HERE, "lib", "librbk.so")

RBK_LAYOUT_VEC3 = 0
RBK_LAYOUT_SOA = 1
RBK_OPENMM_SINGLE, RBK_OPENMM_MIXED, RBK_OPENMM_DOUBLE = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
FORCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)
HOOK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)     # rbk_positions_fn / rbk_velocities_fn

# every symbol include/rbk.h declares, with its signature
SIGNATURES = {
    "rbk_version": (C.c_int, []),
    "rbk_last_error": (C.c_char_p, []),
    "rbk_debug_series_order": (C.c_int, [C.c_void_p, _ip, C.c_void_p]),
    "rbk_debug_copy_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong)]),
    "rbk_debug_launches_per_call": (C.c_int, [C.c_void_p, _ip]),
    "rbk_create": (C.c_int, [C.c_int, _ip, _dp, C.c_char_p, C.c_int, _ip, C.c_int, C.POINTER(C.c_void_p)]),
    "rbk_destroy": (None, [C.c_void_p]),
    "rbk_get_counts": (C.c_int, [C.c_void_p, _ip]),
    "rbk_get_body_index": (C.c_int, [C.c_void_p, _ip]),
    "rbk_get_atom_index": (C.c_int, [C.c_void_p, _ip]),
    "rbk_update": (C.c_int, [C.c_void_p, _dp, _dp, _dp, C.c_int, C.c_int]),
    "rbk_get_host_bodies": (C.c_int, [C.c_void_p, _ip, _ip, _ip] + [_dp] * 10),
    "rbk_get_body_fixed": (C.c_int, [C.c_void_p, _dp]),
    "rbk_upload": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rbk_update_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "rbk_update_device_openmm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "rbk_set_atom_location": (C.c_int, [C.c_void_p, _ip, C.c_void_p]),
    "rbk_part1": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]),
    "rbk_part2": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]),
    "rbk_part2_part1": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]),
    "rbk_kinetic": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, _dp, C.c_void_p]),
    "rbk_part1_openmm": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rbk_part2_openmm": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rbk_part2_part1_openmm": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rbk_reorder_openmm": (C.c_int, [C.c_void_p, _ip, C.c_void_p, C.c_int, C.c_void_p]),
    "rbk_kinetic_openmm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _dp, C.c_void_p]),
    "rbk_kinetic_host": (C.c_int, [C.c_void_p, C.c_void_p, _dp, C.c_void_p]),
    "rbk_download_bodies": (C.c_int, [C.c_void_p] + [_dp] * 6 + [C.c_void_p]),
    "rbk_execute_host": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, FORCE_FN, C.c_void_p, C.c_void_p]),
    "rbk_set_refined_energies": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "rbk_refined_kinetic": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_longlong, _dp, C.c_void_p]),
    "rbk_potential_refinement": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_longlong, _dp, C.c_void_p]),
    "rbk_refined_kinetic_openmm": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_int, _dp, C.c_void_p]),
    "rbk_potential_refinement_openmm": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_int, _dp, C.c_void_p]),
    "rbk_free_dot_openmm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]),
    "rbk_refined_kinetic_host": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, _dp, C.c_void_p]),
    "rbk_potential_refinement_host": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, _dp, C.c_void_p]),
    "rbk_execute_host_hooks": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, FORCE_FN, HOOK_FN, HOOK_FN,
                                         C.c_void_p, C.c_void_p]),
    "rbk_free_delta_openmm": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "rbk_part1_delta_openmm": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p]),
}

_lib = None


class OpenMMException(RuntimeError):
    """Stand-in for OpenMM::OpenMMException, the exception type of the reference's API."""


class RbkError(OpenMMException):
    """Raised for any non-zero return code of the C ABI (the reference throws OpenMMException)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback for the rigid-body step.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RbkError(rc, load().rbk_last_error().decode())
