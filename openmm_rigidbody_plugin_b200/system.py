"""Thin object wrapper over the C ABI: one `DeviceRigidBodySystem` = one rbk_system handle.

Device arrays are passed as torch CUDA tensors (float64) or raw device pointers; torch is only the
allocator / stream provider here, the arithmetic is librbk's CUDA kernels."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import RBK_LAYOUT_SOA, RBK_LAYOUT_VEC3, RbkError, check  # noqa: F401

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _host(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _ptr(t):
    """Device (or pinned host) pointer of a torch tensor / int / None."""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    return C.c_void_p(t.data_ptr())


def _stream(stream):
    if stream is None:
        try:
            import torch
            if torch.cuda.is_available():
                return C.c_void_p(torch.cuda.current_stream().cuda_stream)
        except ImportError:
            pass
        return None
    return C.c_void_p(int(stream))


def _layout(t, layout):
    """(layout, stride) for a tensor: [N,3] contiguous -> VEC3; [3,N] contiguous -> SOA with stride N."""
    if layout is not None:
        return layout
    shape = tuple(t.shape)
    if len(shape) == 2 and shape[1] == 3:
        return (RBK_LAYOUT_VEC3, 0)
    if len(shape) == 2 and shape[0] == 3:
        return (RBK_LAYOUT_SOA, shape[1])
    raise ValueError("atom arrays must be [N,3] (Vec3) or [3,N] (SoA planes)")


class DeviceRigidBodySystem:
    def __init__(self, bodyIndices, masses, rotationMode=0, isVirtual=None, constraints=None):
        self.lib = _lib.load()
        bi = np.ascontiguousarray(bodyIndices, dtype=np.int32)
        ms = np.ascontiguousarray(masses, dtype=np.float64)
        if bi.ndim != 1 or bi.shape != ms.shape:
            raise ValueError("bodyIndices and masses must be 1-D and of equal length")
        self.numAtoms = int(bi.shape[0])
        iv = None if isVirtual is None else np.ascontiguousarray(isVirtual, dtype=np.uint8).tobytes()
        cons = np.zeros((0, 2), np.int32) if constraints is None else np.ascontiguousarray(constraints, dtype=np.int32).reshape(-1, 2)
        h = C.c_void_p()
        check(self.lib.rbk_create(self.numAtoms, _i(bi), _d(ms), iv, int(cons.shape[0]), _i(cons), int(rotationMode), C.byref(h)))
        self.h = h
        self.rotationMode = int(rotationMode)

    def close(self):
        if getattr(self, "h", None):
            self.lib.rbk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host model
    def counts(self):
        out = np.zeros(5, np.int32)
        check(self.lib.rbk_get_counts(self.h, _i(out)))
        return dict(zip(["numBodies", "numFree", "numActualAtoms", "numBodyAtoms", "numDOF"], out.tolist()))

    def body_index(self):
        out = np.zeros(self.numAtoms, np.int32)
        check(self.lib.rbk_get_body_index(self.h, _i(out)))
        return out

    def atom_index(self):
        out = np.zeros(self.counts()["numActualAtoms"], np.int32)
        check(self.lib.rbk_get_atom_index(self.h, _i(out)))
        return out

    def update(self, R=None, V=None, F=None, geometry=True, velocities=True):
        R, V, F = _host(R), _host(V), _host(F)
        check(self.lib.rbk_update(self.h, _d(R), _d(V), _d(F), int(geometry), int(velocities)))

    def host_bodies(self):
        nb = self.counts()["numBodies"]
        o = {
            "N": np.zeros(nb, np.int32), "dof": np.zeros(nb, np.int32), "loc": np.zeros(nb, np.int32),
            "mass": np.zeros(nb), "I": np.zeros((nb, 3)), "invI": np.zeros((nb, 3)), "rcm": np.zeros((nb, 3)),
            "pcm": np.zeros((nb, 3)), "q": np.zeros((nb, 4)), "pi": np.zeros((nb, 4)), "force": np.zeros((nb, 3)),
            "torque": np.zeros((nb, 4)), "twoK": np.zeros((nb, 2)),
        }
        check(self.lib.rbk_get_host_bodies(self.h, _i(o["N"]), _i(o["dof"]), _i(o["loc"]),
                                           *[_d(o[k]) for k in ("mass", "I", "invI", "rcm", "pcm", "q", "pi", "force", "torque", "twoK")]))
        return o

    def body_fixed(self):
        d = np.zeros((self.counts()["numBodyAtoms"], 3))
        check(self.lib.rbk_get_body_fixed(self.h, _d(d)))
        return d

    def series_order(self, stream=None):
        """Taylor order the mode-0 water kernels will use at their next launch (rbk_debug_series_order)."""
        out = np.zeros(1, np.int32)
        check(self.lib.rbk_debug_series_order(self.h, _i(out), _stream(stream)))
        return int(out[0])

    def launches_per_call(self):
        """Kernel launches of part1, part2, part2_part1 on this system (rbk_debug_launches_per_call)."""
        out = np.zeros(3, np.int32)
        check(self.lib.rbk_debug_launches_per_call(self.h, _i(out)))
        return tuple(int(x) for x in out)

    # ---- device
    def upload(self, stream=None):
        check(self.lib.rbk_upload(self.h, _stream(stream)))

    def update_device(self, pos=None, vel=None, force=None, geometry=True, velocities=True, layout=None, stream=None):
        """GPU-side body build from device arrays (rbk_update_device): no host rebuild, no upload."""
        ref = pos if pos is not None else vel
        lay, stride = _layout(ref, layout)
        check(self.lib.rbk_update_device(self.h, _ptr(pos), _ptr(vel), _ptr(force), lay, stride, int(geometry), int(velocities),
                                         _stream(stream)))

    def update_device_openmm(self, posq, posqCorrection, velm, force, paddedNumAtoms, precision, geometry=True,
                             velocities=True, stream=None):
        check(self.lib.rbk_update_device_openmm(self.h, _ptr(posq), _ptr(posqCorrection), _ptr(velm), _ptr(force),
                                                int(paddedNumAtoms), int(precision), int(geometry), int(velocities), _stream(stream)))

    def set_atom_location(self, location=None, stream=None):
        loc = None if location is None else np.ascontiguousarray(location, dtype=np.int32)
        check(self.lib.rbk_set_atom_location(self.h, _i(loc), _stream(stream)))

    def part1(self, dt, pos, vel, force, layout=None, stream=None):
        lay, stride = _layout(pos, layout)
        check(self.lib.rbk_part1(self.h, float(dt), _ptr(pos), _ptr(vel), _ptr(force), lay, stride, _stream(stream)))

    def part2(self, dt, pos, vel, force, layout=None, stream=None):
        lay, stride = _layout(pos, layout)
        check(self.lib.rbk_part2(self.h, float(dt), _ptr(pos), _ptr(vel), _ptr(force), lay, stride, _stream(stream)))

    def part2_part1(self, dt, pos, vel, force, layout=None, stream=None):
        """Part 2 of this step fused with Part 1 of the next one (rbk_part2_part1)."""
        lay, stride = _layout(pos, layout)
        check(self.lib.rbk_part2_part1(self.h, float(dt), _ptr(pos), _ptr(vel), _ptr(force), lay, stride, _stream(stream)))

    def kinetic(self, vel, layout=None, stream=None):
        lay, stride = _layout(vel, layout)
        out = np.zeros(2)
        check(self.lib.rbk_kinetic(self.h, _ptr(vel), lay, stride, _d(out), _stream(stream)))
        return out

    # ---- OpenMM-CUDA boundary formats (posq real4 [+ correction], velm mixed4, int64 fixed-point force planes)
    def part1_openmm(self, dt, posq, posqCorrection, velm, force, paddedNumAtoms, precision, stream=None):
        check(self.lib.rbk_part1_openmm(self.h, float(dt), _ptr(posq), _ptr(posqCorrection), _ptr(velm), _ptr(force),
                                        int(paddedNumAtoms), int(precision), _stream(stream)))

    def part2_openmm(self, dt, posq, posqCorrection, velm, force, paddedNumAtoms, precision, stream=None):
        check(self.lib.rbk_part2_openmm(self.h, float(dt), _ptr(posq), _ptr(posqCorrection), _ptr(velm), _ptr(force),
                                        int(paddedNumAtoms), int(precision), _stream(stream)))

    def part2_part1_openmm(self, dt, posq, posqCorrection, velm, force, paddedNumAtoms, precision, stream=None):
        """Part 2 of this step fused with Part 1 of the next on the OpenMM-CUDA formats (rbk_part2_part1_openmm)."""
        check(self.lib.rbk_part2_part1_openmm(self.h, float(dt), _ptr(posq), _ptr(posqCorrection), _ptr(velm), _ptr(force),
                                              int(paddedNumAtoms), int(precision), _stream(stream)))

    def reorder_openmm(self, location, force, paddedNumAtoms, stream=None):
        """The CUDA platform's ReorderListener on the device: permute the owned atoms' forces to the new order and install
        the new plugin-order -> device-index map."""
        loc = None if location is None else np.ascontiguousarray(location, dtype=np.int32)
        check(self.lib.rbk_reorder_openmm(self.h, _i(loc), _ptr(force), int(paddedNumAtoms), _stream(stream)))

    def free_delta_openmm(self, dt, velm, force, paddedNumAtoms, precision, posDelta, stream=None):
        """posDelta.xyz = (v + f invMass dt/2) dt for the free atoms (the hook before integration.applyConstraints)."""
        check(self.lib.rbk_free_delta_openmm(self.h, float(dt), _ptr(velm), _ptr(force), int(paddedNumAtoms), int(precision),
                                             _ptr(posDelta), _stream(stream)))

    def part1_delta_openmm(self, dt, posq, posqCorrection, velm, force, paddedNumAtoms, precision, posDelta, stream=None):
        check(self.lib.rbk_part1_delta_openmm(self.h, float(dt), _ptr(posq), _ptr(posqCorrection), _ptr(velm), _ptr(force),
                                              int(paddedNumAtoms), int(precision), _ptr(posDelta), _stream(stream)))

    def kinetic_openmm(self, velm, precision, stream=None):
        out = np.zeros(2)
        check(self.lib.rbk_kinetic_openmm(self.h, _ptr(velm), int(precision), _d(out), _stream(stream)))
        return out

    # ---- refined ("shadow") energy diagnostics (rbk_refined.cu)
    def set_refined_energies(self, mode=1, stream=None):
        """0 = off, 1 = bodies and free atoms, 2 = bodies only (caller drives free_dot_openmm)."""
        check(self.lib.rbk_set_refined_energies(self.h, int(mode), _stream(stream)))

    def refined_kinetic(self, dt, vel, layout=None, stream=None):
        lay, stride = _layout(vel, layout)
        out = np.zeros(2)
        check(self.lib.rbk_refined_kinetic(self.h, float(dt), _ptr(vel), lay, stride, _d(out), _stream(stream)))
        return out

    def potential_refinement(self, dt, force, layout=None, stream=None):
        lay, stride = _layout(force, layout)
        out = np.zeros(2)
        check(self.lib.rbk_potential_refinement(self.h, float(dt), _ptr(force), lay, stride, _d(out), _stream(stream)))
        return float(out[0])

    def refined_kinetic_openmm(self, dt, velm, precision, stream=None):
        out = np.zeros(2)
        check(self.lib.rbk_refined_kinetic_openmm(self.h, float(dt), _ptr(velm), int(precision), _d(out), _stream(stream)))
        return out

    def potential_refinement_openmm(self, dt, force, paddedNumAtoms, stream=None):
        out = np.zeros(2)
        check(self.lib.rbk_potential_refinement_openmm(self.h, float(dt), _ptr(force), int(paddedNumAtoms), _d(out), _stream(stream)))
        return float(out[0])

    def free_dot_openmm(self, posDelta, precision, factor, restart, stream=None):
        check(self.lib.rbk_free_dot_openmm(self.h, _ptr(posDelta), int(precision), float(factor), int(bool(restart)), _stream(stream)))

    def refined_kinetic_host(self, dt, V, stream=None):
        out = np.zeros(2)
        check(self.lib.rbk_refined_kinetic_host(self.h, float(dt), _ptr(V), _d(out), _stream(stream)))
        return out

    def potential_refinement_host(self, dt, F, stream=None):
        out = np.zeros(2)
        check(self.lib.rbk_potential_refinement_host(self.h, float(dt), _ptr(F), _d(out), _stream(stream)))
        return float(out[0])

    def kinetic_host(self, V, stream=None):
        """Kinetic energies for host-resident velocities [N,3] (numpy or pinned torch CPU tensor)."""
        out = np.zeros(2)
        check(self.lib.rbk_kinetic_host(self.h, _ptr(V), _d(out), _stream(stream)))
        return out

    def download_bodies(self, stream=None):
        nb = self.counts()["numBodies"]
        o = {"rcm": np.zeros((nb, 3)), "pcm": np.zeros((nb, 3)), "q": np.zeros((nb, 4)), "pi": np.zeros((nb, 4)),
             "force": np.zeros((nb, 3)), "torque": np.zeros((nb, 4))}
        check(self.lib.rbk_download_bodies(self.h, *[_d(o[k]) for k in ("rcm", "pcm", "q", "pi", "force", "torque")], _stream(stream)))
        return o

    def execute_host(self, dt, steps, R, V, F, forces=None, stream=None, constrain_positions=None, constrain_velocities=None):
        """R, V, F: host float64 buffers [N,3] (numpy arrays or pinned torch CPU tensors), updated in place; V is written
        once per call (after its last step) and may be None: the velocities then stay on the device until a later call.
        constrain_positions(oldR_ptr, R_ptr, n, user) / constrain_velocities(R_ptr, V_ptr, n, user) are the free-atom
        constraint / virtual-site hooks of rbk_execute_host_hooks; they return nonzero when they changed the array."""
        cb = _lib.FORCE_FN(forces) if forces is not None else C.cast(None, _lib.FORCE_FN)
        if constrain_positions is None and constrain_velocities is None:
            check(self.lib.rbk_execute_host(self.h, float(dt), int(steps), _ptr(R), _ptr(V), _ptr(F), cb, None, _stream(stream)))
            return
        hp = _lib.HOOK_FN(constrain_positions) if constrain_positions is not None else C.cast(None, _lib.HOOK_FN)
        hv = _lib.HOOK_FN(constrain_velocities) if constrain_velocities is not None else C.cast(None, _lib.HOOK_FN)
        check(self.lib.rbk_execute_host_hooks(self.h, float(dt), int(steps), _ptr(R), _ptr(V), _ptr(F), cb, hp, hv, None,
                                              _stream(stream)))
