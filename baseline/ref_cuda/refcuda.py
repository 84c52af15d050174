"""ctypes front-end of baseline/ref_cuda/_build/libref_cuda_*.so: the reference's own CUDA kernels
(platforms/cuda/src/kernels/rigidbodyintegrator.cu of /root/reference, compiled in place by baseline/ref_cuda/Makefile).

BENCH / TEST INFRASTRUCTURE ONLY - "the design to beat" timed on the same B200 and used to pin the refined-energy
diagnostics.  Nothing in openmm_rigidbody_plugin_b200 imports this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
PRECISION = {"mixed": 1, "double": 2}


def lib_path(precision: str, mode: int, compmod: int) -> str:
    return os.path.join(_HERE, "_build", f"libref_cuda_p{PRECISION[precision]}_m{mode}_c{compmod}.so")


def available(precision="mixed", mode=0, compmod=0) -> bool:
    return os.path.exists(lib_path(precision, mode, compmod))


def build() -> None:
    """Compile every variant (needs /root/reference and nvcc; no GPU)."""
    subprocess.run(["make", "-s", "-C", _HERE, "-j4"], check=True)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class RefCudaSystem:
    """Body data of one system on the device in the reference's AoS `BodyData` form + its kernels."""

    def __init__(self, precision, mode, compmod, bodies, body_fixed, location, num_free, padded):
        """bodies: dict as returned by DeviceRigidBodySystem.host_bodies() (N, loc, mass, invI, rcm, pcm, force, q, pi, torque)."""
        path = lib_path(precision, mode, compmod)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        lib = self.lib = C.CDLL(path)
        lib.refcuda_create.restype = C.c_void_p
        lib.refcuda_create.argtypes = [C.c_int] * 4
        lib.refcuda_destroy.argtypes = [C.c_void_p]
        lib.refcuda_upload.argtypes = [C.c_void_p, _ip, _ip] + [_dp] * 9 + [_ip]
        lib.refcuda_launch.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.refcuda_time_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_int, C.POINTER(C.c_float)]
        lib.refcuda_kinetic.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _dp]
        lib.refcuda_potential_refinement.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _dp]
        lib.refcuda_download.argtypes = [C.c_void_p] + [_dp] * 6
        assert lib.refcuda_precision() == PRECISION[precision] and lib.refcuda_mode() == mode and lib.refcuda_compmod() == compmod
        self.precision, self.mode, self.compmod = precision, mode, compmod
        self.num_free, self.num_bodies = int(num_free), int(len(bodies["N"]))
        self.num_body_atoms = int(body_fixed.shape[0])
        self.padded = int(padded)
        self.h = lib.refcuda_create(self.num_free, self.num_bodies, self.num_body_atoms, self.padded)
        if not self.h:
            raise MemoryError("refcuda_create failed")
        c = lambda a, t=np.float64: np.ascontiguousarray(a, dtype=t)   # noqa: E731
        args = [c(bodies["N"], np.int32), c(bodies["loc"], np.int32), c(bodies["mass"]), c(bodies["invI"]), c(bodies["rcm"]), c(bodies["pcm"]),
                c(bodies["force"]), c(bodies["q"]), c(bodies["pi"]), c(bodies["torque"]), c(body_fixed), c(location, np.int32)]
        rc = lib.refcuda_upload(self.h, _i(args[0]), _i(args[1]), *[_d(a) for a in args[2:11]], _i(args[11]))
        if rc:
            raise RuntimeError("refcuda_upload failed")
        self.blocks_per_sm = 6                     # CudaContext: numThreadBlocks = 6 x multiprocessors (OpenMM 7.2)

    def close(self):
        if getattr(self, "h", None):
            self.lib.refcuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _launch(self, which, dt, posq, corr, velm, force, restart=0, factor=0):
        if self.lib.refcuda_launch(self.h, which, float(dt), _p(posq), _p(corr), _p(velm), _p(force), self.blocks_per_sm, restart, factor):
            raise RuntimeError("reference CUDA kernel launch failed")

    def part1(self, dt, posq, corr, velm, force):
        """CudaIntegrateRigidBodyStepKernel::execute up to the force evaluation (CudaRigidBodyKernels.cpp:405-422), no constraints."""
        if self.num_free:
            if self.compmod:
                self._launch(0, -dt, posq, corr, velm, force)
                self._launch(1, dt, posq, corr, velm, force, 1, -1)
            self._launch(0, dt, posq, corr, velm, force)
        self._launch(2, dt, posq, corr, velm, force)

    def part2(self, dt, posq, corr, velm, force):
        """... and after it (:426-438)."""
        self._launch(3, dt, posq, corr, velm, force)
        if self.num_free and self.compmod:
            self._launch(1, dt, posq, corr, velm, force, 0, 5)
            self._launch(0, dt, posq, corr, velm, force)
            self._launch(1, dt, posq, corr, velm, force, 0, 2)

    def time_steps(self, dt, steps, posq, corr, velm, force_a, force_b, start_with=0):
        ms = C.c_float(0.0)
        cur = self.lib.refcuda_time_steps(self.h, float(dt), int(steps), _p(posq), _p(corr), _p(velm), _p(force_a), _p(force_b),
                                          self.blocks_per_sm, int(start_with), C.byref(ms))
        if cur < 0:
            raise RuntimeError("reference CUDA kernels failed")
        return float(ms.value), cur

    def kinetic(self, velm, dt=None, refined=False):
        out = np.zeros(2)
        if self.lib.refcuda_kinetic(self.h, _p(velm), int(refined), self.blocks_per_sm, _d(out)):
            raise RuntimeError("refcuda_kinetic failed")
        return out / (6.0 * dt) if refined else out          # CudaRigidBodyKernels.cpp:158-161

    def potential_refinement(self, dt, velm, force):
        out = np.zeros(2)
        if self.lib.refcuda_potential_refinement(self.h, _p(velm), _p(force), self.blocks_per_sm, _d(out)):
            raise RuntimeError("refcuda_potential_refinement failed")
        return -out[0] * dt * dt / 24.0                      # CudaRigidBodyKernels.cpp:490-491

    def bodies(self):
        nb = self.num_bodies
        o = {"r": np.zeros((nb, 3)), "v": np.zeros((nb, 3)), "q": np.zeros((nb, 4)), "pi": np.zeros((nb, 4)), "F": np.zeros((nb, 3)),
             "Ctau": np.zeros((nb, 4))}
        if self.lib.refcuda_download(self.h, *[_d(o[k]) for k in ("r", "v", "q", "pi", "F", "Ctau")]):
            raise RuntimeError("refcuda_download failed")
        return o
