"""The C++ plugin glue (openmm_rigidbody_plugin_b200/glue: RigidBodyIntegrator, IntegrateRigidBodyStepKernel, B200
kernel + factory with the extern "C" registration symbols, serialization proxy) compiled against the OpenMM header
shim and linked to librbk.so.  CPU: it builds, exports the plugin entry points, the reference's serialization test
passes, and stepping without a GPU fails loudly.  GPU: the reference-style C++ integration tests run."""
import os
import subprocess

import pytest

import common

LIB = os.path.join(common.ROOT, "openmm_rigidbody_plugin_b200", "lib")


@pytest.fixture(scope="module")
def glue():
    import __graft_entry__ as g
    g.build()
    return LIB


def run(path):
    return subprocess.run([path], capture_output=True, text=True, timeout=600)


def test_glue_builds_and_exports_plugin_entry_points(glue):
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(glue, "libRigidBodyPluginB200.so")],
                         capture_output=True, text=True, check=True).stdout
    for sym in ("registerPlatforms", "registerKernelFactories", "registerRigidBodyB200KernelFactories",
                "registerRigidBodySerializationProxies"):
        assert f" T {sym}" in out, sym
    undefined = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(glue, "libRigidBodyPluginB200.so")],
                               capture_output=True, text=True, check=True).stdout
    for sym in ("rbk_create", "rbk_update", "rbk_upload", "rbk_execute_host", "rbk_kinetic_host"):
        assert sym in undefined, sym                      # the glue reaches the kernels only through the C ABI


def test_serialization_proxy_round_trip(glue):
    r = run(os.path.join(glue, "TestSerializeRigidBodyIntegrator"))
    assert r.returncode == 0 and r.stdout.strip().endswith("Done"), r.stdout + r.stderr


def test_stepping_without_gpu_fails_loudly(glue):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run(os.path.join(glue, "TestB200RigidBodyIntegrator"))
    assert r.returncode == 1 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_integration_tests_on_gpu(glue):
    r = run(os.path.join(glue, "TestB200RigidBodyIntegrator"))
    assert r.returncode == 0 and r.stdout.strip().endswith("Done"), r.stdout + r.stderr


def test_cmake_packaging_configures_and_builds_against_the_shim(glue, tmp_path):
    """glue/CMakeLists.txt (OPENMM_DIR-based, like the reference's CMake): without an OpenMM installation it must fall
    back to the header shim and build the plugin library and the three test programs."""
    import shutil
    if shutil.which("cmake") is None:
        pytest.skip("cmake not available")
    src = os.path.join(common.ROOT, "openmm_rigidbody_plugin_b200", "glue")
    cfg = subprocess.run(["cmake", "-S", src, "-B", str(tmp_path), "-DOPENMM_DIR=/nonexistent"], capture_output=True, text=True)
    assert cfg.returncode == 0, cfg.stdout + cfg.stderr
    assert "header shim" in cfg.stdout
    bld = subprocess.run(["cmake", "--build", str(tmp_path), "-j", "8"], capture_output=True, text=True, timeout=900)
    assert bld.returncode == 0, bld.stdout[-2000:] + bld.stderr[-2000:]
    assert os.path.exists(os.path.join(str(tmp_path), "libRigidBodyPluginB200.so"))


def test_glue_registers_on_the_cuda_platform(glue):
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(glue, "libRigidBodyPluginB200.so")],
                         capture_output=True, text=True, check=True).stdout
    assert " T registerRigidBodyCudaKernelFactories" in out          # the symbol the reference's CUDA plugin exports
    undefined = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(glue, "libRigidBodyPluginB200.so")],
                               capture_output=True, text=True, check=True).stdout
    for sym in ("rbk_part1_openmm", "rbk_part2_openmm", "rbk_part2_part1_openmm", "rbk_reorder_openmm", "rbk_update_device_openmm",
                "rbk_free_delta_openmm", "rbk_part1_delta_openmm", "rbk_kinetic_openmm"):
        assert sym in undefined, sym


@pytest.mark.gpu
def test_cuda_platform_kernel_on_gpu_and_against_the_oracle(glue, tmp_path):
    """The CUDA-platform KernelImpl (B200CudaIntegrateRigidBodyStepKernel) driven through RigidBodyIntegrator::step on
    the shim CudaContext: the C++ program runs the reference-style tests in mixed and double precision, checks both
    kernels of the plugin against each other, asserts zero host<->device copies by librbk inside step(n), and dumps one
    run (700 tethered waters, fused step(8) x 3 with the atoms reordered every 3 steps, mixed precision) that is
    replayed here on the CPU oracle."""
    import numpy as np
    from oracle.checkers import CpuStepper
    dump = str(tmp_path / "cuda_run.bin")
    r = subprocess.run([os.path.join(glue, "TestB200CudaRigidBodyIntegrator"), dump], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("Done"), r.stdout + r.stderr
    raw = np.fromfile(dump, dtype=np.float64)
    n, n_mol, mode, steps, dt, k = int(raw[0]), int(raw[1]), int(raw[2]), int(raw[3]), raw[4], raw[5]
    ke = raw[6:8]
    field = raw[8:11]
    o = 11
    charges = raw[o:o + n]; o += n
    R0 = raw[o:o + 3 * n].reshape(n, 3); o += 3 * n
    V0 = raw[o:o + 3 * n].reshape(n, 3); o += 3 * n
    R1 = raw[o:o + 3 * n].reshape(n, 3); o += 3 * n
    V1 = raw[o:o + 3 * n].reshape(n, 3); o += 3 * n
    assert o == raw.shape[0] and n == 3 * n_mol
    masses = np.tile([15.99943, 1.007947, 1.007947], n_mol)
    body = np.repeat(np.arange(1, n_mol + 1, dtype=np.int32), 3)
    s = CpuStepper("oracle", body, masses, mode)
    # the CUDA platform holds forces in fixed point (2^-32 kJ/mol/nm): the oracle steps with exact fp64 forces, which
    # bounds the agreement at ~1e-10 relative per force evaluation
    s.set_state(R0, np.zeros((n, 3)), np.zeros((n, 3)))
    s.set_tether(k, field, charges, R0)
    s.compute_forces()
    s.update(True, True)
    s.set_state(V=V0)
    s.update(False, True)
    s.step(dt, steps)
    Ro, Vo, _ = s.get_state()
    eR, eV = common.rel_inf(R1, Ro), common.rel_inf(V1, Vo)
    # max over 700 bodies x 24 steps with position-dependent forces: dominated by the few rotor states where the
    # reference's elliptic-integral route (= the oracle) itself keeps only 9-11 digits ("reduction
    # axis"); the two kernels of the plugin agree with each other to 1e-11 on this run (asserted in the C++ program)
    assert eR <= 2e-8 and eV <= 5e-7, (eR, eV)
    assert common.rel_inf(ke, s.kinetic()) <= 1e-7
