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
