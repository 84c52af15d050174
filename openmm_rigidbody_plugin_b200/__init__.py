"""B200-native RigidBodyIntegrator step (drop-in for one hot path of craabreu/openmm_rigidbody_plugin).

The arithmetic lives in lib/librbk.so (hand-written CUDA for sm_100a behind the C ABI in
include/rbk.h); this package is the host-side mirror of the reference's interface.  Importing the
package does not load the library; constructing any system does, and fails loudly if it is missing.
"""
from ._lib import RBK_LAYOUT_SOA, RBK_LAYOUT_VEC3, OpenMMException, RbkError  # noqa: F401
from .integrator import Context, HarmonicBondForce, RigidBodyIntegrator, RigidBodySystem, State, System  # noqa: F401
from .system import DeviceRigidBodySystem  # noqa: F401
from . import serialization  # noqa: F401
from .forcefield import ForceField  # noqa: F401
from .statedatareporter import StateDataReporter  # noqa: F401

__all__ = ["DeviceRigidBodySystem", "RbkError", "OpenMMException", "RBK_LAYOUT_VEC3", "RBK_LAYOUT_SOA",
           "RigidBodyIntegrator", "RigidBodySystem", "System", "Context", "State", "HarmonicBondForce",
           "ForceField", "StateDataReporter", "serialization"]
