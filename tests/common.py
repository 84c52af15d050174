"""Shared helpers for the test-suite: deterministic small systems, the reference call protocol,
error metrics.  Imports the CPU checkers (oracle/) - test infrastructure, never the product."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from openmm_rigidbody_plugin_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# analytic test potential of SURVEY.md section 8c (energy-drift oracle)
TETHER_K = 5000.0
TETHER_E = np.array([300.0, -500.0, 800.0])


def rel_inf(a, b):
    """||a-b||_inf / ||b||_inf"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = float(np.max(np.abs(b))) if b.size else 0.0
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def quat_rel(qa, qb):
    """Quaternion distance up to a global sign per body."""
    qa = np.asarray(qa)
    qb = np.asarray(qb)
    s = np.sign(np.sum(qa * qb, axis=1, keepdims=True))
    s[s == 0] = 1.0
    return rel_inf(qa * s, qb)


def init_like_reference(stepper, sysd, tether=False):
    """Drive a stepper through the reference's own start-up protocol
    (openmmapi/src/RigidBodyIntegrator.cpp:63-74): setPositions -> forces at the initial positions
    -> update(geometry, velocities) with zero velocities; setVelocities -> update(velocities)."""
    n = sysd["R"].shape[0]
    stepper.set_state(sysd["R"], np.zeros((n, 3)), sysd["F"])
    if tether:
        stepper.set_tether(TETHER_K, TETHER_E, sysd["charges"], sysd["R"])
        stepper.compute_forces()
    stepper.update(True, True)
    stepper.set_state(V=sysd["V"])
    stepper.update(False, True)


def _rng(seed):
    return np.random.Generator(np.random.Philox(key=seed))


def edge_cases():
    """Small deterministic systems exercising the index-mapping and degenerate-geometry paths."""
    cases = {}
    rng = _rng(77)

    def mk(body, masses=None, R=None):
        body = np.asarray(body, dtype=np.int32)
        n = body.shape[0]
        masses_ = rng.uniform(1.0, 16.0, n) if masses is None else np.asarray(masses, dtype=np.float64)
        R_ = rng.uniform(-0.3, 0.3, (n, 3)) if R is None else np.asarray(R, dtype=np.float64)
        return {
            "bodyIndices": body, "masses": masses_, "R": np.ascontiguousarray(R_),
            "V": rng.standard_normal((n, 3)) * np.sqrt(synth.KT_300K / masses_)[:, None],
            "F": rng.standard_normal((n, 3)) * 300.0, "charges": rng.uniform(-0.5, 0.5, n),
        }

    cases["free_only"] = (mk([0, 0, 0, 0, 0]), [0, 3])
    cases["single_body"] = (mk([1, 1, 1, 1, 1]), [0, 1, 4])
    # labels with gaps, negatives, zeros, unordered first appearance, one singleton-free region
    cases["gaps_unordered"] = (mk([7, -3, 7, 2, 0, 2, 2, 7, 100, 100, 100, 0, 7, 2, 41, 41, 41, -1]), [0, 2])
    # two diatomics (collinear, dof 5) + a 3-atom straight rod + one free atom; NO-SQUISH only
    R = np.array([[0.0, 0.0, 0.0], [0.11, 0.02, -0.05], [1.0, 1.0, 1.0], [1.05, 0.93, 1.08],
                  [2.0, 0.0, 0.0], [2.1, 0.1, -0.1], [2.2, 0.2, -0.2], [3.0, 3.0, 3.0]])
    cases["linear_bodies"] = (mk([1, 1, 2, 2, 3, 3, 3, 0], R=R), [1, 3])
    # prolate symmetric top (I0 = I1 > I2): four light atoms on a square plus two heavy atoms on the axis.
    # (The oblate arrangement I0 > I1 = I2 makes the reference's eigenvalue rounding produce I1 < I2 by one
    # ulp and its mode-0 rotation NaN; that case is documented in DESIGN.md, not pinned.)
    a, h = 0.05, 0.15
    R = np.array([[a, 0, 0], [-a, 0, 0], [0, a, 0], [0, -a, 0], [0, 0, h], [0, 0, -h]], dtype=np.float64)
    Rrot = R @ np.array([[0.36, 0.48, -0.8], [-0.8, 0.6, 0.0], [0.48, 0.64, 0.6]]).T + 0.5
    cases["symmetric_top"] = (mk([1] * 6, masses=[3.0] * 4 + [12.0] * 2, R=Rrot), [0, 2])
    # massless-free handling is reference-undefined; virtual sites are covered by a separate test
    return cases
