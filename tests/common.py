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


def molecule_permutation(body, seed):
    """A random permutation of whole units - runs of consecutive atoms with the same positive body label, and single
    free atoms - the way OpenMM's reorderAtoms moves whole molecules: returns perm with atom i stored at slot perm[i]."""
    n = body.shape[0]
    start = np.nonzero(np.concatenate([[True], (body[1:] != body[:-1]) | (body[1:] <= 0)]))[0]
    length = np.diff(np.concatenate([start, [n]]))
    order = np.random.Generator(np.random.Philox(key=seed)).permutation(start.shape[0])       # unit order[k] is stored k-th
    new_start = np.empty_like(start)
    new_start[order] = np.concatenate([[0], np.cumsum(length[order])[:-1]])
    return (np.repeat(new_start - start, length) + np.arange(n)).astype(np.int64)


class GpuStepper:
    """Same call protocol as oracle.checkers.CpuStepper, but every step runs librbk's CUDA kernels
    through the C ABI.  layout: 'vec3' ([N,3] like std::vector<Vec3>) or 'soa' ([3,N] planes).
    shuffle=True stores the atoms in a permuted order and hands the permutation to
    rbk_set_atom_location (the CUDA platform's atom reordering)."""

    def __init__(self, bodyIndices, masses, mode=0, layout="vec3", shuffle=False, isVirtual=None, constraints=None, seed=5):
        import torch
        from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem
        self.torch = torch
        self.dev = torch.device("cuda:0")
        self.sys = DeviceRigidBodySystem(bodyIndices, masses, mode, isVirtual=isVirtual, constraints=constraints)
        self.n = len(masses)
        self.layout = layout
        self.perm = None
        if shuffle == "molecules":
            self.perm = molecule_permutation(np.asarray(bodyIndices), seed)
        elif shuffle:
            self.perm = np.random.Generator(np.random.Philox(key=seed)).permutation(self.n)   # atom i lives at perm[i]
        self.hR = np.zeros((self.n, 3))
        self.hV = np.zeros((self.n, 3))
        self.hF = np.zeros((self.n, 3))
        self.dR = self.dV = self.dF = None
        self.tether = None
        self.uploaded = False

    # host <-> device helpers -------------------------------------------------------------
    def _to_dev(self, a):
        t = self.torch
        if self.perm is not None:
            b = np.empty_like(a)
            b[self.perm] = a
            a = b
        x = t.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        return x.t().contiguous() if self.layout == "soa" else x.contiguous()

    def _to_host(self, x):
        a = (x.t() if self.layout == "soa" else x).contiguous().cpu().numpy()
        return a[self.perm] if self.perm is not None else a

    def _natural(self, x):
        """device tensor as [N,3] in ORIGINAL atom order (torch, on device)"""
        a = x.t() if self.layout == "soa" else x
        if self.perm is not None:
            a = a[self.torch.from_numpy(self.perm).to(self.dev)]
        return a

    def set_state(self, R=None, V=None, F=None):
        if R is not None:
            self.hR = np.array(R, dtype=np.float64)
            self.dR = self._to_dev(self.hR)
        if V is not None:
            self.hV = np.array(V, dtype=np.float64)
            self.dV = self._to_dev(self.hV)
        if F is not None:
            self.hF = np.array(F, dtype=np.float64)
            self.dF = self._to_dev(self.hF)

    def get_state(self):
        return self._to_host(self.dR), self._to_host(self.dV), self._to_host(self.dF)

    def set_tether(self, k, E, charges, x0):
        t = self.torch
        self.tether = (float(k), t.tensor(np.asarray(E), dtype=t.float64, device=self.dev),
                       t.from_numpy(np.ascontiguousarray(charges)).to(self.dev),
                       t.from_numpy(np.ascontiguousarray(x0)).to(self.dev))

    def compute_forces(self):
        """Analytic test potential evaluated on the device with the same operation order as the CPU checkers."""
        k, E, ch, x0 = self.tether
        x = self._natural(self.dR)
        dx = x - x0
        F = dx * (-k) + E[None, :] * ch[:, None]
        U = (0.5 * k * (dx * dx).sum(1) - ch * (x * E[None, :]).sum(1)).sum()
        Fh = F
        if self.perm is not None:
            inv = self.torch.empty_like(F)
            inv[self.torch.from_numpy(self.perm).to(self.dev)] = F
            Fh = inv
        self.dF.copy_(Fh.t() if self.layout == "soa" else Fh)
        return float(U)

    def update(self, geometry=True, velocities=True):
        R, V, F = self.get_state() if self.dR is not None else (self.hR, self.hV, self.hF)
        self.sys.update(R, V, F, geometry, velocities)
        self.sys.upload()
        if self.perm is not None:
            loc = self.perm[self.sys.atom_index()].astype(np.int32)
            self.sys.set_atom_location(loc)
        self.uploaded = True

    def part1(self, dt):
        self.sys.part1(dt, self.dR, self.dV, self.dF)

    def part2(self, dt):
        self.sys.part2(dt, self.dR, self.dV, self.dF)

    def step(self, dt, steps=1):
        if getattr(self, "fused", False) and steps > 0:
            # part1, forces, [part2+part1 fused, forces] x (steps-1), part2  (rbk_part2_part1)
            self.sys.part1(dt, self.dR, self.dV, self.dF)
            for _ in range(steps - 1):
                if self.tether is not None:
                    self.compute_forces()
                self.sys.part2_part1(dt, self.dR, self.dV, self.dF)
            if self.tether is not None:
                self.compute_forces()
            self.sys.part2(dt, self.dR, self.dV, self.dF)
            return
        for _ in range(steps):
            self.sys.part1(dt, self.dR, self.dV, self.dF)
            if self.tether is not None:
                self.compute_forces()
            self.sys.part2(dt, self.dR, self.dV, self.dF)

    def kinetic(self):
        return self.sys.kinetic(self.dV)

    def counts(self):
        return self.sys.counts()

    def body_index(self):
        return self.sys.body_index()

    def atom_index(self):
        return self.sys.atom_index()

    def bodies(self):
        b = self.sys.host_bodies()
        if self.uploaded:
            b.update(self.sys.download_bodies())
        return b

    def body_fixed(self):
        return self.sys.body_fixed()

    def close(self):
        self.sys.close()
