"""Synthetic rigid-body systems shared by tests/ and bench.py.

All generators are deterministic functions of (size, seed) through numpy's counter-based Philox
bit generator, so the CPU checkers and the CUDA path see bit-identical inputs on every box.
Units follow OpenMM: nm, ps, amu, kJ/mol.
"""
from __future__ import annotations

import numpy as np

KT_300K = 2.4943387854  # kJ/mol at 300 K
# TIP3P geometry / masses / charges
R_OH = 0.09572
ANGLE_HOH = np.deg2rad(104.52)
M_O, M_H = 15.99943, 1.007947
Q_O, Q_H = -0.834, 0.417


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=int(seed)))


def _rotate(quat: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Rotate vectors x[n,k,3] by unit quaternions quat[n,4] (scalar first)."""
    w, v = quat[:, None, 0:1], quat[:, None, 1:4]
    t = 2.0 * np.cross(v, x)
    return x + w * t + np.cross(v, t)


def water_box(n_mol: int, seed: int = 20240001, force_sigma: float = 500.0, spacing: float = 0.31):
    """n_mol rigid TIP3P waters on a simple-cubic lattice, random orientations, 300 K atomic
    velocities and fixed Gaussian forces.  Atom order O,H,H; bodyIndices[3m+k] = m+1."""
    rng = _rng(seed)
    n = int(n_mol)
    side = int(np.ceil(n ** (1.0 / 3.0)))
    idx = np.arange(n)
    centre = np.stack([idx % side, (idx // side) % side, idx // (side * side)], axis=1) * spacing
    half = 0.5 * ANGLE_HOH
    site = np.array([[0.0, 0.0, 0.0],
                     [R_OH * np.sin(half), 0.0, R_OH * np.cos(half)],
                     [-R_OH * np.sin(half), 0.0, R_OH * np.cos(half)]])
    quat = rng.standard_normal((n, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    pos = centre[:, None, :] + _rotate(quat, np.broadcast_to(site, (n, 3, 3)))
    masses = np.tile(np.array([M_O, M_H, M_H]), n)
    charges = np.tile(np.array([Q_O, Q_H, Q_H]), n)
    vel = rng.standard_normal((3 * n, 3)) * np.sqrt(KT_300K / masses)[:, None]
    force = rng.standard_normal((3 * n, 3)) * force_sigma
    body = np.repeat(np.arange(1, n + 1, dtype=np.int32), 3)
    return {
        "masses": np.ascontiguousarray(masses),
        "bodyIndices": np.ascontiguousarray(body, dtype=np.int32),
        "R": np.ascontiguousarray(pos.reshape(3 * n, 3)),
        "V": np.ascontiguousarray(vel),
        "F": np.ascontiguousarray(force),
        "charges": np.ascontiguousarray(charges),
    }


def mixed_system(n_bodies: int, n_free: int, seed: int = 20240004, min_atoms: int = 3, max_atoms: int = 60,
                 force_sigma: float = 500.0, shuffle_free: bool = True):
    """BASELINE config 4: n_bodies rigid bodies of min_atoms..max_atoms atoms (Gaussian clouds,
    sigma 0.15 nm, masses U[1,16]) plus n_free free atoms of mass 12.  Bodies are assembled the way
    the reference's Python layer does it: 3-atom "template" fragments with consecutive labels are
    merged via a mergeList-style union to min(label) (python/forcefield.py:96-105), which leaves
    non-contiguous body labels for cleanBodyIndices to compact.  Free atoms are interleaved between
    bodies so the atom->body map is not the identity."""
    rng = _rng(seed)
    nb = int(n_bodies)
    sizes = rng.integers(min_atoms, max_atoms + 1, size=nb)
    # fragment labels: body b is built from ceil(size/3) fragments labelled consecutively
    nfrag = (sizes + 2) // 3
    first_label = np.concatenate([[1], 1 + np.cumsum(nfrag)[:-1]])
    n_body_atoms = int(sizes.sum())
    # atoms of body b all get the merged label min(set) = first_label[b]
    body_of_atom = np.repeat(np.arange(nb), sizes)
    label = first_label[body_of_atom].astype(np.int32)
    spacing = 1.2
    side = int(np.ceil(max(nb, 1) ** (1.0 / 3.0)))
    ib = np.arange(nb)
    centre = np.stack([ib % side, (ib // side) % side, ib // (side * side)], axis=1) * spacing
    cloud = rng.standard_normal((n_body_atoms, 3)) * 0.15
    pos_b = centre[body_of_atom] + cloud
    mass_b = rng.uniform(1.0, 16.0, size=n_body_atoms)
    # free atoms
    nf = int(n_free)
    pos_f = rng.uniform(0.0, side * spacing, size=(nf, 3))
    mass_f = np.full(nf, 12.0)
    # interleave: free atom k is inserted after body (k mod nb) when shuffle_free
    N = n_body_atoms + nf
    if shuffle_free and nf > 0 and nb > 0:
        owner_f = (np.arange(nf) % nb)
        key_b = body_of_atom.astype(np.float64)
        key_f = owner_f.astype(np.float64) + 0.5
        order = np.argsort(np.concatenate([key_b, key_f]), kind="stable")
    else:
        order = np.arange(N)
    pos = np.concatenate([pos_b, pos_f])[order]
    masses = np.concatenate([mass_b, mass_f])[order]
    body = np.concatenate([label, np.zeros(nf, dtype=np.int32)])[order]
    vel = rng.standard_normal((N, 3)) * np.sqrt(KT_300K / masses)[:, None]
    force = rng.standard_normal((N, 3)) * force_sigma
    charges = rng.uniform(-0.5, 0.5, size=N)
    return {
        "masses": np.ascontiguousarray(masses),
        "bodyIndices": np.ascontiguousarray(body, dtype=np.int32),
        "R": np.ascontiguousarray(pos),
        "V": np.ascontiguousarray(vel),
        "F": np.ascontiguousarray(force),
        "charges": np.ascontiguousarray(charges),
    }


def algorithmic_bytes(n_bodies: int, n_body_atoms: int, n_free: int) -> int:
    """Compulsory HBM bytes of one integrator step: 560 per body + 128 per
    body atom + 288 per free atom."""
    return 560 * int(n_bodies) + 128 * int(n_body_atoms) + 288 * int(n_free)
