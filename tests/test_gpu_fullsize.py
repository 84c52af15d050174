"""Parity at the JUDGED sizes (BASELINE.json configs 2-5), stepping the way bench.py does - part1, rbk_part2_part1 x (n-1),
part2 - so that the kernels the benchmark times are the kernels checked here: a random subsample of the bodies (and free
atoms) is re-run on the CPU oracle (bodies are independent under prescribed forces) and must agree to REL_TIGHT, plus
size-independent rigid-body properties on ALL bodies."""
import numpy as np
import pytest

import common
from common import GpuStepper, rel_inf
from oracle.checkers import CpuStepper

pytestmark = pytest.mark.gpu

REL_TIGHT = 2e-10


def subsample(sysd, n_bodies, n_free, seed):
    body = sysd["bodyIndices"]
    rng = np.random.Generator(np.random.Philox(key=seed))
    labels = np.unique(body[body > 0])
    pick = rng.choice(labels, min(n_bodies, labels.shape[0]), replace=False)
    mask = np.isin(body, pick)
    free = np.nonzero(body <= 0)[0]
    if free.shape[0]:
        mask[rng.choice(free, min(n_free, free.shape[0]), replace=False)] = True
    atoms = np.nonzero(mask)[0]
    return atoms, {k: np.ascontiguousarray(sysd[k][atoms]) for k in ("masses", "R", "V", "F", "charges", "bodyIndices")}


def check_against_oracle(sysd, mode, steps, layout="vec3", shuffle=False, dt=0.001, n_bodies=2000, n_free=2000):
    import torch
    s = GpuStepper(sysd["bodyIndices"], sysd["masses"], mode, layout=layout, shuffle=shuffle)
    s.fused = True
    common.init_like_reference(s, sysd)
    norm0 = np.linalg.norm(s.bodies()["q"], axis=1)
    s.step(dt, steps)
    R, V, _ = s.get_state()
    assert np.isfinite(R).all() and np.isfinite(V).all()
    atoms, sub = subsample(sysd, n_bodies, n_free, seed=31 + mode)
    o = CpuStepper("oracle", sub["bodyIndices"], sub["masses"], mode)
    common.init_like_reference(o, sub)
    o.step(dt, steps)
    Ro, Vo, _ = o.get_state()
    eR, eV = rel_inf(R[atoms], Ro), rel_inf(V[atoms], Vo)
    assert eR <= REL_TIGHT and eV <= REL_TIGHT, (eR, eV)
    # kinetic energy from the reconstructed atomic velocities of the BODY atoms + free atoms = KE_t + KE_r
    ke = s.kinetic()
    ke_atoms = 0.5 * float(np.sum(sysd["masses"][:, None] * V * V))
    assert abs(ke_atoms - ke.sum()) <= 1e-10 * ke_atoms, (ke_atoms, ke)
    # |q|: the exact rotation renormalises every step; NO-SQUISH (like the reference's) only preserves the norm the body
    # build left (1 to ~1e-11, set by the orthonormality of the eigenvectors), so it is compared with the initial one
    norm = np.linalg.norm(s.bodies()["q"], axis=1)
    assert np.max(np.abs(norm - (1.0 if mode == 0 else norm0))) < 1e-13
    order = s.sys.series_order()
    s.close()
    del s
    torch.cuda.empty_cache()
    return R, V, eR, eV, order


@pytest.mark.parametrize("mode", [0, 10])
def test_c2_c3_1M_waters_fused(mode):
    """BASELINE configs 2 (mode 0) and 3 (mode 10): 1,000,000 rigid waters stepped with the one-pass kernel."""
    n_mol = 1_000_000
    sysd = common.synth.water_box(n_mol, seed=20240001)
    R, V, eR, eV, order = check_against_oracle(sysd, mode, 3)
    # 1 fs at 300 K: the series ladder leaves its starting rung for the lowest one after the first launch (order 11); the
    # CONSTANT synthetic forces of this test then spin the bodies up, which may send it back to 13 within a few steps
    assert mode != 0 or order in (11, 13)
    Rm = R.reshape(n_mol, 3, 3)
    tol = 1e-12 if mode == 0 else 1e-10          # NO-SQUISH keeps |q| as built (1 to ~1e-11), and the bond lengths with it
    assert np.max(np.abs(np.linalg.norm(Rm[:, 1] - Rm[:, 0], axis=1) - common.synth.R_OH)) < tol
    assert np.max(np.abs(np.linalg.norm(Rm[:, 2] - Rm[:, 1], axis=1) - 2 * common.synth.R_OH * np.sin(0.5 * common.synth.ANGLE_HOH))) < tol
    print(f"1M waters mode {mode}, fused stepping: subsample vs oracle rel err R {eR:.1e} V {eV:.1e}")


def test_c4_mixed_200k_bodies_500k_free():
    """BASELINE config 4: 200,000 bodies of 3-60 atoms (merged labels) + 500,000 interleaved free atoms, mode 0."""
    sysd = common.synth.mixed_system(200_000, 500_000)
    _, _, eR, eV, _ = check_against_oracle(sysd, 0, 3)
    print(f"config 4, fused stepping: subsample vs oracle rel err R {eR:.1e} V {eV:.1e}")


def test_c5_250k_waters_fused():
    """BASELINE config 5's per-GPU replica: 250,000 rigid waters (working set about the size of the L2)."""
    sysd = common.synth.water_box(250_000, seed=20240001)
    _, _, eR, eV, _ = check_against_oracle(sysd, 0, 4)
    print(f"250k waters, fused stepping: subsample vs oracle rel err R {eR:.1e} V {eV:.1e}")


@pytest.mark.parametrize("shuffle,layout", [(True, "soa"), ("molecules", "vec3")])
def test_1M_waters_reordered_fused(shuffle, layout):
    """Reordered atoms at full size: every atom on its own (the one-pass kernel's gather instantiation, SoA planes) and
    whole molecules (the handle re-sorts its bodies to the caller's order and streams)."""
    sysd = common.synth.water_box(1_000_000, seed=20240002)
    _, _, eR, eV, _ = check_against_oracle(sysd, 0, 3, layout=layout, shuffle=shuffle)
    print(f"1M waters reordered ({shuffle}), fused stepping: subsample vs oracle rel err R {eR:.1e} V {eV:.1e}")


@pytest.mark.parametrize("dt_fs", [2.0, 4.0])
def test_long_time_steps_1M_waters(dt_fs):
    """Rigid bodies exist to allow 2-5 fs steps: the mode-0 kernel at 2 and 4 fs against the oracle."""
    sysd = common.synth.water_box(300_000, seed=20240003)
    _, _, eR, eV, order = check_against_oracle(sysd, 0, 4, dt=dt_fs * 1e-3)
    # the device-side series ladder has climbed: order 13 (or 16) at 2 fs, 16 at 4 fs
    assert order == 16 if dt_fs == 4.0 else order in (13, 16), order
    print(f"300k waters, dt {dt_fs} fs: subsample vs oracle rel err R {eR:.1e} V {eV:.1e}, series order {order}")
