"""Generate the committed golden fixtures from the TRUE reference (oracle/_ref/librb_ref.so).

Run in the CPU container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

Every fixture stores its inputs next to the reference outputs, so the tests do not depend on the
random-number generator reproducing them.  Also writes special-function known answers computed
independently with mpmath (50 digits), which pin oracle/rb_oracle.c's elliptic functions without
going through the reference at all.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
from common import synth  # noqa: E402
from oracle.checkers import CpuStepper  # noqa: E402

DT = 0.001


def snapshot(s, prefix, out, bodies=True):
    R, V, _ = s.get_state()
    out[prefix + "R"] = R
    out[prefix + "V"] = V
    out[prefix + "KE"] = s.kinetic()
    if bodies:
        b = s.bodies()
        for k in ("rcm", "pcm", "q", "pi", "force", "torque"):
            out[prefix + k] = b[k]


def run_case(name, sysd, modes, checkpoints, tether=False, full_bodies=True):
    for mode in modes:
        s = CpuStepper("reference", sysd["bodyIndices"], sysd["masses"], mode)
        common.init_like_reference(s, sysd, tether=tether)
        out = {k: sysd[k] for k in ("bodyIndices", "masses", "R", "V", "F", "charges")}
        out["mode"] = np.int32(mode)
        out["dt"] = np.float64(DT)
        out["tether"] = np.int32(tether)
        c = s.counts()
        out["counts"] = np.array([c[k] for k in ("numBodies", "numFree", "numActualAtoms", "numBodyAtoms", "numDOF")], np.int32)
        out["cleanIndex"] = s.body_index()
        out["atomIndex"] = s.atom_index()
        b = s.bodies()
        for k in ("N", "dof", "loc", "mass", "I", "invI"):
            out["b0_" + k] = b[k]
        out["b0_d"] = s.body_fixed()
        snapshot(s, "s0_", out, bodies=full_bodies)
        done = 0
        for cp in checkpoints:
            s.step(DT, cp - done)
            done = cp
            snapshot(s, f"s{cp}_", out, bodies=full_bodies)
        out["checkpoints"] = np.array(checkpoints, np.int32)
        path = os.path.join(HERE, f"{name}_mode{mode}.npz")
        np.savez_compressed(path, **out)
        print(f"{path}: {os.path.getsize(path)/1e3:.0f} kB")
        s.close()


def _energy_series(sysd, mode, steps, every):
    s = CpuStepper("reference", sysd["bodyIndices"], sysd["masses"], mode)
    common.init_like_reference(s, sysd, tether=True)
    E = []
    for i in range(steps // every + 1):
        U = s.compute_forces()
        ke = s.kinetic()
        E.append([i * every * DT, U, ke[0], ke[1]])
        if i < steps // every:
            s.step(DT, every)
    return s, E


def drift_case(name, sysd, modes, steps=10000, every=50, twin=False):
    """twin=True also records the TRUE reference run from initial velocities scaled by (1 + 1e-13): how far the
    reference's own energy series and fitted slope move under a perturbation at rounding level (the dynamics is chaotic)."""
    for mode in modes:
        s, E = _energy_series(sysd, mode, steps, every)
        out = {k: sysd[k] for k in ("bodyIndices", "masses", "R", "V", "F", "charges")}
        out.update(mode=np.int32(mode), dt=np.float64(DT), every=np.int32(every), series=np.array(E))
        if twin:
            s2, E2 = _energy_series(dict(sysd, V=sysd["V"] * (1.0 + 1e-13)), mode, steps, every)
            out["series_twin"] = np.array(E2)
            s2.close()
        R, V, _ = s.get_state()
        out["R_end"], out["V_end"] = R, V
        path = os.path.join(HERE, f"{name}_mode{mode}.npz")
        np.savez_compressed(path, **out)
        e = np.array(E)
        tot = e[:, 1] + e[:, 2] + e[:, 3]
        slope = np.polyfit(e[:, 0], tot, 1)[0]
        print(f"{path}: E0={tot[0]:.6f} slope={slope:.4e} kJ/mol/ps rms={np.std(tot - np.polyval(np.polyfit(e[:,0], tot, 1), e[:,0])):.3e}")
        s.close()


def gravity_case(name, sysd, modes, steps=10000, every=250):
    """Integrable 10k-step run: fixed forces F_i = m_i g (no torque).  E = -sum F.x + KE."""
    gvec = np.array([0.3, -0.5, 0.8])
    sysd = dict(sysd)
    sysd["F"] = sysd["masses"][:, None] * gvec[None, :]
    for mode in modes:
        s = CpuStepper("reference", sysd["bodyIndices"], sysd["masses"], mode)
        common.init_like_reference(s, sysd)
        E = []
        for i in range(steps // every + 1):
            R, V, _ = s.get_state()
            ke = s.kinetic()
            E.append([i * every * DT, -float(np.sum(sysd["F"] * R)), ke[0], ke[1]])
            if i < steps // every:
                s.step(DT, every)
        out = {k: sysd[k] for k in ("bodyIndices", "masses", "R", "V", "F", "charges")}
        out.update(mode=np.int32(mode), dt=np.float64(DT), every=np.int32(every), series=np.array(E))
        out["R_end"], out["V_end"] = s.get_state()[:2]
        path = os.path.join(HERE, f"{name}_mode{mode}.npz")
        np.savez_compressed(path, **out)
        e = np.array(E)
        tot = e[:, 1:].sum(1)
        print(f"{path}: E0={tot[0]:.6f} max|E-E0|={np.max(np.abs(tot - tot[0])):.3e}")
        s.close()


def special_function_kats():
    import mpmath as mp
    mp.mp.dps = 50
    rng = np.random.Generator(np.random.Philox(key=99))
    rows_j, rows_f, rows_jj, rows_c = [], [], [], []
    for _ in range(200):
        u = float(rng.uniform(-6.0, 6.0))
        m = float(rng.uniform(1e-6, 1.0 - 1e-6))
        rows_j.append([u, m, float(mp.ellipfun("sn", u, m=m)), float(mp.ellipfun("cn", u, m=m)), float(mp.ellipfun("dn", u, m=m))])
        x, y, z = (float(v) for v in rng.uniform(0.0, 2.0, 3))
        p = float(rng.uniform(0.05, 3.0))
        rows_f.append([x, y, z, float(mp.elliprf(x, y, z))])
        rows_jj.append([x, y, z, p, float(mp.elliprj(x, y, z, p))])
        rows_c.append([x, y + 1e-3, float(mp.elliprc(x, y + 1e-3))])
    path = os.path.join(HERE, "special_functions_mpmath.npz")
    np.savez_compressed(path, jacobi=np.array(rows_j), rf=np.array(rows_f), rj=np.array(rows_jj), rc=np.array(rows_c))
    print(path)


def main():
    # BASELINE config 1: 4,096 TIP3P waters, mode 0, one step (positions/velocities only)
    run_case("c1_water4096", synth.water_box(4096), [0], [1], full_bodies=False)
    # small water box, all three rotation modes, full body state after 1/10/100 steps
    run_case("water256", synth.water_box(256, seed=11), [0, 1, 10], [1, 10, 100])
    # ragged bodies (3..40 atoms, merged labels with gaps) + interleaved free atoms
    run_case("mixed120", synth.mixed_system(120, 200, seed=12, max_atoms=40), [0, 3], [1, 10])
    # position-dependent forces (tether + field)
    run_case("water64_tether", synth.water_box(64, seed=13), [0, 10], [1, 50], tether=True)
    for nm, (sysd, modes) in common.edge_cases().items():
        run_case("edge_" + nm, sysd, modes, [1, 5])
    drift_case("drift_water128", synth.water_box(128, seed=14), [0, 10])
    # SURVEY.md section 8c's probe size: 512 waters, with the reference's own rounding-level twin
    drift_case("drift_water512", synth.water_box(512, seed=14), [0], twin=True)
    gravity_case("gravity_water128", synth.water_box(128, seed=15), [0, 10])
    special_function_kats()


if __name__ == "__main__":
    main()
