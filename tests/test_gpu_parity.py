"""GPU parity tests: librbk's CUDA kernels, called through the C ABI, against
  (a) the committed golden fixtures generated from the TRUE reference,
  (b) the oracle on the same seeded inputs (sizes the oracle finishes in seconds),
  (c) size-independent properties at BASELINE.json's full sizes (1M rigid waters).
Tolerances: integer/index work bit-exact; positions and velocities within 1e-6 relative after one
step (BASELINE.json north star) - and we additionally hold the much tighter REL_TIGHT below, which is
what an fp64 implementation of the same mathematics should reach."""
import glob
import os

import numpy as np
import pytest

import common
from common import GOLDEN_DIR, GpuStepper, quat_rel, rel_inf
from oracle.checkers import CpuStepper

pytestmark = pytest.mark.gpu

REL_NORTH_STAR = 1e-6
REL_TIGHT = 2e-10

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*_mode*.npz"))
               if not os.path.basename(p).startswith(("drift_", "gravity_")))


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def sysd_of(g):
    return {k: g[k] for k in ("bodyIndices", "masses", "R", "V", "F", "charges")}


def insert_free_atoms(sysd, rng, every):
    """a copy of a water system with one free atom (mass 35.45) inserted after every `every`-th molecule"""
    n = sysd["masses"].shape[0]
    after = np.arange(3 * every - 1, n, 3 * every)                 # insert behind these atoms
    out = {}
    k = after.shape[0]
    extra = {"masses": np.full(k, 35.45), "bodyIndices": np.zeros(k, np.int32), "charges": np.full(k, -1.0),
             "R": sysd["R"][after] + 0.15, "V": rng.standard_normal((k, 3)) * 0.26, "F": rng.standard_normal((k, 3)) * 300.0}
    for key in ("masses", "bodyIndices", "charges", "R", "V", "F"):
        out[key] = np.ascontiguousarray(np.insert(sysd[key], after + 1, extra[key], axis=0))
    return out


def compare_state(tag, s, ref_R, ref_V, ref_KE, ref_b=None, tol=REL_TIGHT):
    R, V, _ = s.get_state()
    eR, eV = rel_inf(R, ref_R), rel_inf(V, ref_V)
    assert eR <= REL_NORTH_STAR and eV <= REL_NORTH_STAR, (tag, eR, eV)
    assert eR <= tol and eV <= tol, (tag, eR, eV)
    ke = s.kinetic()
    assert rel_inf(ke, ref_KE) <= tol, (tag, ke, ref_KE)
    if ref_b is not None:
        b = s.bodies()
        for k in ("rcm", "pcm", "pi", "force", "torque"):
            assert rel_inf(b[k], ref_b[k]) <= 50 * tol, (tag, k, rel_inf(b[k], ref_b[k]))
        assert quat_rel(b["q"], ref_b["q"]) <= tol, (tag, "q")


@pytest.mark.parametrize("layout,shuffle", [("vec3", False), ("soa", True)])
@pytest.mark.parametrize("name", CASES)
def test_golden_from_true_reference(name, layout, shuffle):
    g = load(name)
    s = GpuStepper(g["bodyIndices"], g["masses"], int(g["mode"]), layout=layout, shuffle=shuffle)
    common.init_like_reference(s, sysd_of(g), tether=bool(g["tether"]))
    c = s.counts()
    assert [c[k] for k in ("numBodies", "numFree", "numActualAtoms", "numBodyAtoms", "numDOF")] == g["counts"].tolist()
    assert np.array_equal(s.body_index(), g["cleanIndex"])
    assert np.array_equal(s.atom_index(), g["atomIndex"])
    hb = s.sys.host_bodies()
    for k in ("N", "dof", "loc"):
        assert np.array_equal(hb[k], g["b0_" + k])
    for k in ("mass", "I", "invI"):
        assert rel_inf(hb[k], g["b0_" + k]) <= 1e-13
    assert rel_inf(s.body_fixed(), g["b0_d"]) <= 1e-13
    done = 0
    for cp in [0] + g["checkpoints"].tolist():
        s.step(float(g["dt"]), cp - done)
        done = cp
        ref_b = {k: g[f"s{cp}_{k}"] for k in ("rcm", "pcm", "q", "pi", "force", "torque")} if f"s{cp}_q" in g else None
        # error grows with the number of steps (chaotic amplification is absent here, rounding accumulates)
        tol = REL_TIGHT * max(1, cp) ** 0.5 * (10 if cp >= 50 else 1)
        compare_state((name, layout, cp), s, g[f"s{cp}_R"], g[f"s{cp}_V"], g[f"s{cp}_KE"], ref_b, tol)
    s.close()


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("mode", [0, 1, 10])
@pytest.mark.parametrize("layout,shuffle", [("vec3", False), ("soa", False), ("vec3", "molecules")])
def test_water_vs_oracle(mode, layout, shuffle, fused):
    """fused=True steps with rbk_part2_part1 (one pass per step) instead of separate part1/part2 launches.
    shuffle="molecules": the caller stores whole molecules in a permuted order (what OpenMM's reorderAtoms produces) - the
    handle then keeps its own bodies sorted by the caller's slots and streams through the caller's arrays."""
    sysd = common.synth.water_box(20000, seed=100 + mode)
    if shuffle:        # + free atoms (ions) sprinkled between the waters: they are re-sorted as well
        rng = np.random.Generator(np.random.Philox(key=7))
        sysd = insert_free_atoms(sysd, rng, every=37)
    o = CpuStepper("oracle", sysd["bodyIndices"], sysd["masses"], mode)
    s = GpuStepper(sysd["bodyIndices"], sysd["masses"], mode, layout=layout, shuffle=shuffle)
    s.fused = fused
    for st in (o, s):
        common.init_like_reference(st, sysd)
        st.step(0.001, 3)
    R, V, _ = o.get_state()
    compare_state(("water", mode, layout), s, R, V, o.kinetic(), o.bodies())


@pytest.mark.parametrize("mode", [0, 4])
@pytest.mark.parametrize("layout,shuffle", [("vec3", False), ("vec3", True), ("soa", False)])
def test_mixed_vs_oracle(mode, layout, shuffle):
    """BASELINE config 4 at reduced size: ragged bodies of 3..60 atoms (merged labels) + free atoms."""
    sysd = common.synth.mixed_system(4000, 10000, seed=200 + mode)
    o = CpuStepper("oracle", sysd["bodyIndices"], sysd["masses"], mode)
    s = GpuStepper(sysd["bodyIndices"], sysd["masses"], mode, layout=layout, shuffle=shuffle)
    for st in (o, s):
        common.init_like_reference(st, sysd, tether=True)
        st.step(0.001, 4)
    assert s.counts() == {k: v for k, v in o.counts().items() if k != "numAtoms"}
    R, V, _ = o.get_state()
    compare_state(("mixed", mode, layout, shuffle), s, R, V, o.kinetic(), o.bodies())


@pytest.mark.parametrize("dt", [0.004, 0.02, 0.1])
def test_exact_rotation_fallback_paths(dt):
    """Mode 0 with long steps / fast rotors: the order-14 series fails its truncation check for a growing share of the
    bodies, which then take 2 or 4 exact sub-steps and finally the complete elliptic-integral route on the device.
    All routes must agree with the oracle (the reference algorithm) on every body."""
    sysd = common.synth.water_box(6000, seed=300)
    sysd = dict(sysd, V=sysd["V"] * 2.0)                    # hotter than 300 K: more bodies off the fast path
    o = CpuStepper("oracle", sysd["bodyIndices"], sysd["masses"], 0)
    s = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
    for st in (o, s):
        common.init_like_reference(st, sysd)
        st.step(dt, 2)
    R, V, _ = o.get_state()
    assert np.isfinite(R).all()
    # at 100 fs the rotation angles reach many radians and the reference's own Omega/atan differences lose ~7 digits,
    # so two correct evaluations (different libm, FMA contraction) agree to ~1e-8 there; still 100x inside the 1e-6 bar
    compare_state(("fallback", dt), s, R, V, o.kinetic(), o.bodies(), tol=5e-9 if dt <= 0.02 else 1e-7)


def test_huge_body_and_tile_boundaries():
    """One 5000-atom body (spans many warps and tiles of its own), bodies of exactly 32/33/64/128 atoms and
    a run of 300 tiny bodies: exercises every branch of the segmented reduction."""
    rng = np.random.Generator(np.random.Philox(key=9))
    sizes = [5000, 32, 33, 64, 128, 1, 2] + [3] * 300 + [31, 97, 1025]
    body = np.concatenate([np.full(n, i + 1, np.int32) for i, n in enumerate(sizes)])
    sizes = [n for n in sizes if n >= 3] if False else sizes
    n = body.shape[0]
    masses = rng.uniform(1.0, 16.0, n)
    centres = rng.uniform(0, 20, (len(sizes), 3))
    R = centres[body - 1] + rng.standard_normal((n, 3)) * 0.2
    # 1- and 2-atom bodies are degenerate in the reference (NaN); turn them into free atoms
    for i, sz in enumerate(sizes):
        if sz < 3:
            body[body == i + 1] = 0
    sysd = {"bodyIndices": body, "masses": masses, "R": R, "V": rng.standard_normal((n, 3)) * 0.3,
            "F": rng.standard_normal((n, 3)) * 100.0, "charges": rng.uniform(-0.5, 0.5, n)}
    for mode in (0, 2):
        o = CpuStepper("oracle", body, masses, mode)
        s = GpuStepper(body, masses, mode, layout="vec3", shuffle=True)
        for st in (o, s):
            common.init_like_reference(st, sysd, tether=True)
            st.step(0.0005, 3)
        R1, V1, _ = o.get_state()
        compare_state(("huge", mode), s, R1, V1, o.kinetic(), o.bodies(), tol=2e-9)


def test_c1_config_one_step():
    """BASELINE config 1: 4,096 TIP3P waters, mode 0, one step, against the true reference."""
    g = load("c1_water4096_mode0")
    s = GpuStepper(g["bodyIndices"], g["masses"], 0)
    common.init_like_reference(s, sysd_of(g))
    assert s.counts()["numDOF"] == 24576
    s.step(float(g["dt"]), 1)
    compare_state("c1", s, g["s1_R"], g["s1_V"], g["s1_KE"])


def test_bitwise_reproducible_and_layout_independent():
    sysd = common.synth.mixed_system(2000, 3000, seed=77)
    outs = []
    for layout, shuffle in (("vec3", False), ("vec3", False), ("soa", True)):
        s = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0, layout=layout, shuffle=shuffle)
        common.init_like_reference(s, sysd)
        s.step(0.001, 5)
        R, V, _ = s.get_state()
        outs.append((R, V, s.kinetic()))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)                      # run-to-run: bit identical (no atomics)
    for a, b in zip(outs[0], outs[2]):
        assert np.array_equal(a, b)                      # layout / atom order do not change the arithmetic


def test_energy_drift_tracks_reference():
    """10k steps in the analytic tether+field potential, against the series recorded
    from the TRUE reference.  The dynamics is chaotic (rounding differences grow ~e^(2.8 t/ps)), so after
    ~2 ps the two runs are different samples of the same process and are compared as such:
      (a) while the trajectories still coincide (first 1,500 steps) E(t) must track the reference to 1e-9;
      (b) the energy-drift measure - the rms excursion of E(t) from E(0) over all 10k steps, which is what
          the integrator's accuracy controls - must be within 10% of the reference's (BASELINE.json bar);
      (c) the fitted linear slope must be statistically indistinguishable from the reference's: both are
          far below the slope's own standard error (reference 1.7e-4 vs s.e. 4.6e-3 kJ/mol/ps), so a
          relative bar is meaningless there; we require |slope_gpu - slope_ref| <= 3*sqrt(2)*s.e."""
    for mode in (0, 10):
        g = load(f"drift_water128_mode{mode}")
        s = GpuStepper(g["bodyIndices"], g["masses"], mode)
        common.init_like_reference(s, sysd_of(g), tether=True)
        every, dt = int(g["every"]), float(g["dt"])
        ref = g["series"]
        series = []
        for i in range(ref.shape[0]):
            U = s.compute_forces()
            ke = s.kinetic()
            series.append([i * every * dt, U, ke[0], ke[1]])
            if i < ref.shape[0] - 1:
                s.step(dt, every)
        e = np.array(series)
        t, tot, tot_ref = e[:, 0], e[:, 1:].sum(1), ref[:, 1:].sum(1)
        early = t <= 1500 * dt + 1e-12
        assert np.max(np.abs(tot[early] - tot_ref[early])) <= 1e-9 * abs(tot_ref[0]), np.max(np.abs(tot[early] - tot_ref[early]))
        exc, exc_ref = np.sqrt(np.mean((tot - tot[0]) ** 2)), np.sqrt(np.mean((tot_ref - tot_ref[0]) ** 2))
        assert abs(exc - exc_ref) <= 0.1 * exc_ref, (mode, exc, exc_ref)
        fit, fit_ref = np.polyfit(t, tot, 1), np.polyfit(t, tot_ref, 1)
        resid = tot_ref - np.polyval(fit_ref, t)
        slope_se = np.std(resid) / (np.std(t) * np.sqrt(len(t)))
        assert abs(fit[0] - fit_ref[0]) <= 3.0 * np.sqrt(2.0) * slope_se, (mode, fit[0], fit_ref[0], slope_se)
        print(f"drift mode {mode}: rms excursion gpu {exc:.4f} ref {exc_ref:.4f} kJ/mol ({100*abs(exc-exc_ref)/exc_ref:.2f}% apart); "
              f"slope gpu {fit[0]:.3e} ref {fit_ref[0]:.3e} (s.e. {slope_se:.1e}) kJ/mol/ps")


def test_energy_drift_512_waters_literal_slope():
    """SURVEY.md section 8c's probe: 512 TIP3P waters in the tether + field potential, 10k steps of 1 fs, mode 0.  The
    survey hoped the literal bar - fitted slope within 10% of the reference's - would be resolvable at this size.  It is
    not: the TRUE reference started from velocities scaled by (1 + 1e-13) (fixture `series_twin`) moves its own slope by
    18% (2.02e-2 -> 1.65e-2 kJ/mol/ps) while its rms excursion moves by 0.1%.  So the test reports the literal slope
    ratio, requires the slope to lie as close to the reference's as the reference's twin does (x2) or within 10%, and
    holds the 10% bar on the drift measure that IS reproducible, the rms excursion of E(t) from E(0)."""
    g = load("drift_water512_mode0")
    s = GpuStepper(g["bodyIndices"], g["masses"], 0)
    s.fused = True
    common.init_like_reference(s, sysd_of(g), tether=True)
    every, dt = int(g["every"]), float(g["dt"])
    ref, twin = g["series"], g["series_twin"]
    series = []
    for i in range(ref.shape[0]):
        U = s.compute_forces()
        ke = s.kinetic()
        series.append([i * every * dt, U, ke[0], ke[1]])
        if i < ref.shape[0] - 1:
            s.step(dt, every)
    e = np.array(series)
    t, tot, tot_ref, tot_twin = e[:, 0], e[:, 1:].sum(1), ref[:, 1:].sum(1), twin[:, 1:].sum(1)
    early = t <= 1000 * dt + 1e-12
    assert np.max(np.abs(tot[early] - tot_ref[early])) <= 1e-9 * abs(tot_ref[0])
    slope, slope_ref, slope_twin = (np.polyfit(t, x, 1)[0] for x in (tot, tot_ref, tot_twin))
    exc, exc_ref = (np.sqrt(np.mean((x - x[0]) ** 2)) for x in (tot, tot_ref))
    print(f"drift 512 waters: slope gpu {slope:.4e} ref {slope_ref:.4e} (ratio {slope/slope_ref:.3f}; reference twin {slope_twin:.4e}, "
          f"ratio {slope_twin/slope_ref:.3f}); rms excursion gpu {exc:.4f} ref {exc_ref:.4f} kJ/mol")
    assert abs(exc - exc_ref) <= 0.1 * exc_ref, (exc, exc_ref)
    assert abs(slope - slope_ref) <= max(0.1 * abs(slope_ref), 2.0 * abs(slope_twin - slope_ref)), (slope, slope_ref, slope_twin)


def test_energy_conservation_integrable_10k_steps():
    """Deterministic companion of the drift test: forces proportional to mass (uniform gravity) exert no
    torque, so every body is a free rotor on a parabola - integrable, no chaotic amplification - and the
    GPU run must reproduce the TRUE reference's trajectory and energy over all 10k steps."""
    for mode in (0, 10):
        g = load(f"gravity_water128_mode{mode}")
        s = GpuStepper(g["bodyIndices"], g["masses"], mode)
        common.init_like_reference(s, sysd_of(g))
        every, dt = int(g["every"]), float(g["dt"])
        ref = g["series"]
        worst = 0.0
        for i in range(ref.shape[0]):
            R, V, _ = s.get_state()
            ke = s.kinetic()
            E = -float(np.sum(g["F"] * R)) + ke[0] + ke[1]
            worst = max(worst, abs(E - ref[i, 1:].sum()))
            if i < ref.shape[0] - 1:
                s.step(dt, every)
        E0 = abs(ref[0, 1:].sum())
        assert worst <= 1e-9 * E0, (mode, worst, E0)
        R, V, _ = s.get_state()
        assert rel_inf(R, g["R_end"]) <= 1e-8 and rel_inf(V, g["V_end"]) <= 1e-8, (rel_inf(R, g["R_end"]), rel_inf(V, g["V_end"]))
        print(f"integrable mode {mode}: max |E_gpu - E_ref| over 10k steps = {worst:.2e} kJ/mol (E0 = {E0:.1f})")


def test_full_size_properties_1M_waters():
    """BASELINE config 2 size: rigid-body invariants, determinism, and a 2,000-molecule random subsample
    stepped by the oracle (bodies are independent under fixed forces, so the subsample must agree)."""
    import torch
    n_mol = 1_000_000
    sysd = common.synth.water_box(n_mol, seed=20240001)
    s = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0, layout="soa")
    common.init_like_reference(s, sysd)
    s.step(0.001, 2)
    R, V, _ = s.get_state()
    assert np.isfinite(R).all() and np.isfinite(V).all()
    Rm = R.reshape(n_mol, 3, 3)
    d_oh1 = np.linalg.norm(Rm[:, 1] - Rm[:, 0], axis=1)
    d_oh2 = np.linalg.norm(Rm[:, 2] - Rm[:, 0], axis=1)
    d_hh = np.linalg.norm(Rm[:, 2] - Rm[:, 1], axis=1)
    assert np.max(np.abs(d_oh1 - common.synth.R_OH)) < 1e-12 and np.max(np.abs(d_oh2 - common.synth.R_OH)) < 1e-12
    assert np.max(np.abs(d_hh - 2 * common.synth.R_OH * np.sin(0.5 * common.synth.ANGLE_HOH))) < 1e-12
    q = s.bodies()["q"]
    assert np.max(np.abs(np.linalg.norm(q, axis=1) - 1.0)) < 1e-14
    # rigid velocities: relative velocity along each bond vanishes
    Vm = V.reshape(n_mol, 3, 3)
    bond = Rm[:, 1] - Rm[:, 0]
    assert np.max(np.abs(np.sum((Vm[:, 1] - Vm[:, 0]) * bond, axis=1))) < 1e-12
    pick = np.sort(np.random.Generator(np.random.Philox(key=3)).choice(n_mol, 2000, replace=False))
    atoms = (pick[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
    sub = {k: sysd[k][atoms] for k in ("masses", "R", "V", "F", "charges")}
    sub["bodyIndices"] = np.repeat(np.arange(1, 2001, dtype=np.int32), 3)
    o = CpuStepper("oracle", sub["bodyIndices"], sub["masses"], 0)
    common.init_like_reference(o, sub)
    o.step(0.001, 2)
    Ro, Vo, _ = o.get_state()
    assert rel_inf(R[atoms], Ro) <= REL_TIGHT and rel_inf(V[atoms], Vo) <= REL_TIGHT
    ke1 = s.kinetic()
    ke2 = s.kinetic()
    assert np.array_equal(ke1, ke2)
    # kinetic energy from the reconstructed atomic velocities must equal KE_t + KE_r (rigid-body identity)
    ke_atoms = 0.5 * float(np.sum(sysd["masses"][:, None] * V * V))
    assert abs(ke_atoms - ke1.sum()) <= 1e-10 * ke_atoms
    del s
    torch.cuda.empty_cache()


def test_execute_host_matches_device_path():
    import torch
    sysd = common.synth.mixed_system(500, 700, seed=5)
    a = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
    common.init_like_reference(a, sysd)
    a.step(0.001, 3)
    Ra, Va, _ = a.get_state()
    b = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
    common.init_like_reference(b, sysd)
    R = torch.from_numpy(sysd["R"].copy()).pin_memory()
    V = torch.from_numpy(sysd["V"].copy()).pin_memory()
    F = torch.from_numpy(sysd["F"].copy()).pin_memory()
    calls = []

    def forces(Rp, Fp, n, user):
        calls.append(n)

    b.sys.execute_host(0.001, 2, R, V, F, forces=forces)
    b.sys.execute_host(0.001, 1, R, V, F)
    assert calls == [len(sysd["masses"])] * 2
    assert np.array_equal(R.numpy(), Ra) and np.array_equal(V.numpy(), Va)


def test_error_behaviour_on_device():
    from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem, RbkError
    import torch
    s = DeviceRigidBodySystem([1, 1, 1, 0], [1.0, 2.0, 3.0, 4.0], 0)
    x = torch.zeros(4, 3, dtype=torch.float64, device="cuda")
    with pytest.raises(RbkError, match="not uploaded"):
        s.part1(0.001, x, x, x)
    with pytest.raises(RbkError, match="not uploaded"):
        s.kinetic(x)
