"""The step-fused kernel (rbk_part2_part1 = Part 2 of step k + Part 1 of step k+1 in one pass) against the two separate
kernels, for every system shape.  Where it uses the same segmented scan (bodies larger than 8 atoms somewhere in the
system) or falls back to the two kernels (large mean body size) the results are bit-identical; for small bodies it sums
each body's forces sequentially in the body's own thread (the reference's summation order), which differs from the
scan's tree order by rounding only."""
import numpy as np
import pytest

import common
from common import GpuStepper

pytestmark = pytest.mark.gpu


def run(sysd, mode, fused, layout, shuffle, tether, steps):
    s = GpuStepper(sysd["bodyIndices"], sysd["masses"], mode, layout=layout, shuffle=shuffle)
    s.fused = fused
    common.init_like_reference(s, sysd, tether=tether)
    s.step(0.001, steps)
    R, V, _ = s.get_state()
    return R, V, s.kinetic(), s.bodies()


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("case", ["water", "small_mixed", "medium_mixed", "large_mixed", "free_only"])
def test_fused_is_bit_identical(case, mode):
    if case == "water":
        sysd, layout, shuffle = common.synth.water_box(3000, seed=91), "vec3", False
    elif case == "small_mixed":      # bodies of 3..7 atoms + interleaved free atoms, reordered, SoA layout
        sysd, layout, shuffle = common.synth.mixed_system(2500, 3000, seed=92, max_atoms=7), "soa", True
    elif case == "medium_mixed":     # mean body size <= 8 but some bodies > 8 atoms: fused kernel with the segmented scan
        sysd, layout, shuffle = common.synth.mixed_system(2500, 1000, seed=95, max_atoms=12), "vec3", True
    elif case == "large_mixed":      # mean body size > 8: the fused entry point takes the two-kernel route
        sysd, layout, shuffle = common.synth.mixed_system(400, 500, seed=93, max_atoms=60), "vec3", True
    else:
        n = 700
        rng = np.random.Generator(np.random.Philox(key=94))
        sysd = {"bodyIndices": np.zeros(n, np.int32), "masses": rng.uniform(1, 16, n), "R": rng.uniform(0, 3, (n, 3)),
                "V": rng.standard_normal((n, 3)), "F": rng.standard_normal((n, 3)) * 100, "charges": rng.uniform(-1, 1, n)}
        layout, shuffle = "vec3", False
    for tether in (False, True):
        a = run(sysd, mode, False, layout, shuffle, tether, 6)
        b = run(sysd, mode, True, layout, shuffle, tether, 6)
        if case in ("water", "small_mixed"):          # sequential per-body sums vs tree order: rounding-level differences
            assert common.rel_inf(b[0], a[0]) <= 1e-12 and common.rel_inf(b[1], a[1]) <= 1e-11
            assert common.rel_inf(b[2], a[2]) <= 1e-13
            for k in ("rcm", "pcm", "pi", "force", "torque"):
                assert common.rel_inf(b[3][k], a[3][k]) <= 1e-11, k
            assert common.quat_rel(b[3]["q"], a[3]["q"]) <= 1e-11
            continue
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        for k in ("rcm", "pcm", "q", "pi", "force", "torque"):
            assert np.array_equal(a[3][k], b[3][k]), k


@pytest.mark.parametrize("n_mol", [1003, 33, 5])
@pytest.mark.parametrize("n_free", [0, 7])
def test_one_warp_tiles_ragged_ends_vs_oracle(n_mol, n_free):
    """Water counts that are not a multiple of 32 (partial last one-warp tile; an odd atom count sends that tile down
    the per-thread cp.async route instead of the TMA boxes) and an odd number of leading free atoms (which shifts
    every tile's force range off the 16-byte rule): fused and two-launch stepping against the CPU oracle."""
    from oracle.checkers import CpuStepper
    w = common.synth.water_box(n_mol, seed=77)
    rng = np.random.Generator(np.random.Philox(key=78))
    sysd = {"bodyIndices": np.concatenate([np.zeros(n_free, np.int32), w["bodyIndices"]]),
            "masses": np.concatenate([rng.uniform(1, 16, n_free), w["masses"]]),
            "R": np.vstack([rng.uniform(0, 3, (n_free, 3)), w["R"]]), "V": np.vstack([rng.standard_normal((n_free, 3)), w["V"]]),
            "F": np.vstack([rng.standard_normal((n_free, 3)) * 100, w["F"]])}
    o = CpuStepper("oracle", sysd["bodyIndices"], sysd["masses"], 0)
    common.init_like_reference(o, sysd)
    o.step(0.001, 5)
    Ro, Vo, _ = o.get_state()
    for fused in (False, True):
        g = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
        g.fused = fused
        common.init_like_reference(g, sysd)
        g.step(0.001, 5)
        Rg, Vg, _ = g.get_state()
        assert common.rel_inf(Rg, Ro) < 1e-12 and common.rel_inf(Vg, Vo) < 1e-11, (fused, n_mol, n_free)
        assert common.rel_inf(g.kinetic(), o.kinetic()) < 1e-12


def test_lean_rung0_kernel_and_full_ladder_kernel_give_the_same_bits(monkeypatch):
    """Mode-0 water kernels: the launcher picks the lean order-11 kernel or the full-ladder kernel from a host-side copy
    of the rung that may be a few launches old (csrc/rbk_kernels.cu, launchFusedShape).  Both run the same inlined
    series on rung 0, so which one ran must not show in the results: RBK_FULL_LADDER=1 (never the lean kernel) against
    the default, bit for bit - otherwise runs would not be reproducible."""
    sysd = common.synth.water_box(5000, seed=96)
    sysd = dict(sysd, F=sysd["F"] * 0.05)          # constant forces: keep them small, so that the bodies stay on rung 0
    out = []
    for full in ("0", "1"):
        monkeypatch.setenv("RBK_FULL_LADDER", full)
        for fused in (True, False):
            s = GpuStepper(sysd["bodyIndices"], sysd["masses"], 0)
            s.fused = fused
            common.init_like_reference(s, sysd)
            s.step(0.001, 8)
            R, V, _ = s.get_state()
            assert s.sys.series_order() == 11
            out.append((R, V, s.kinetic()))
            s.close()
    assert np.array_equal(out[0][0], out[2][0]) and np.array_equal(out[0][1], out[2][1]) and np.array_equal(out[0][2], out[2][2])
    assert np.array_equal(out[1][0], out[3][0]) and np.array_equal(out[1][1], out[3][1])
