#!/usr/bin/env python
"""bench.py - RigidBodyIntegrator step throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (rbk_part1 + rbk_part2, fixed synthetic forces) over the
workload: BASELINE.json configs[1], 1,000,000 rigid TIP3P waters, rotation mode 0 (exact), one
independent replica of that system per GPU ("replicas only": the path has no exchange step, so
N GPUs = N replicas, no collective on the data path; scaling "weak").

Printed (rank 0, ONE JSON line):
  value     whole-job body-steps/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       the same metric through the host-buffer call rbk_execute_host (pinned host R/V/F;
            per step: forces H2D, positions + velocities D2H inside the timed region)
  roofline  dominant kernel: algorithmic bytes per launch / CUDA-event launch duration vs the measured
            HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the reference's CPU arithmetic timed on this box's host cores on a bounded sample
--impl reference times only that CPU arm (oracle/_ref = the unmodified reference sources when the
prebuilt library is present, else the oracle port) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from openmm_rigidbody_plugin_b200 import replicas, synth  # noqa: E402

DT = 0.001                      # ps (1 fs, README.md:200 of the reference)
METRIC = "integrator body-steps/s at 1M rigid waters (mode 0 exact rotation, fixed synthetic forces)"


def metric_name(args):
    """BASELINE.json's metric for the workload it is quoted on (the defaults); other workloads say what they are."""
    if args.workload == "water" and args.molecules == 1_000_000 and args.mode == 0:
        return METRIC
    what = f"{args.molecules} rigid waters" if args.workload == "water" else "the mixed workload (BASELINE config 4 shape)"
    return f"integrator body-steps/s at {what} (mode {args.mode}, fixed synthetic forces)"
UNIT = "body-steps/s"
# algorithmic bytes per body-step, split per kernel
P1_BODY, P1_ATOM, P2_BODY, P2_ATOM = 320, 52, 240, 76
FREE_P1, FREE_P2 = 156, 132


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--molecules", type=int, default=1_000_000)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--workload", default="water", choices=["water", "mixed"])
    ap.add_argument("--layout", default="vec3", choices=["vec3", "soa", "openmm-mixed", "openmm-double"],
                    help="caller-owned atom arrays: fp64 Vec3 / SoA planes, or the OpenMM-CUDA boundary formats")
    ap.add_argument("--shuffle", nargs="?", const="molecules", default=None, choices=["molecules", "atoms"],
                    help="caller's atom order differs from the plugin's (rbk_set_atom_location): 'molecules' = whole bodies / free atoms "
                         "permuted as units, what OpenMM's reorderAtoms does; 'atoms' = every atom on its own (worst case)")
    ap.add_argument("--forces", default="alternating", choices=["alternating", "constant"],
                    help="fixed synthetic forces: sign flipping every step (default) or literally constant (SURVEY 8d)")
    ap.add_argument("--no-parity", action="store_true", help="skip the CPU-oracle subsample check after the timed region")
    ap.add_argument("--no-fuse", action="store_true", help="step with separate part1/part2 launches only")
    ap.add_argument("--graph", action="store_true",
                    help="replay the interior steps of step(K) from a CUDA graph (two fused steps captured, K/2 - 1 replays): what a caller "
                         "with a fixed launch sequence can do about launch gaps; the step calls are capturable (rbk.h)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip timing the reference's own CUDA kernels (baseline/ref_cuda)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--dt-fs", type=float, default=1.0, help="time step in fs (BASELINE: 1 fs)")
    ap.add_argument("--free-per-body", type=float, default=2.5, help="mixed workload: free atoms per rigid body (BASELINE config 4: 2.5)")
    return ap.parse_args()


def make_workload(args, seed):
    if args.workload == "water":
        sysd = synth.water_box(args.molecules, seed=seed)
        name = f"{args.molecules} rigid TIP3P waters ({3*args.molecules} atoms), mode {args.mode}, integrator-only, fixed synthetic forces ({'sign alternating per step' if args.forces == 'alternating' else 'constant'})"
    else:
        nb = max(args.molecules // 5, 1)
        nf = int(args.free_per_body * nb)
        sysd = synth.mixed_system(nb, nf, seed=seed)
        name = f"mixed: {nb} rigid bodies of 3-60 atoms + {nf} free atoms, mode {args.mode}"
    return sysd, name


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own arithmetic on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference(sysd, mode, n_bodies_sample, steps, warmup, threads, alternate=True):
    """Time `steps` passes of the reference CPU stepper over the first n_bodies_sample molecules of the
    workload, bodies split into `threads` disjoint slices (the reference itself is single-threaded:
    plain loops in RigidBodySystem.cpp:170-204; bodies are independent, so slicing is exact)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    from oracle import checkers
    kind = "reference" if checkers.available("reference") else "oracle"
    body = sysd["bodyIndices"]
    # water: 3 atoms per molecule, contiguous; general: cut at body boundaries of the label array
    labels = np.unique(body[body > 0])
    n_bodies_sample = min(n_bodies_sample, labels.shape[0])
    threads = max(1, min(threads, n_bodies_sample))
    cut = np.linspace(0, n_bodies_sample, threads + 1).astype(np.int64)
    steppers = []
    for t in range(threads):
        lo, hi = labels[cut[t]], labels[cut[t + 1] - 1]
        sel = np.nonzero((body >= lo) & (body <= hi))[0]
        sub = {k: np.ascontiguousarray(sysd[k][sel]) for k in ("masses", "R", "V", "F", "charges", "bodyIndices")}
        s = checkers.CpuStepper(kind, sub["bodyIndices"], sub["masses"], mode)
        common.init_like_reference(s, sub)
        s.set_alternate(alternate)             # same workload as the GPU arm (sign of F flips every step unless --forces constant)
        steppers.append(s)

    def run(n):
        ths = [threading.Thread(target=s.step, args=(DT, n)) for s in steppers]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return time.perf_counter() - t0

    if warmup:
        run(warmup)
    elapsed = run(steps)
    for s in steppers:
        s.close()
    return {
        "value": n_bodies_sample * steps / elapsed, "unit": UNIT, "cores": threads, "kind": "reference" if kind == "reference" else "port",
        "sample": f"first {n_bodies_sample} bodies of the workload x {steps} steps, {threads} threads on disjoint body slices "
                  f"({'unmodified reference sources, oracle/_ref' if kind == 'reference' else 'oracle/rb_oracle.c port'}, g++ -O2)",
        "seconds": elapsed,
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sysd, name = make_workload(args, seed=20240001)
    cores = os.cpu_count() or 1
    # bounded sample: ~0.1 s of work per step per core
    sample = int(min(args.molecules, max(2000, 40000 * cores // (10 if args.mode >= 10 else 1) // 4)))
    res = cpu_reference(sysd, args.mode, sample, args.steps, args.warmup, cores, args.forces == "alternating")
    value = res["value"]
    line = {
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "note": "CPU arm: each step is one pass over a bounded sample of the workload (body-steps/s is size-independent on the CPU)"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "ns_per_day": None,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks: sample NVML during the measurement
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self._th = threading.Thread(target=self._loop, daemon=True)
            self._th.start()

    def stop(self):
        self._stop.set()
        if self._th is not None:
            self._th.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "how": "NVML polled every 5 ms from the first warm-up step to the end of the per-kernel timing pass"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "of fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel_key):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if one exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


class AtomArrays:
    """The caller-owned device arrays of one replica in one of the layouts librbk accepts, with the step calls that go
    with it.  vec3 / soa: plain fp64 (RBK_LAYOUT_VEC3 = the Reference platform's std::vector<Vec3>, RBK_LAYOUT_SOA);
    openmm-mixed / openmm-double: the OpenMM-CUDA boundary formats (posq float4 + posqCorrection | double4, velm double4,
    fixed-point long long force planes).  shuffle: atoms stored in a random permutation handed to rbk_set_atom_location
    (the CUDA platform's atom reordering)."""

    def __init__(self, system, sysd, layout, shuffle, dev, alternate):
        import torch
        from openmm_rigidbody_plugin_b200._lib import RBK_OPENMM_DOUBLE, RBK_OPENMM_MIXED
        self.t, self.sys, self.layout, self.dev = torch, system, layout, dev
        n = self.n = sysd["masses"].shape[0]
        self.order = None
        if shuffle and system is not None:
            rng = np.random.Generator(np.random.Philox(key=12345))
            if shuffle == "atoms":
                self.order = rng.permutation(n)                   # atom i lives at slot order[i]
            else:
                # units = runs of consecutive atoms with the same positive body label, and single free atoms
                body = sysd["bodyIndices"]
                start = np.nonzero(np.concatenate([[True], (body[1:] != body[:-1]) | (body[1:] <= 0)]))[0]
                length = np.diff(np.concatenate([start, [n]]))
                perm = rng.permutation(start.shape[0])            # unit perm[k] is stored k-th
                new_start = np.empty_like(start)
                new_start[perm] = np.concatenate([[0], np.cumsum(length[perm])[:-1]])
                self.order = (np.repeat(new_start - start, length) + np.arange(n)).astype(np.int64)
            system.set_atom_location(self.order[system.atom_index()].astype(np.int32))
        self.openmm = layout.startswith("openmm")
        F = sysd["F"]
        if self.openmm:
            self.precision = RBK_OPENMM_MIXED if layout == "openmm-mixed" else RBK_OPENMM_DOUBLE
            self.padded = ((n + 31) // 32) * 32
            Fq = np.round(F * 4294967296.0).astype(np.int64)         # exact: run_b200_arm quantised sysd["F"] already
            pos4, vel4 = np.zeros((self.padded, 4)), np.zeros((self.padded, 4))
            sl = self.order if self.order is not None else np.arange(n)
            pos4[sl, :3], pos4[sl, 3] = sysd["R"], sysd["charges"]
            vel4[sl, :3], vel4[sl, 3] = sysd["V"], 1.0 / sysd["masses"]
            p64 = torch.from_numpy(pos4).to(dev)
            if self.precision == RBK_OPENMM_MIXED:
                self.posq = p64.float().contiguous()
                self.corr = (p64 - self.posq.double()).float().contiguous()
            else:
                self.posq, self.corr = p64.contiguous(), None
            self.velm = torch.from_numpy(vel4).to(dev).contiguous()
            planes = np.zeros((3, self.padded), np.int64)
            planes[:, sl] = Fq.T
            f0 = torch.from_numpy(planes).to(dev).contiguous()
            self.forces = (f0, (-f0).contiguous()) if alternate else (f0, f0)
        else:
            self.pos, self.vel = self._dev(sysd["R"]), self._dev(sysd["V"])
            f0 = self._dev(F)
            self.forces = (f0, self._dev(-F)) if alternate else (f0, f0)

    def _dev(self, a):
        if self.order is not None:
            b = np.empty_like(a)
            b[self.order] = a
            a = b
        x = self.t.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        return x.t().contiguous() if self.layout == "soa" else x

    def bind(self, dt, stream):
        """Prebuild the ctypes argument tuples of the three step calls (per force buffer), so that a step costs one foreign
        call - the Python wrapper's per-call work (shape checks, data_ptr, current_stream) is measurable next to a 35 us step."""
        import ctypes as C
        from openmm_rigidbody_plugin_b200 import _lib
        lib, h, st = self.sys.lib, self.sys.h, C.c_void_p(stream.cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None      # noqa: E731
        self._check = _lib.check
        self._calls = {}
        for name in ("part1", "part2", "part2_part1"):
            per_buffer = []
            for f in self.forces:
                if self.openmm:
                    fn = getattr(lib, f"rbk_{name}_openmm")
                    args = (h, C.c_double(dt), p(self.posq), p(self.corr), p(self.velm), p(f), self.padded, self.precision, st)
                else:
                    fn = getattr(lib, f"rbk_{name}")
                    lay, stride = (1, self.pos.shape[1]) if self.layout == "soa" else (0, 0)
                    args = (h, C.c_double(dt), p(self.pos), p(self.vel), p(f), lay, C.c_longlong(stride), st)
                per_buffer.append((fn, args))
            self._calls[name] = per_buffer

    def _call(self, name, k):
        fn, args = self._calls[name][k]
        rc = fn(*args)
        if rc:
            self._check(rc)

    def part1(self, dt, k):
        self._call("part1", k)

    def part2(self, dt, k):
        self._call("part2", k)

    def part2_part1(self, dt, k):
        self._call("part2_part1", k)

    def kinetic(self):
        return self.sys.kinetic_openmm(self.velm, self.precision) if self.openmm else self.sys.kinetic(self.vel)

    def rows(self, atoms):
        """positions and velocities (fp64, host) of the given atoms (original numbering)"""
        t = self.t
        idx = t.from_numpy(np.ascontiguousarray(self.order[atoms] if self.order is not None else atoms)).to(self.dev)
        if self.openmm:
            R = self.posq[idx, :3].double()
            if self.corr is not None:
                R = R + self.corr[idx, :3].double()
            V = self.velm[idx, :3]
        elif self.layout == "soa":
            R, V = self.pos[:, idx].t(), self.vel[:, idx].t()
        else:
            R, V = self.pos[idx], self.vel[idx]
        return R.cpu().numpy(), V.cpu().numpy()

    # bytes moved per atom by the position write / velocity write / force read in this layout
    def atom_io_bytes(self):
        if not self.openmm:
            return 24, 24, 24
        return (32 if self.corr is not None else 32), 32, 24


def parity_subsample(sysd, arrays, mode, total_steps, alternate, n_bodies=2000, n_free=2000):
    """The timed kernels did the work: re-run a random subsample of the workload (bodies and free atoms are independent
    under prescribed forces) on the CPU oracle for exactly the number of steps the GPU arrays have seen and compare."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    from oracle import checkers
    body = sysd["bodyIndices"]
    rng = np.random.Generator(np.random.Philox(key=4242))
    labels = np.unique(body[body > 0])
    pick = rng.choice(labels, min(n_bodies, labels.shape[0]), replace=False) if labels.shape[0] else labels
    free = np.nonzero(body <= 0)[0]
    mask = np.isin(body, pick)
    if free.shape[0]:
        mask[rng.choice(free, min(n_free, free.shape[0]), replace=False)] = True
    atoms = np.nonzero(mask)[0]
    sub = {k: np.ascontiguousarray(sysd[k][atoms]) for k in ("masses", "R", "V", "F", "charges", "bodyIndices")}
    o = checkers.CpuStepper("oracle", sub["bodyIndices"], sub["masses"], mode)
    common.init_like_reference(o, sub)
    o.set_alternate(alternate)
    o.step(DT, total_steps)
    Ro, Vo, _ = o.get_state()
    Rg, Vg = arrays.rows(atoms)
    o.close()

    def quantiles(a, b):
        err = np.max(np.abs(a - b), axis=1) / float(np.max(np.abs(b)))
        return float(np.median(err)), float(np.quantile(err, 0.99)), float(np.max(err))
    (r50, r99, rmax), (v50, v99, vmax) = quantiles(Rg, Ro), quantiles(Vg, Vo)
    return {"max_rel_R": rmax, "max_rel_V": vmax, "p99_rel_R": r99, "p99_rel_V": v99, "median_rel_R": r50, "median_rel_V": v50,
            "steps": total_steps, "bodies": int(pick.shape[0]), "atoms": int(atoms.shape[0]),
            "how": "random subsample re-run on the CPU oracle (oracle/rb_oracle.c) for every step the device arrays have seen; per atom "
                   "|gpu - cpu|_inf / ||cpu||_inf, max / 99th percentile / median over the sample.  The bar (1e-6, BASELINE.json's for ONE "
                   "step) is applied to the 99th percentile after all steps: over hundreds of steps the few rotors that pass near the "
                   "unstable intermediate-axis rotation amplify the rounding difference between two exact-rotation algorithms "
                   "exponentially, and they set the max at long time steps"}


def gpu_reference(system, sysd, mode, steps, alternate, dev):
    """The kernels this library replaces, on the same GPU and workload: the reference's own CUDA kernels
    (platforms/cuda/src/kernels/rigidbodyintegrator.cu, compiled in place for sm_100a by baseline/ref_cuda/Makefile) driven as
    CudaIntegrateRigidBodyStepKernel::execute drives them - integrateRigidBodyPart1, integrateRigidBodyPart2, blocks of 128
    threads - on OpenMM-format arrays (mixed and double precision).  Bench infrastructure, like cpu_baseline."""
    try:
        from baseline.ref_cuda import refcuda
    except Exception:
        return None
    out = {}
    n = sysd["masses"].shape[0]
    quant = dict(sysd, F=np.round(sysd["F"] * 4294967296.0) / 4294967296.0)
    for precision in ("mixed", "double"):
        if not refcuda.available(precision, mode, 0):
            continue
        A = AtomArrays(None, quant, "openmm-" + precision, None, dev, alternate)
        best = None
        for blocks in (6, 12, 24):                           # OpenMM's grid cap is 6 blocks per SM; larger grids tried in the reference's favour
            ref = refcuda.RefCudaSystem(precision, mode, 0, system.host_bodies(), system.body_fixed(), system.atom_index(),
                                        system.counts()["numFree"], A.padded)
            ref.blocks_per_sm = blocks
            B = AtomArrays(None, quant, "openmm-" + precision, None, dev, alternate)
            _, cur = ref.time_steps(DT, 3, B.posq, B.corr, B.velm, B.forces[0], B.forces[1], 0)
            ms, _ = ref.time_steps(DT, steps, B.posq, B.corr, B.velm, B.forces[0], B.forces[1], cur)
            if best is None or ms < best[0]:
                best = (ms, blocks)
            ref.close()
        nB = system.counts()["numBodies"]
        out[precision] = {"value": nB * steps / (best[0] * 1e-3), "unit": UNIT, "ms_per_step": best[0] / steps, "blocks_per_sm": best[1], "steps": steps}
    if not out:
        return None
    out["kernels"] = ("integrateRigidBodyPart1 + integrateRigidBodyPart2 of the reference (one thread per body, AoS BodyData), "
                      "sources compiled in place from /root/reference, CUDA-event timed, inputs resident")
    return out


def committed_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return None


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem

    rank, world, local = replicas.rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the rigid-body step has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own log lines ("NCCL version ..." at NCCL_DEBUG=VERSION/WARN) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    # one independent replica per GPU, distinct seed per replica (BASELINE.json: replicas only)
    sysd, name = make_workload(args, seed=replicas.replica_seed(20240001, rank))
    alternate = args.forces == "alternating"
    n = sysd["masses"].shape[0]
    if args.layout.startswith("openmm"):
        # OpenMM's force arrays are fixed point (scale 2^32): quantise once so that the body build, the kernels and the CPU
        # check all see the same numbers
        sysd = dict(sysd, F=np.round(sysd["F"] * 4294967296.0) / 4294967296.0)
    system = DeviceRigidBodySystem(sysd["bodyIndices"], sysd["masses"], args.mode)
    system.update(sysd["R"], np.zeros((n, 3)), sysd["F"], True, True)
    system.update(V=sysd["V"], geometry=False, velocities=True)
    system.upload()
    c = system.counts()
    nB, nF, nA = c["numBodies"], c["numFree"], c["numBodyAtoms"]
    A = AtomArrays(system, sysd, args.layout, args.shuffle, dev, alternate)
    stream = torch.cuda.current_stream()
    A.bind(DT, stream)
    cur = [0]                 # index of the force buffer of the most recent force "evaluation"
    done = [0]                # integrator steps the device arrays have seen
    graphs = {}
    capture_stream = torch.cuda.Stream() if args.graph else None

    # Forces: fixed synthetic arrays.  --forces alternating (default): the SIGN flips from step to step (two resident
    # buffers, no extra kernel) - constant forces spin the bodies up without bound (25x thermal angular momentum after
    # 100 steps), alternating ones keep the system at its 300 K state for any number of steps.  --forces constant: the
    # literal contract workload of SURVEY.md section 8(d).
    def barrier():
        torch.cuda.synchronize()
        replicas.barrier(dist if world > 1 else None)
        torch.cuda.synchronize()

    def run_steps(k, marks=None):
        """k integrator steps the way RigidBodyIntegrator::step(k) runs on this library: part1, then (k-1) times
        [new forces, part2+part1 in one pass (rbk_part2_part1)], then new forces, part2.  --no-fuse: k x [part1, part2].
        Part 1 kicks with the forces of the previous evaluation, Part 2 with the new ones - the order in which the
        reference's execute() sees them (ReferenceRigidBodyKernels.cpp:97-102)."""
        done[0] += k
        if args.no_fuse:
            for _ in range(k):
                A.part1(DT, cur[0])
                cur[0] ^= 1
                A.part2(DT, cur[0])
            return 2 * k
        A.part1(DT, cur[0])
        if marks is not None:
            marks[0].record(stream)                      # (inside the timed region: the k - 1 one-pass launches lie between the marks)
        interior = k - 1
        first = cur[0] ^ 1                               # force buffer of the first interior step
        if args.graph and first in graphs:
            for _ in range(interior // 2):
                graphs[first].replay()                   # = two fused steps: buffer `first`, then the other one
            interior -= 2 * (interior // 2)
        for _ in range(interior):
            cur[0] ^= 1
            A.part2_part1(DT, cur[0])
        if marks is not None:
            marks[1].record(stream)
        cur[0] ^= 1
        A.part2(DT, cur[0])
        if args.graph and not graphs:
            # after the first (warm-up) call - kernel attributes set, side stream created, series ladder settled - capture
            # the two-step graphs for either buffer parity; capturing runs nothing and stays outside the timed region
            torch.cuda.synchronize()
            A.bind(DT, capture_stream)
            for f0 in (0, 1):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=capture_stream):
                    A.part2_part1(DT, f0)
                    A.part2_part1(DT, f0 ^ 1)
                graphs[f0] = g
            A.bind(DT, stream)
        return k + 1

    ke_start = A.kinetic()
    clocks = ClockSampler(local)
    clocks.start()
    run_steps(max(args.warmup, 3))
    # ---- timed region: exactly K steps, CUDA events on the launching stream, barrier + sync both sides
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    marks = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    e0.record(stream)
    launches = run_steps(args.steps, marks)
    e1.record(stream)
    barrier()
    ms = replicas.max_over_ranks(e0.elapsed_time(e1), dist if world > 1 else None, dev)
    # ---- per-kernel pass (same workload, same K): events around every launch, for the roofline split
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps)]
    torch.cuda.synchronize()
    for i in range(args.steps):
        ev[3*i].record(stream)
        A.part1(DT, cur[0])
        ev[3*i+1].record(stream)
        cur[0] ^= 1
        A.part2(DT, cur[0])
        ev[3*i+2].record(stream)
    done[0] += args.steps
    torch.cuda.synchronize()
    t1 = float(np.mean([ev[3*i].elapsed_time(ev[3*i+1]) for i in range(args.steps)]))
    t2 = float(np.mean([ev[3*i+1].elapsed_time(ev[3*i+2]) for i in range(args.steps)]))
    large = nB > 0 and nA > 8 * nB
    fused_ok = (not args.no_fuse) and nB > 0 and not large
    tf = None
    if not args.no_fuse:  # event-time the one-pass call (part 2 of step k + part 1 of step k+1 = one step of work):
        # ONE event pair around K back-to-back launches (an event between every two launches costs the GPU ~8 us of idle
        # time per launch at this launch length, which is not the kernel's)
        # Three such passes, the fastest one counts (the host has to stay ahead of a 0.1 ms kernel; one pass in a few is
        # caught by scheduler jitter on a busy box and then measures the host).
        tf = None
        for _ in range(3):
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            A.part1(DT, cur[0])
            f0.record(stream)
            for i in range(args.steps):
                cur[0] ^= 1
                A.part2_part1(DT, cur[0])
            f1.record(stream)
            cur[0] ^= 1
            A.part2(DT, cur[0])
            done[0] += args.steps + 1
            torch.cuda.synchronize()
            t = f0.elapsed_time(f1) / args.steps
            if os.environ.get("BENCH_DEBUG_TF"):
                print(f"[bench] one-pass launch pass: {t:.4f} ms per launch, series order {system.series_order()}", file=sys.stderr)
            tf = t if tf is None else min(tf, t)
    # The one-pass launch as timed INSIDE the timed region (the marks around its K - 1 interior launches): the roofline's launch
    # duration, under the clocks `value` was measured at.  The separate passes above run after >= 60 ms of sustained fp64 load,
    # by which time a B200 may sit at its power cap (sw_power_cap, SM clock 1740 instead of 1965 MHz: +4..10 % per launch);
    # they are kept as `launch_ms_later_pass`.
    tf_later = tf
    if tf is not None and args.steps > 1:
        tf = marks[0].elapsed_time(marks[1]) / (args.steps - 1)
    clk = clocks.stop()
    ke = A.kinetic()
    if not np.isfinite(ke).all():
        raise SystemExit("bench.py: non-finite kinetic energy after the timed region")
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_subsample(sysd, A, args.mode, done[0], alternate)
        # BASELINE.json's bar for ONE step (1e-6), held on the 99th percentile after every step run here - up to ~1,200 steps.
        # Beyond that the driven rotors' chaos takes over (each water is a rigid rotor under a torque that depends on its
        # orientation: rounding differences between two exact-rotation algorithms grow exponentially for a growing share of
        # the bodies - after 15,000 steps the median is still 2e-13 but 1 % of the bodies have decorrelated), so long runs are
        # judged on the median, which stays at rounding level as long as the kernels do their work.
        bar = 1e-6
        if parity["steps"] <= 1200:
            parity["ok"] = bool(parity["p99_rel_R"] <= bar and parity["p99_rel_V"] <= bar)
        else:
            parity["ok"] = bool(parity["median_rel_R"] <= 1e-9 and parity["median_rel_V"] <= 1e-7)
        if not parity["ok"] and args.forces == "alternating" and DT <= 0.0021:
            raise SystemExit(f"bench.py: the timed kernels disagree with the CPU oracle: {parity}")

    value = world * nB * args.steps / (ms * 1e-3)
    # kernels launched inside the timed region: the library's own count per call (rbk_debug_launches_per_call; launch
    # structure in rbk_kernels.cu launchPart1 / launchPart2 / launchPart2Part1: large bodies split part 1 into rotation +
    # position kernels, free atoms have their own launch unless they ride along in the large-body Part 2 kernel)
    per_p1, per_p2, per_pp = system.launches_per_call()
    gpu_launches = args.steps * (per_p1 + per_p2) if args.no_fuse else per_p1 + (args.steps - 1) * per_pp + per_p2

    # ---- roofline.  Two byte counts per launch, both stated:
    #   compulsory  what the kernel(s) ACTUALLY timed must move (this layout, this launch structure) - `frac` uses it
    #   work-equivalent  SURVEY.md section 8(d)'s model of the two-kernel formulation (560 + 128 n per body-step, 288 per
    #                free atom-step) - the contract's per-unit figure, kept under work_equiv_*; the one-pass kernel moves
    #                fewer bytes than that model, so work_equiv_frac can exceed 1 and is NOT a bandwidth
    wpos, wvel, rfor = A.atom_io_bytes()
    identity = A.order is None and not (nF > 0 and nB > 0 and args.workload == "mixed")
    loc4 = 0 if identity else 4                      # atomLoc look-up when the plugin-order -> array map is not the identity
    # per body: state planes read + written (8 B each) + 4 B prefix offset; per atom: body-frame coordinates 24, body byte 1,
    # [slot 4], force, velocity, position in this layout; per free atom: v, f, x, 1/m (+ savedPos) in, v, x (+ savedPos) out
    p1_bytes = (192 + 4 + 112) * nB + (24 + 1 + loc4 + wpos) * nA + (loc4 + 2 * wvel + rfor + 2 * wpos + 8 + 24) * nF
    p2_bytes = (120 + 4 + 104) * nB + (24 + 1 + (loc4 if (not large or args.shuffle == "atoms") else 0) + rfor + wvel) * nA + (loc4 + 2 * wvel + rfor + wpos + 24 + 8) * nF
    free_pp = (loc4 + 2 * wvel + rfor + 2 * wpos + 24 + 8 + 24) * nF
    if large:
        # one rbk_part2_part1 call = part2LargeKernel (state in 120 + 4, out 104; per atom d, byte, slot, f in, v out) +
        # rotation kernel (in 192, out 112) + atomPositionKernel (r, q in 56; per atom d, byte, slot in, x out) + free atoms
        # (bodies that are runs of slots - everything but --shuffle atoms: part2LargeKernel reads one slot per body, not per atom)
        loc4_p2 = loc4 if args.shuffle == "atoms" else 0
        pp_bytes = (124 + 104 + 192 + 112 + 56) * nB + (2 * (24 + 1) + loc4 + loc4_p2 + rfor + wvel + wpos) * nA + free_pp
    else:
        # the one-pass kernel: r p q pi 1/m 1/I in (144 + 4), r p q pi out (112) [+ F tau out (48) when the stores are kept]
        ft = 0 if system_lazy_ft() else 48
        pp_bytes = (148 + 112 + ft) * nB + (24 + 1 + loc4 + rfor + wvel + wpos) * nA + free_pp
    work1 = P1_BODY * nB + P1_ATOM * nA + FREE_P1 * nF
    work2 = P2_BODY * nB + P2_ATOM * nA + FREE_P2 * nF
    peak, peak_src = measured_peak()
    if tf is not None:
        dom = ("part2Part1", pp_bytes, tf)
    else:
        dom = ("part1", p1_bytes, t1) if t1 >= t2 else ("part2", p2_bytes, t2)
    ach = dom[1] / (dom[2] * 1e-3) / 1e9
    kernel_name = {"part2Part1": "rbk::part2Part1Kernel" if not large else "rbk_part2_part1 call = part2LargeKernel (bodies' Part 2 + the free atoms' Part 2 and Part 1 riding along) + part1Kernel (rotation) + atomPositionKernel",
                   "part1": "rbk_part1 call", "part2": "rbk_part2 call"}[dom[0]]
    traffic = ncu_traffic(f"{dom[0]}_mode{args.mode}_{args.workload}{args.molecules}_{args.layout}{'_shuffle_' + args.shuffle if args.shuffle else ''}")
    work_ach = (work1 + work2) / (dom[2] * 1e-3) / 1e9 if tf is not None else None
    fp64 = committed_json("fp64_peak.json")
    roofline = {
        "bound": "hbm", "kernel": kernel_name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
        "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom[1], "launch_ms": dom[2],
        "launch_ms_how": ("events inside the timed region around its K - 1 one-pass launches" if tf is not None and args.steps > 1 else
                          "mean over the per-kernel pass (an event between every two launches)"),
        "launch_ms_later_pass": tf_later,
        "bytes_model": "compulsory bytes of the launch(es) actually timed: per body state in + out, per atom body-frame coordinates, "
                       "body byte, force, velocity, position in this layout; DESIGN.md section 4 lists them per kernel",
        "kernels": {"part1": {"ms": t1, "bytes": p1_bytes, "GBps": p1_bytes / (t1 * 1e-3) / 1e9},
                    "part2": {"ms": t2, "bytes": p2_bytes, "GBps": p2_bytes / (t2 * 1e-3) / 1e9}},
        "work_equiv": {"bytes_per_step": work1 + work2, "achieved": work_ach, "frac": None if work_ach is None else work_ach / peak,
                       "note": "SURVEY.md 8(d) two-kernel byte model (560 + 128 n per body-step) over the one-pass launch time: a work rate, "
                               "not a bandwidth (can exceed 1)"},
        "step": {"ms": ms / args.steps, "compulsory_GBps": (pp_bytes if tf is not None else p1_bytes + p2_bytes) / ((ms / args.steps) * 1e-3) / 1e9,
                 "note": "whole step from the 2-event timed region (all launches, incl. the opening part1 / closing part2)"},
        "secondary_bound_fp64": fp64,
    }
    if tf is not None:
        roofline["kernels"]["part2Part1"] = {"ms": tf, "bytes": pp_bytes, "GBps": pp_bytes / (tf * 1e-3) / 1e9}
    if traffic:
        roofline["traffic_frac"] = traffic / (dom[2] * 1e-3) / 1e9 / peak

    # ---- e2e: the host-buffer call (what a Reference-platform RigidBodyIntegrator::step drives): pinned host R/V/F; every
    # step that step's forces go host -> device and the new positions device -> host inside the timed region; velocities
    # come back once, after the last step (nothing on the host reads them in between - rbk.h, rbk_execute_host)
    e2e = None
    if not args.no_e2e and not A.openmm and A.order is None and args.layout == "vec3":
        hR = torch.from_numpy(sysd["R"].copy()).pin_memory()
        hV = torch.from_numpy(sysd["V"].copy()).pin_memory()
        hF = (torch.from_numpy(sysd["F"].copy()).pin_memory(), torch.from_numpy((-sysd["F"]) if alternate else sysd["F"].copy()).pin_memory())
        system.upload()                                   # reset body state + device mirrors
        k2 = max(3, min(args.steps, 20))
        system.execute_host(DT, 1, hR, hV, hF[0])         # warm-up (the first call also uploads the mirrors)
        for i in range(1, 4):
            system.execute_host(DT, 1, hR, None, hF[i & 1])
        barrier()
        t0 = time.perf_counter()
        for i in range(k2):                               # one call per step, that step's forces from the host
            system.execute_host(DT, 1, hR, hV if i == k2 - 1 else None, hF[i & 1])
        torch.cuda.synchronize()
        el = replicas.max_over_ranks(time.perf_counter() - t0, dist if world > 1 else None, dev)
        e2e = {"value": world * nB * k2 / el, "unit": UNIT, "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": 24 * n + 24 * n // k2,
               "steps": k2, "ms_per_step": 1e3 * el / k2,
               "call": "one rbk_execute_host per step with that step's forces in pinned host memory: forces H2D (copy stream) under part1 + "
                       "positions D2H, part2, sync; velocities D2H once, after the last step"}

    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_gpu_reference and args.workload == "water":
        system.upload()                                   # host model = the initial state; the reference starts from it too
        gpu_ref = gpu_reference(system, sysd, args.mode, max(3, min(args.steps, 30)), alternate, dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = int(min(nB, max(2000, 25000 * cores // (4 if args.mode >= 10 else 1))))
        res = cpu_reference(sysd, args.mode, sample, 40, 2, cores, alternate)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": name, "per_gpu": "one independent replica per GPU (replicas only, no collective)",
                       "layout": args.layout, "shuffle": args.shuffle, "cuda_graph": bool(args.graph), "series_order": system.series_order(), "forces": args.forces, "dt_ps": DT, "bodies": nB, "body_atoms": nA, "free_atoms": nF,
                       "l2": "no flush needed: the per-step working set (state + atoms, >500 MB at 1M waters) exceeds the 126 MB L2"},
            "ns_per_day": (args.steps / (ms * 1e-3)) * DT * 1e-3 * 86400.0,
            "clocks": clk, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu, "gpu_reference": gpu_ref,
            "parity_subsample": parity,
            "kinetic_energy_kJmol": {"start": [float(ke_start[0]), float(ke_start[1])], "end": [float(ke[0]), float(ke[1])],
                                     "note": "translational, rotational"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def system_lazy_ft():
    return os.environ.get("RBK_EAGER_FORCE_TORQUE", "0")[:1] != "1"


def main():
    global DT
    args = parse_args()
    DT = args.dt_fs * 1e-3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
