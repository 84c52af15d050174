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
    ap.add_argument("--layout", default="vec3", choices=["vec3", "soa"])
    ap.add_argument("--no-fuse", action="store_true", help="step with separate part1/part2 launches only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--dt-fs", type=float, default=1.0, help="time step in fs (BASELINE: 1 fs)")
    return ap.parse_args()


def make_workload(args, seed):
    if args.workload == "water":
        sysd = synth.water_box(args.molecules, seed=seed)
        name = f"{args.molecules} rigid TIP3P waters ({3*args.molecules} atoms), mode {args.mode}, integrator-only, fixed synthetic forces (sign alternating per step)"
    else:
        nb = max(args.molecules // 5, 1)
        sysd = synth.mixed_system(nb, int(2.5 * nb), seed=seed)
        name = f"mixed: {nb} rigid bodies of 3-60 atoms + {int(2.5*nb)} free atoms, mode {args.mode}"
    return sysd, name


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own arithmetic on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference(sysd, mode, n_bodies_sample, steps, warmup, threads):
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
        s.set_alternate(True)                  # same workload as the GPU arm: sign of F flips every step
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
    res = cpu_reference(sysd, args.mode, sample, args.steps, args.warmup, cores)
    value = res["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
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
    n = sysd["masses"].shape[0]
    system = DeviceRigidBodySystem(sysd["bodyIndices"], sysd["masses"], args.mode)
    system.update(sysd["R"], np.zeros((n, 3)), sysd["F"], True, True)
    system.update(V=sysd["V"], geometry=False, velocities=True)
    system.upload()
    c = system.counts()
    nB, nF, nA = c["numBodies"], c["numFree"], c["numBodyAtoms"]

    def dev_array(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        return t.t().contiguous() if args.layout == "soa" else t

    pos, vel = dev_array(sysd["R"]), dev_array(sysd["V"])
    # Fixed synthetic forces whose SIGN alternates from step to step (two resident buffers, no extra
    # kernel): constant forces would spin the bodies up without bound (25x thermal angular momentum
    # after 100 steps), alternating ones keep the system at its 300 K state for any number of steps.
    forces = (dev_array(sysd["F"]), dev_array(-sysd["F"]))
    stream = torch.cuda.current_stream()
    cur = [0]                 # index of the force buffer of the most recent force "evaluation"

    def step():
        # Part 1 kicks with the forces of the previous evaluation, Part 2 with the new ones - the order in which
        # the reference's execute() sees them (ReferenceRigidBodyKernels.cpp:97-102)
        system.part1(DT, pos, vel, forces[cur[0]])
        cur[0] ^= 1
        system.part2(DT, pos, vel, forces[cur[0]])

    def barrier():
        torch.cuda.synchronize()
        replicas.barrier(dist if world > 1 else None)
        torch.cuda.synchronize()

    def run_steps(k):
        """k integrator steps the way RigidBodyIntegrator::step(k) runs on this library: part1, then (k-1) times
        [new forces, part2+part1 in one pass (rbk_part2_part1)], then new forces, part2.  --no-fuse: k x [part1, part2]."""
        if args.no_fuse:
            for _ in range(k):
                step()
            return 2 * k
        system.part1(DT, pos, vel, forces[cur[0]])
        for _ in range(k - 1):
            cur[0] ^= 1
            system.part2_part1(DT, pos, vel, forces[cur[0]])
        cur[0] ^= 1
        system.part2(DT, pos, vel, forces[cur[0]])
        return k + 1

    ke_start = system.kinetic(vel)
    clocks = ClockSampler(local)
    clocks.start()
    run_steps(max(args.warmup, 3))
    # ---- timed region: exactly K steps, CUDA events on the launching stream, barrier + sync both sides
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    launches = run_steps(args.steps)
    e1.record(stream)
    barrier()
    ms = replicas.max_over_ranks(e0.elapsed_time(e1), dist if world > 1 else None, dev)
    # ---- per-kernel pass (same workload, same K): events around every launch, for the roofline split
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps)]
    torch.cuda.synchronize()
    for i in range(args.steps):
        ev[3*i].record(stream)
        system.part1(DT, pos, vel, forces[cur[0]])
        ev[3*i+1].record(stream)
        cur[0] ^= 1
        system.part2(DT, pos, vel, forces[cur[0]])
        ev[3*i+2].record(stream)
    torch.cuda.synchronize()
    t1 = float(np.mean([ev[3*i].elapsed_time(ev[3*i+1]) for i in range(args.steps)]))
    t2 = float(np.mean([ev[3*i+1].elapsed_time(ev[3*i+2]) for i in range(args.steps)]))
    fused_ok = (not args.no_fuse) and nB > 0 and nA <= 8 * nB
    tf = None
    if fused_ok:          # event-time the one-pass kernel (part 2 of step k + part 1 of step k+1 = one step of work)
        fe = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        system.part1(DT, pos, vel, forces[cur[0]])
        for i in range(args.steps):
            fe[i].record(stream)
            cur[0] ^= 1
            system.part2_part1(DT, pos, vel, forces[cur[0]])
        fe[args.steps].record(stream)
        cur[0] ^= 1
        system.part2(DT, pos, vel, forces[cur[0]])
        torch.cuda.synchronize()
        tf = float(np.mean([fe[i].elapsed_time(fe[i+1]) for i in range(args.steps)]))
    clk = clocks.stop()
    ke = system.kinetic(vel)
    if not np.isfinite(ke).all():
        raise SystemExit("bench.py: non-finite kinetic energy after the timed region")

    value = world * nB * args.steps / (ms * 1e-3)
    # kernels launched inside the timed region (librbk's launch structure, rbk_kernels.cu launchPart1 / launchPart2 /
    # launchPart2Part1): free atoms have their own launch; large bodies split part 1 into rotation + position kernels
    large, fr = nB > 0 and nA > 8 * nB, 1 if nF > 0 else 0
    per_p1, per_p2 = fr + (2 if large else 1 if nB else 0), fr + (1 if nB else 0)
    per_pp = fr + (3 if large else 1 if nB else 0)
    gpu_launches = args.steps * (per_p1 + per_p2) if args.no_fuse else per_p1 + (args.steps - 1) * per_pp + per_p2
    bytes1 = P1_BODY * nB + P1_ATOM * nA + FREE_P1 * nF
    bytes2 = P2_BODY * nB + P2_ATOM * nA + FREE_P2 * nF
    peak, peak_src = measured_peak()
    dom = ("part1", bytes1, t1) if t1 >= t2 else ("part2", bytes2, t2)
    if tf is not None:
        dom = ("part2Part1", bytes1 + bytes2, tf)
    ach = dom[1] / (dom[2] * 1e-3) / 1e9
    step_ach = (bytes1 + bytes2) / ((ms / args.steps) * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": f"rbk::{'part2Large' if large and dom[0] == 'part2' else dom[0]}Kernel" + (" (+ freeAtomsKernel of the same call)" if fr and tf is None else ""), "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
        "traffic": ncu_traffic(f"{dom[0]}_mode{args.mode}_{args.workload}{args.molecules}"),
        "peak_source": peak_src, "algorithmic_bytes_per_launch": dom[1], "launch_ms": dom[2],
        "kernels": {"part1": {"ms": t1, "bytes": bytes1, "GBps": bytes1 / (t1 * 1e-3) / 1e9},
                    "part2": {"ms": t2, "bytes": bytes2, "GBps": bytes2 / (t2 * 1e-3) / 1e9}},
        "step": {"achieved": step_ach, "frac": step_ach / peak, "bytes": bytes1 + bytes2,
                 "note": "whole step from the 2-event timed region (all launches, incl. the opening part1 / closing part2)"},
    }
    if tf is not None:
        roofline["kernels"]["part2Part1"] = {"ms": tf, "bytes": bytes1 + bytes2, "GBps": (bytes1 + bytes2) / (tf * 1e-3) / 1e9}
        # what the one-pass kernel itself has to move: state read once (r p q pi 1/m 1/I), written once (r p q pi F tau);
        # per atom: force + coordinates + body byte in, velocity + position out
        fused_bytes = (144 + 160) * nB + (49 + 48) * nA + 288 * nF
        roofline["one_pass_compulsory_bytes"] = fused_bytes
        roofline["note"] = ("achieved/frac use SURVEY.md's algorithmic bytes of the two-kernel formulation (560+128n per body-step); the "
                            "one-pass kernel moves fewer compulsory bytes (state resident across the step boundary), so frac can exceed "
                            "what a two-kernel step could reach; frac_of_one_pass_bytes = " + f"{fused_bytes / (tf * 1e-3) / 1e9 / peak:.3f}")

    # ---- e2e: host-buffer call, pinned host R/V/F, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        hR = torch.from_numpy(sysd["R"].copy()).pin_memory()
        hV = torch.from_numpy(sysd["V"].copy()).pin_memory()
        hF = (torch.from_numpy(sysd["F"].copy()).pin_memory(), torch.from_numpy(-sysd["F"]).pin_memory())
        system.upload()                                   # reset body state + device mirrors
        k2 = max(3, min(args.steps, 20))
        for i in range(4):                                # warm-up (first call also uploads the mirrors)
            system.execute_host(DT, 1, hR, hV, hF[i & 1])
        barrier()
        t0 = time.perf_counter()
        for i in range(k2):                               # one call per step, that step's forces from the host
            system.execute_host(DT, 1, hR, hV, hF[i & 1])
        torch.cuda.synchronize()
        el = replicas.max_over_ranks(time.perf_counter() - t0, dist if world > 1 else None, dev)
        e2e = {"value": world * nB * k2 / el, "unit": UNIT, "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": 48 * n,
               "steps": k2, "ms_per_step": 1e3 * el / k2,
               "call": "one rbk_execute_host per step with that step's forces in pinned host memory: forces H2D (copy stream) under part1 + positions D2H, part2, velocities D2H, sync"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = int(min(nB, max(2000, 25000 * cores // (4 if args.mode >= 10 else 1))))
        res = cpu_reference(sysd, args.mode, sample, 40, 2, cores)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": name, "per_gpu": "one independent replica per GPU (replicas only, no collective)",
                       "layout": args.layout, "dt_ps": DT, "bodies": nB, "body_atoms": nA, "free_atoms": nF,
                       "l2": "no flush needed: the per-step working set (state + atoms, >500 MB at 1M waters) exceeds the 126 MB L2"},
            "ns_per_day": (args.steps / (ms * 1e-3)) * DT * 1e-3 * 86400.0,
            "clocks": clk, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu,
            "kinetic_energy_kJmol": {"start": [float(ke_start[0]), float(ke_start[1])], "end": [float(ke[0]), float(ke[1])],
                                     "note": "translational, rotational; the workload stays at its initial ~300 K state"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
