"""A/B timing of the fused step with another build of librbk (older builds may lack newer diagnostics entry points, which
bench.py needs): python tools/ab_step.py --lib <librbk.so> [--workload mixed] [--molecules N].  Prints ms per step, best of 3 passes."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--lib", required=True)
ap.add_argument("--workload", default="water")
ap.add_argument("--molecules", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=200)
args = ap.parse_args()

import openmm_rigidbody_plugin_b200._lib as L
lib = C.CDLL(args.lib)
for name, (res, a) in L.SIGNATURES.items():
    try:
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, a
    except AttributeError:
        pass
L._lib = lib
import torch
from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem, synth

sysd = synth.water_box(args.molecules, seed=20240001) if args.workload == "water" else synth.mixed_system(args.molecules // 5, int(2.5 * (args.molecules // 5)))
dev = torch.device("cuda:0")
s = DeviceRigidBodySystem(sysd["bodyIndices"], sysd["masses"], 0)
s.update(sysd["R"], sysd["V"], sysd["F"], True, True)
s.upload()
R, V = (torch.from_numpy(sysd[k].copy()).to(dev) for k in ("R", "V"))
F = (torch.from_numpy(sysd["F"].copy()).to(dev), torch.from_numpy(-sysd["F"]).to(dev))
dt = 0.001
s.part1(dt, R, V, F[0])
for i in range(10):
    s.part2_part1(dt, R, V, F[(i + 1) & 1])
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        s.part2_part1(dt, R, V, F[i & 1])
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / args.steps)
print(f"{os.path.relpath(args.lib, ROOT)} {args.workload} {args.molecules}: {best:.4f} ms per fused step")
