"""GPU diagnostic: series-ladder rung and launch time, launch by launch (1M waters)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from openmm_rigidbody_plugin_b200 import DeviceRigidBodySystem, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dt = float(sys.argv[2]) if len(sys.argv) > 2 else 0.001
sysd = synth.water_box(n, seed=20240001)
s = DeviceRigidBodySystem(sysd["bodyIndices"], sysd["masses"], 0)
s.update(sysd["R"], np.zeros((3*n, 3)), sysd["F"], True, True); s.update(V=sysd["V"], geometry=False, velocities=True); s.upload()
dev = torch.device("cuda:0")
R, V = (torch.from_numpy(sysd[k].copy()).to(dev) for k in ("R", "V"))
F = (torch.from_numpy(sysd["F"].copy()).to(dev), torch.from_numpy(-sysd["F"]).to(dev))
print("start order", s.series_order())
s.part1(dt, R, V, F[0]); print("after part1 order", s.series_order())
for i in range(14):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.part2_part1(dt, R, V, F[(i + 1) & 1]); e1.record(); torch.cuda.synchronize()
    print(f"fused {i}: {e0.elapsed_time(e1)*1e3:.1f} us, next order {s.series_order()}")
