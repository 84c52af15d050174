"""Record the reference's own CUDA kernels (baseline/ref_cuda) on a B200 as committed vectors - run ON THE GPU BOX:

    gpurun -- 'python tests/golden/make_golden_refcuda.py gpurun_out/golden_refcuda'

then copy gpurun_out/golden_refcuda/*.npz into tests/golden/.  Each file holds the inputs (quantised forces, atom order),
the trajectory end point and the refined kinetic energies / potential-energy refinement the reference kernels computed
(COMPMOD = 1, rigidbodyintegrator.cu:433-469) - the only implementation of those diagnostics in the reference."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_gpu_ref_cuda as t  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for case in ("water", "mixed"):
        for precision, mode in (("double", 0), ("mixed", 0), ("double", 3)):
            steps = 4
            s, ref, A, B, dt = t.both(case, precision, mode, 1, steps=steps)
            theirs = np.concatenate([ref.kinetic(B.velm, dt, refined=True), [ref.potential_refinement(dt, B.velm, B.force)]])
            sysd, _ = t.make_case(case)
            R_end, V_end = B.host()
            out = {k: sysd[k] for k in ("bodyIndices", "masses", "R", "V", "F", "charges")}
            out.update(order=B.order, precision=precision, mode=np.int32(mode), steps=np.int32(steps), dt=np.float64(dt),
                       reference=theirs, R_end=R_end, V_end=V_end)
            path = os.path.join(out_dir, f"refcuda_refined_{case}_{precision}_m{mode}.npz")
            np.savez_compressed(path, **out)
            print(path, theirs)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_refcuda"))
