#!/usr/bin/env python
"""Per-kernel SASS mnemonic histogram of librbk.so (cuobjdump -sass), the evidence for what the kernels use:
LDGSTS = cp.async staging, UBLKCP + SYNCS = TMA bulk copies (cp.async.bulk) completing on an mbarrier, UTMALDG = 2-D TMA tensor copies, SHFL = warp-shuffle segmented reduction, DFMA/DMUL/DADD = fp64 pipe, MUFU.RCP64H/RSQ64H
= fp64 reciprocal / rsqrt seeds, no HMMA/UTC*MMA (nothing here is a dense contraction).

    python tools/sass_summary.py > profiles/rNN_sass_summary.txt ; the full listing goes to profiles/rNN_librbk_sass.txt.gz
"""
import collections
import gzip
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "openmm_rigidbody_plugin_b200", "lib", "librbk.so")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    if len(sys.argv) > 1:
        with gzip.open(sys.argv[1], "wt") as f:
            f.write(sass)
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"rbk::\(anonymous namespace\)::", "", cur)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            kernels[cur][m.group(2)] += 1
    for name, ops in kernels.items():
        tot = sum(ops.values())
        base = collections.Counter()
        for o, n in ops.items():
            base[o.split(".")[0]] += n
        print(f"== {name}\n   {tot} SASS instructions")
        print("   " + ", ".join(f"{o} {n}" for o, n in base.most_common(14)))
        special = {k: v for k, v in ops.items() if k.startswith(("LDGSTS", "SHFL", "MUFU", "BAR", "LDGDEPBAR", "DEPBAR", "UBLKCP", "UTMALDG", "UTMAPF", "SYNCS", "FENCE", "HMMA", "UTC", "ATOM", "RED", "STL", "LDL"))}
        print("   notable: " + ", ".join(f"{k} {v}" for k, v in sorted(special.items())))
        print()


if __name__ == "__main__":
    main()
