import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith("{"): continue
    d=json.loads(line); print(d["value"], d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["roofline"]["kernels"].items()}, round(d["roofline"]["step"]["frac"],3))
