"""stdin -> the JSON line(s) a bench.py run printed (launchers such as torchrun / NCCL add lines of their own around it)."""
import sys
for line in sys.stdin:
    if line.lstrip().startswith("{"):
        sys.stdout.write(line.strip() + "\n")
