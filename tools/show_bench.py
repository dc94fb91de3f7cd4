import json
import sys
for line in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin):   # pass the file name: reading a terminal-less stdin blocks under gpurun
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    for k in ("value", "ms_per_step", "gpu_launches", "clocks", "stage_ms", "e2e", "roofline", "cpu_baseline"):
        if k in d:
            print(k, "=", d[k])
