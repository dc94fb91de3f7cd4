import json
import sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    for k in ("value", "ms_per_step", "gpu_launches", "clocks", "stage_ms", "e2e", "roofline", "cpu_baseline"):
        if k in d:
            print(k, "=", d[k])
