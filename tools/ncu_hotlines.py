"""Print the hottest SASS lines (warp-stall samples) of each kernel in an .ncu-rep (run where ncu is available).
usage: python tools/ncu_hotlines.py <report.ncu-rep> [min_samples]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}
        blocks.append(cur)
    elif cur is not None:
        if cur["hdr"] is None:
            cur["hdr"] = row
        else:
            cur["rows"].append(row)
for b in blocks:
    hdr = b["hdr"]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    total = 0
    lines = []
    for k, r in enumerate(b["rows"]):
        if len(r) < len(hdr):
            continue
        n = int(r[ix["# Samples"]] or 0)
        total += n
        st = {s: int(r[ix[s]] or 0) for s in stalls}
        for s, v in st.items():
            tot[s] += v
        lines.append((k, n, int(r[ix["Instructions Executed"]] or 0), r[ix["Source"]][:95], max(st.items(), key=lambda kv: kv[1])))
    print("==", b["name"][:100], "samples", total)
    print("   ", ", ".join(f"{s[6:]} {100 * v / max(1, total):.0f}%" for s, v in tot.most_common(7)))
    for k, n, ie, src, top in lines:
        if n >= thr:
            print(f"   {k:5d} {n:6d} {ie:9d}  {src}  [{top[0][6:]} {top[1]}]")
