"""Summarise ncu output for profiles/:
  python tools/ncu_summary.py launches <ncu --csv log> <out.csv> [header comment lines...]
  python tools/ncu_summary.py full <report.ncu-rep> <out.csv> [header comment lines...]
"""
import collections
import csv
import subprocess
import sys

FULL_COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
             "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
             "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def launches(src, dst, comments):
    rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1000.0 if r[mu] in ("ns", "nsecond") else v * 1000.0 if r[mu] in ("ms", "msecond") else v
        name = r[kn].replace("void ", "").replace("<unnamed>::", "").split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        for c in comments:
            f.write("# " + c + "\n")
        f.write("kernel,launches,total_us,share_pct,avg_us\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{a[0]},{a[1]:.1f},{100 * a[1] / tot:.2f},{a[1] / a[0]:.1f}\n')
        f.write(f"# total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches\n")


def full(src, dst, comments):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) for c in FULL_COLS if c in hdr]
    with open(dst, "w") as f:
        for c in comments:
            f.write("# " + c + "\n")
        w = csv.writer(f)
        w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx])
        for r in rows[2:]:
            w.writerow([r[i].replace("void ", "").replace("<unnamed>::", "").split("(")[0] if hdr[i] == "Kernel Name" else r[i]
                        for i in idx])


def metrics(src, dst, comments):
    """ncu --metrics a,b,c --csv log (one row per launch per metric) -> one row per launch."""
    rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    idc, kn, mn, mu, mv = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    per, names, order = {}, [], []
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        key = r[idc]
        if key not in per:
            per[key] = {"kernel": r[kn].replace("void ", "").replace("<unnamed>::", "").split("(")[0]}
            order.append(key)
        col = f"{r[mn]} [{r[mu]}]"
        if col not in names:
            names.append(col)
        per[key][col] = r[mv].replace(",", "")
    with open(dst, "w") as f:
        for c in comments:
            f.write("# " + c + "\n")
        w = csv.writer(f)
        w.writerow(["launch", "kernel"] + names)
        for i, key in enumerate(order):
            w.writerow([i, per[key]["kernel"]] + [per[key].get(n, "") for n in names])


def traffic(src, dst, comments):
    """per-launch metrics CSV (the `metrics` mode's output) -> the JSON bench.py's roofline.traffic is read from."""
    import json
    rows = list(csv.DictReader(l for l in open(src) if not l.startswith("#")))
    col = lambda key: next(c for c in rows[0] if c.startswith(key))
    rd, wr, tm, tp = (col(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                         "sm__pipe_tensor_cycles_active"))
    tot_r = sum(float(r[rd]) for r in rows)
    tot_w = sum(float(r[wr]) for r in rows)
    tot_t = sum(float(r[tm]) for r in rows)
    out = {
        "kernel": "gemm_tc_kernel",
        "launches": len(rows),
        "dram_bytes_per_launch": (tot_r + tot_w) / len(rows),
        "dram_read_bytes_total": tot_r,
        "dram_write_bytes_total": tot_w,
        "ncu_time_us_total": tot_t / 1e3,
        "tensor_pipe_pct_time_weighted": sum(float(r[tp]) * float(r[tm]) for r in rows) / tot_t,
        "source": " ".join(comments),
    }
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    {"launches": launches, "full": full, "metrics": metrics, "traffic": traffic}[mode](src, dst, sys.argv[4:])
