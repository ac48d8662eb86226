"""Per-kernel DRAM traffic and achieved GB/s from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` launch list (the HBM-bound kernels are reported as GB/s against the
measured copy bandwidth in MEASURED_PEAKS.json).  Per-launch times under ncu are cold-cache and serialised."""
import collections
import csv
import json
import os
import re
import sys


def to_us(v, unit):
    return v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v * 1e6 if unit in ("s", "second") else v


def to_bytes(v, unit):
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main(path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = 6650.0
    pk = os.path.join(root, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.defaultdict(dict)
    names = {}
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        per[row["ID"]][row["Metric Name"]] = (v, row["Metric Unit"])
        name = re.sub(r"^void |skf::", "", row["Kernel Name"])
        names[row["ID"]] = re.sub(r"\(CUtensorMap.*|\(const .*|\(float.*|\(long.*|\(int.*|\(unsigned.*|\(double.*|\(__nv.*|\(unsigned char.*", "", name)[:60]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i, m in per.items():
        if "gpu__time_duration.sum" not in m:
            continue
        a = agg[names[i]]
        a[0] += 1
        a[1] += to_us(*m["gpu__time_duration.sum"])
        a[2] += to_bytes(*m.get("dram__bytes_read.sum", (0, "byte")))
        a[3] += to_bytes(*m.get("dram__bytes_write.sum", (0, "byte")))
    print(f"HBM peak {peak:.1f} GB/s (MEASURED_PEAKS.json hbm_gbs, else fallback)")
    print(f"{'kernel':60s} {'n':>4s} {'us/launch':>10s} {'MB read':>9s} {'MB write':>9s} {'GB/s':>8s} {'of peak':>8s}")
    for k, (n, us, rd, wr) in sorted(agg.items(), key=lambda t: -t[1][1]):
        gbs = (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else 0.0
        print(f"{k:60s} {n:4d} {us / n:10.1f} {rd / n / 1e6:9.2f} {wr / n / 1e6:9.2f} {gbs:8.0f} {gbs / peak:8.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
