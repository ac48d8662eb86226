"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel family."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        name = re.sub(r"^void |skf::", "", row["Kernel Name"])
        name = re.sub(r"\(CUtensorMap.*|\(const .*|\(float.*|\(long.*|\(int.*|\(unsigned.*|\(double.*", "", name)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"total {T:.1f} us over {sum(cnt.values())} launches")
    for k, v in sorted(tot.items(), key=lambda t: -t[1])[:top]:
        print(f"{v:10.1f} us {100 * v / T:5.1f}%  n={cnt[k]:4d} avg={v / cnt[k]:8.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
