"""Hot spots of an ncu report with source info: per kernel, instruction mix and the most-sampled SASS instructions
with their stall reasons (run where ncu is installed): python tools/ncu_hot.py report.ncu-rep [kernel index] [top]"""
import collections
import csv
import re
import subprocess
import sys


def main(path, which=0, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = [r[1]]
            blocks.append(cur)
        elif cur is not None:
            cur.append(r)
    name, hdr, data = blocks[which][0], blocks[which][1], blocks[which][2:]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in data if len(r) > ix["Instructions Executed"]]

    def I(r, k):
        try:
            return int(r[ix[k]] or 0)
        except ValueError:
            return 0
    tot = sum(I(r, "Instructions Executed") for r in data)
    tots = sum(I(r, "# Samples") for r in data)
    print(f"== {name[:100]}\n   {len(blocks)} kernels in report; warp instructions {tot}, samples {tots}")
    byop, samp = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        op = m.group(2).split(".")[0] if m else r[ix["Source"]][:10]
        byop[op] += I(r, "Instructions Executed")
        samp[op] += I(r, "# Samples")
    for op, n in byop.most_common(14):
        print(f"   {op:10s} {n:10d} {100 * n / max(tot, 1):5.1f}%  samples {samp[op]}")
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {c: sum(I(r, c) for r in data) for c in stall}
    print("   stalls:", ", ".join(f"{k[6:]} {v}" for k, v in sorted(agg.items(), key=lambda t: -t[1])[:8]))
    for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:top]:
        why = sorted(((I(r, c), c[6:]) for c in stall), reverse=True)[:2]
        print(f"   {I(r, '# Samples'):5d} {I(r, 'Instructions Executed'):8d}  {r[ix['Source']][:70]:70s} {why}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
