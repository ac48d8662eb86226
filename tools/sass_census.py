"""Opcode census of the built library: per kernel, how many tcgen05 / TMEM / TMA / mbarrier SASS instructions it holds
(`cuobjdump -sass libskf.so`; mnemonics from /opt/skills/guides/B200_PROFILING.md).  Runs without a GPU.
    python tools/sass_census.py > profiles/rN_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sketchformer_b200", "libskf.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "UTCBAR.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF",
       "SYNCS", "REDG", "MUFU.EX2", "HMMA", "FFMA"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                       text=True).stdout.split("\n")
counts, order, cur, k = {}, [], None, 0
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", names[k]).replace("skf::", "").replace("void ", "")
        k += 1
        if cur not in counts:
            counts[cur] = collections.Counter()
            order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or cur is None:
        continue
    op = m.group(1)
    c = counts[cur]
    c["total"] += 1
    if op.startswith("UTCHMMA"):
        c["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
    elif op.startswith("UTCBAR"):
        c["UTCBAR.2CTA" if ".2CTA" in op else "UTCBAR"] += 1
    else:
        for o in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "REDG", "MUFU.EX2", "HMMA", "FFMA"):
            if op.startswith(o):
                c[o] += 1
                break
print("SASS opcode census of", os.path.relpath(lib, ROOT), "(sm_100a)")
print("UTCHMMA = tcgen05.mma, .2CTA = cta_group::2; UTCBAR = tcgen05.commit; LDTM/STTM = tcgen05.ld/st; "
      "UTMALDG/UTMASTG/UTMAREDG = TMA load/store/reduce; SYNCS = mbarrier ops")
tot = collections.Counter()
print(f"{'kernel':64s} " + " ".join(f"{o:>7s}" for o in OPS) + "   total")
for name in order:
    c = counts[name]
    tot.update(c)
    if not any(c[o] for o in OPS[:10]):
        continue                      # SIMT-only kernels are summed below
    print(f"{name[:64]:64s} " + " ".join(f"{c[o]:7d}" for o in OPS) + f" {c['total']:7d}")
print(f"{'ALL KERNELS (' + str(len(order)) + ')':64s} " + " ".join(f"{tot[o]:7d}" for o in OPS) + f" {tot['total']:7d}")
