"""Builds a timing-experiment variant of the library next to the product one (never loaded unless SKF_LIB points at it):
    python tools/build_variant.py noasplit -DSKF_T3_DEBUG_SKIP_ASPLIT      -> sketchformer_b200/libskf_noasplit.so
Only the sources that see the define are recompiled into tools/probe/; the other objects are the product build's."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sketchformer_b200 import build as b      # noqa: E402

tag, defines = sys.argv[1], sys.argv[2:]
b.build()
objs = []
for src in b.SOURCES:
    obj = os.path.join(b.CSRC, src.replace(".cu", ".o"))
    text = open(os.path.join(b.CSRC, src)).read()
    if any(d[2:].split("=")[0] in text for d in defines):
        obj = os.path.join(ROOT, "tools", "probe", f"{tag}_{src.replace('.cu', '.o')}")
        cmd = [b._nvcc()] + [f for f in b.NVCC_FLAGS if f not in ("--use_fast_math=false", "-Xptxas", "-v")] + defines + \
              ["-c", os.path.join(b.CSRC, src), "-o", obj]
        subprocess.run(cmd, check=True)
    objs.append(obj)
out = os.path.join(ROOT, "sketchformer_b200", f"libskf_{tag}.so")
subprocess.run([b._nvcc(), "-shared", "-o", out] + objs + ["-ldl"], check=True)
print(out)
