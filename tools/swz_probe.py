import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sketchformer_b200 import _lib
lib = C.CDLL(_lib.LIB_PATH)
rows = 16
src = torch.arange(rows * 32, dtype=torch.float32, device="cuda")
for mn in (0, 1):
    out = torch.zeros(rows * 32, device="cuda")
    rc = lib.skf_debug_tma_dump(C.c_void_p(src.data_ptr()), rows, mn, C.c_void_p(out.data_ptr()), None)
    torch.cuda.synchronize()
    o = out.cpu().view(rows, 32).long()
    print("mn", mn, "rc", rc)
    for r in range(rows):
        # for each physical 16B chunk p of row r: which logical chunk (col/4) sits there
        print(r, [(int(o[r, 4 * p]) - 32 * r) // 4 for p in range(8)], "row ok" if all(o[r, i] // 32 == r for i in range(32)) else "ROW MIX")
