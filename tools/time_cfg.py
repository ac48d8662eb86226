"""Time one configuration's train step on the device (eager, CUDA events): python tools/time_cfg.py cfg5 [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sketchformer_b200 import synth
from sketchformer_b200.engine import Engine, make_cfg
from sketchformer_b200.init import init_params
CFGS = {"cfg2": dict(num_layers=4, d_model=256, dff=1024, num_heads=8, lowerdim=256, max_seq_len=200, vocab_size=1004, n_classes=345),
        "cfg5": dict(num_layers=6, d_model=512, dff=2048, num_heads=8, lowerdim=256, max_seq_len=512, vocab_size=1004, n_classes=345),
        "cfg1": dict(num_layers=4, d_model=128, dff=512, num_heads=8, lowerdim=256, max_seq_len=64, vocab_size=1004, n_classes=10)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
kw = CFGS[name]
eng = Engine(make_cfg(B, kw["max_seq_len"], kw["d_model"], kw["dff"], kw["num_heads"], kw["num_layers"], kw["vocab_size"],
                      kw["n_classes"], kw["lowerdim"]))
eng.load_params(init_params(kw, 0))
x, y = synth.token_batch(B, kw["max_seq_len"], kw["n_classes"], 1)
xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
for _ in range(2): eng.train_step(xd, yd, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
e0.record()
for _ in range(n): eng.train_step(xd, yd, True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{name} B={B} [{os.environ.get('SKF_OPTIONS', '')}]: {ms:.2f} ms/step, {B / ms * 1e3:.0f} sketches/s")
