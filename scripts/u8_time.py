import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import yololite_b200 as y
from yololite_b200 import synth
S, B = 640, 64
ck = synth.random_checkpoint(synth.make_meta("edge_n", 80, S), seed=0)
eng = y.YoloLiteB200(ck["state_dict"], ck["meta"], device="cuda:0")
img = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device="cuda")
x, _ = y.preprocess_batch(img, S)
outs = eng(x)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("forward fp32      ms", t(lambda: eng.forward(x, out=outs)))
print("forward u8        ms", t(lambda: eng.forward_u8(img, out=outs)))
print("preprocess        ms", t(lambda: y.preprocess_batch(img, S, out=x)))
