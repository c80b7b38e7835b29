import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import yololite_b200 as y
from conftest import synth_ckpt
S = 64
ck = synth_ckpt("edge_n", 3, S)
eng = y.YoloLiteB200(ck["state_dict"], ck["meta"], device="cuda:0")
img = torch.randint(0, 256, (3, S, S, 3), dtype=torch.uint8, device="cuda")
x, _ = y.preprocess_batch(img, S)
want = eng(x); torch.cuda.synchronize(); print("fp32 path ok")
t0 = time.time()
try:
    got = eng.forward_u8(img); torch.cuda.synchronize()
    print("u8 path ok", time.time() - t0, max(float((g - w).abs().max()) for g, w in zip(got, want)))
except Exception as e:
    print("u8 path failed after", time.time() - t0, str(e)[:200])
