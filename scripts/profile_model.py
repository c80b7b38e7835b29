"""Per-op device times of one forward for any synthetic model: python scripts/profile_model.py edge_m 32 640"""
import sys, os, torch
sys.path.insert(0, os.getcwd())
import yololite_b200 as y
from yololite_b200 import synth, _lib as L
model, B, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
meta = synth.make_meta(model, 80, S)
ck = synth.random_checkpoint(meta, seed=0)
eng = y.YoloLiteB200(ck["state_dict"], ck["meta"], device="cuda:0")
x = torch.randn((B, 3, S, S), device="cuda")
eng(x); torch.cuda.synchronize()
t0, s0 = L.lib().yl_stat(b"tc_launches"), L.lib().yl_stat(b"simt_launches")
res = eng.profile_ops(x)
print("tc launches", L.lib().yl_stat(b"tc_launches") - t0, "simt launches", L.lib().yl_stat(b"simt_launches") - s0)
kinds = {0: "stem", 1: "conv", 2: "dw", 3: "dwpw", 4: "stem2"}
tot = sum(t for _, t in res)
for i, (op, t) in enumerate(res):
    print(f"{i:3d} {kinds[op['kind']]:5s} {op['cin']:4d}->{op['cout']:4d} k{op['k']} s{op['stride']} k2={op['k2']} s2={op.get('stride2',0)} res={op['res']>=0} up={op['up']>=0} {t*1000:8.0f} us")
print("total ms", tot)
