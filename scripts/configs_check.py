"""Run the other BASELINE configs at full size once (crash / capacity check + timing): edge_m 640 (config 3 shard),
YOLOLiteMS-arch FPN+head with P2 on the mobilenetv4 backbone (config 5 analogue), edge_n 320 batch 1 (config 0)."""
import sys, os, time, torch
sys.path.insert(0, os.getcwd())
import yololite_b200 as y
from yololite_b200 import synth

def run(model, B, S, **kw):
    meta = synth.make_meta(model, 80, S, **kw)
    ck = synth.random_checkpoint(meta, seed=0, obj_bias=-3.0)
    eng = y.YoloLiteB200(ck["state_dict"], ck["meta"], device="cuda:0")
    post = y.PostProcessor()
    x = torch.randn((B, 3, S, S), device="cuda")
    outs = eng(x)
    d = post(outs, S, 0.25, 0.5, 300, cap=1024)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        outs = eng(x)
        d = post(outs, S, 0.25, 0.5, 300, cap=1024)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fin = all(bool(torch.isfinite(o).all()) for o in outs)
    print(f"{model} {kw} B={B} S={S}: {ms:.3f} ms/step = {B / ms * 1e3:.0f} img/s, levels {[tuple(o.shape) for o in outs]}, finite={fin}, ops={len(eng.program.ops)}")

run("edge_n", 1, 320)
run("edge_m", 32, 640)
for m in synth.MODEL_YAMLS:
    if m not in ("edge_n", "edge_m"):
        try:
            run(m, 16, 640, use_p2=True)
        except Exception as e:
            print(m, "failed:", str(e)[:200])
