"""Diagnostics (GPU box): per-level logit error of the engine (tcgen05 and fp32 SIMT) against the fp32 and the fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yololite_b200 as y
from oracle import model_ref

cases = [("edge_n", 80, 320, 2, False, False), ("edge_m", 80, 320, 2, False, False), ("edge_m", 3, 640, 1, False, False),
         ("edge_l", 80, 320, 1, True, False), ("edge_s", 13, 352, 2, False, True), ("ms_n_mnv4", 80, 320, 1, False, False),
         ("ms_m_mnv4", 80, 256, 1, True, False)]
for model, nc, S, B, p2, p6 in cases:
    meta = model_ref.make_meta(model, nc, S, use_p2=p2, use_p6=p6)
    ck = model_ref.synth_checkpoint(meta, seed=7)
    x = model_ref.synth_input(B, S, seed=11)
    w32 = model_ref.forward_ref(ck["state_dict"], meta, x)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in ck["state_dict"].items()}
    w64 = model_ref.forward_ref(sd64, meta, x.double())
    print(model, S, "absmax", ["%.1f" % float(b.abs().max()) for b in w64], "| fp32-vs-fp64", ["%.1e" % float((a.double() - b).abs().max()) for a, b in zip(w32, w64)])
    for tc in (True, False):
        eng = y.YoloLiteB200(ck["state_dict"], meta, device="cuda:0", tensor_cores=tc)
        o = [t.cpu() for t in eng(x.cuda())]
        print("   tc=%d  vs fp32" % tc, ["%.1e" % float((a - b).abs().max()) for a, b in zip(o, w32)], " vs fp64", ["%.1e" % float((a.double() - b).abs().max()) for a, b in zip(o, w64)])
        eng.close()
