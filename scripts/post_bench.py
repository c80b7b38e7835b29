"""Postprocess kernel alone on synthetic logits: detection setting (conf 0.25 / iou 0.5 / 300 per class) and the evaluation
setting of scripts/helpers/helpers.py:86-153 (conf 0.001 / iou 0.65 / unlimited)."""
import sys, os, torch
sys.path.insert(0, os.getcwd())
import yololite_b200 as y
B, S, nc = 64, 640, 80
g = torch.Generator(device='cuda').manual_seed(0)
lv = [torch.randn((B, 1, s, s, 5 + nc), device='cuda', generator=g) for s in (80, 40, 20)]
for l in lv:
    l[..., 4] -= 4.0
post = y.PostProcessor()
for name, conf, iou, md, cap in (("detect", 0.25, 0.5, 300, 1024), ("eval", 0.001, 0.65, 0, None)):
    for _ in range(2):
        d = post(lv, S, conf, iou, md, cap=cap)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        d = post(lv, S, conf, iou, md, cap=cap)
    e1.record(); torch.cuda.synchronize()
    print(name, "ms", e0.elapsed_time(e1) / 3, "dets/img", float((d.counts & 0x3FFFFFFF).float().mean()))
