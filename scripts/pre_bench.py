"""Preprocess kernel alone (letterbox resize path + the no-resize fast path) on a synthetic uint8 batch."""
import sys, os, torch
sys.path.insert(0, os.getcwd())
import yololite_b200 as y
g = torch.Generator(device='cuda').manual_seed(0)
for shape in ((16, 480, 640, 3), (16, 640, 640, 3)):
    img = torch.randint(0, 256, shape, dtype=torch.uint8, device='cuda', generator=g)
    out = torch.empty((shape[0], 3, 640, 640), device='cuda')
    for _ in range(2):
        y.preprocess_batch(img, 640, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        y.preprocess_batch(img, 640, out=out)
    e1.record(); torch.cuda.synchronize()
    print(shape, "ms", e0.elapsed_time(e1) / 3)
