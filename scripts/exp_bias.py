"""Experiment (GPU box): signed error of the tcgen05 3xTF32 pointwise conv vs fp64, to characterise accumulator rounding."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yololite_b200 import _lib as L, packer

def run(K, N, M, positive, tc):
    g = np.random.RandomState(1)
    w = g.randn(K, N) / np.sqrt(K)
    x = g.randn(M, K)
    if positive:
        w, x = np.abs(w), np.abs(x)
    x32 = x.astype(np.float32); w32 = w.astype(np.float32)
    exact = x32.astype(np.float64) @ w32.astype(np.float64)
    blob, off = [], [0]
    def add(a):
        a = np.ascontiguousarray(a, np.float32).reshape(-1); o = off[0]; blob.append(a)
        pad = (-a.size) % 64
        if pad: blob.append(np.zeros(pad, np.float32))
        off[0] += a.size + pad; return o
    op = L.YlOp()
    op.kind, op.k, op.stride, op.act, op.anchors = 1, 1, 1, 0, 0
    op.src, op.dst, op.res, op.up = 0, 1, -1, -1
    op.cin, op.cout = K, N
    wm = np.zeros((K, (N + 3) // 4 * 4)); wm[:, :N] = w32
    op.w_off = add(wm); op.wt_off = add(packer.tc_image(wm, N)); op.b_off = add(np.zeros(N))
    bl = torch.from_numpy(np.concatenate(blob)).cuda()
    xd = torch.from_numpy(x32).cuda().reshape(1, M, 1, K).contiguous()
    out = torch.empty((1, M, 1, N), device="cuda")
    L.check(L.lib().yl_run_op(ctypes.byref(op), bl.data_ptr(), xd.data_ptr(), None, None, out.data_ptr(), 1, M, 1, 0, 0, 2 if tc else 0, None))
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(M, N).astype(np.float64)
    rel = (got - exact) / np.maximum(np.abs(exact), 1e-30)
    if not positive:
        scale = np.sqrt((exact ** 2).mean())
        rel = (got - exact) * np.sign(exact) / scale          # signed toward/away from zero, relative to the rms magnitude
    return rel.mean(), np.sqrt((rel ** 2).mean()), np.abs(rel).max()

for positive in (True, False):
    for K in (32, 96, 256, 960):
        for tc in (1, 0):
            m, r, mx = run(K, 64, 1024, positive, tc)
            print(f"positive={positive} K={K:4d} tc={tc}: mean signed rel err {m:+.3e} ({m * 2**24:+.2f} ulp)  rms {r:.3e}  max {mx:.3e}")
