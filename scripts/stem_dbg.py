import ctypes, sys, os, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from yololite_b200 import _lib as L, packer
g = torch.Generator().manual_seed(1)
B, H, W, n2 = 2, 64, 64, 16
x = torch.randn(B, 3, H, W, generator=g)
ws = torch.randn(32, 3, 3, 3, generator=g) / 5
bs = torch.randn(32, generator=g) * 0.3
w2 = torch.randn(n2, 32, 3, 3, generator=g) / 17
b2 = torch.randn(n2, generator=g)
blob, off = [], [0]
def add(a):
    a = np.ascontiguousarray(a, np.float32).reshape(-1); o = off[0]; blob.append(a)
    pad = (-a.size) % 64
    if pad: blob.append(np.zeros(pad, np.float32))
    off[0] += a.size + pad
    return o
op = L.YlOp()
op.kind, op.k, op.stride, op.act, op.anchors, op.k2 = L.OP_STEM2, 3, 2, 1, 0, 32
op.src, op.dst, op.res, op.up = -1, 1, -1, -1
op.cin, op.cout = 3, n2
wm = packer._gemm_w(w2.double().numpy())
op.w_off = add(wm); op.wt_off = add(packer.tc_image(wm, n2))
wsm = np.transpose(ws.double().numpy(), (2, 3, 1, 0)).reshape(27, 32)
op.w2_off = add(np.concatenate([wsm.reshape(-1), bs.double().numpy(), packer.tc_image(np.concatenate([wsm, bs.double().numpy().reshape(1, -1)]), 32).astype(np.float64)]))
op.w3_off = add(packer.stem2_image(wm, n2, wsm, bs.double().numpy()))
op.b_off = add(packer._pad4(b2.double().numpy()))
dblob = torch.from_numpy(np.concatenate(blob)).cuda()
want = F.relu(F.conv2d(F.relu(F.conv2d(x.double(), ws.double(), bs.double(), stride=2, padding=1)), w2.double(), b2.double(), stride=2, padding=1)).permute(0, 2, 3, 1).contiguous()
out = torch.full(want.shape, float("nan"), device="cuda")
xc = x.cuda()
L.check(L.lib().yl_run_op(ctypes.byref(op), dblob.data_ptr(), xc.data_ptr(), None, None, out.data_ptr(), B, H, W, 0, 0, 1, None))
torch.cuda.synchronize()
d = (out.cpu().double() - want).abs()
print(os.environ.get("YL_S2_PASSES"), "max err", float(d.max()), "mean", float(d.mean()), "nan", int(torch.isnan(out).sum()))
