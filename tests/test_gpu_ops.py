"""Single-kernel numerics through yl_run_op: every conv kernel (fp32 SIMT and tcgen05 bf16-triple) against a plain
PyTorch fp32 (CPU) convolution of the same op.  Tolerance 2e-4 abs on O(1) activations (fp32 accumulation-order
noise is ~1e-6; a single-pass TF32 product would show ~1e-2 here)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 2e-4


def _run(kind, x_nhwc, w, bias, k, stride, act=0, res=None, up=None, anchors=0, w2=None, use_tc=1, nchw_input=False,
         b2=None, act2=0, stride2=0, tap_layout=False):
    """w: [Cout,Cin,k,k] torch fp32 (dense) or [C,1,k,k] (depthwise); returns NHWC output (or head layout)."""
    from yololite_b200 import _lib as L, packer
    lib = L.lib()
    blob, off = [], [0]

    def add(a):
        a = np.ascontiguousarray(a, np.float32).reshape(-1)
        o = off[0]
        blob.append(a)
        pad = (-a.size) % 64
        if pad:
            blob.append(np.zeros(pad, np.float32))
        off[0] += a.size + pad
        return o
    cout = w.shape[0]
    wn = w.double().numpy()
    op = L.YlOp()
    op.kind, op.k, op.stride, op.act, op.anchors = kind, k, stride, act, anchors
    op.src, op.dst, op.res, op.up = 0, 1, (2 if res is not None else -1), (3 if up is not None else -1)
    op.k2, op.w2_off, op.wt_off, op.w3_off, op.b2_off, op.act2 = 0, -1, -1, -1, -1, 0
    if kind == L.OP_DW:
        op.cin = op.cout = cout
        op.w_off = add(np.transpose(wn, (2, 3, 1, 0)).reshape(k * k, cout))
    elif kind == L.OP_STEM:
        op.cin, op.cout = 3, cout
        wm = np.transpose(wn, (2, 3, 1, 0)).reshape(-1, cout)
        op.w_off = add(wm)
    else:
        op.cin, op.cout = w.shape[1], cout
        wm = packer._gemm_w(wn)
        op.w_off = add(wm)
        if tap_layout:      # per-tap padded K axis (yl_op.wt_layout = 1): the TMA-fed dense conv path
            op.wt_off = add(packer.tc_image(packer.tap_padded(wm, k * k, w.shape[1]), cout))
            op.wt_layout = 1
        else:
            op.wt_off = add(packer.tc_image(wm, cout))
        if kind == L.OP_DWPW:
            op.k, op.k2 = 1, int(w2.shape[-1])
            op.w2_off = add(np.transpose(w2.double().numpy(), (2, 3, 1, 0)).reshape(op.k2 * op.k2, -1))
            op.act2, op.stride2 = act2, stride2
            if b2 is not None:
                op.b2_off = add(packer._pad4(b2.double().numpy()))
    op.b_off = add(packer._pad4(bias.double().numpy())) if bias is not None else -1
    dblob = torch.from_numpy(np.concatenate(blob)).cuda()
    xin = x_nhwc.cuda().contiguous()
    B = xin.shape[0]
    Hin, Win = (xin.shape[2], xin.shape[3]) if nchw_input else (xin.shape[1], xin.shape[2])
    kk = 3 if kind == L.OP_DWPW else k
    ho, wo = (Hin + 2 * (k // 2) - k) // stride + 1, (Win + 2 * (k // 2) - k) // stride + 1
    if kind == L.OP_DWPW:
        s2, k2 = max(1, stride2), int(w2.shape[-1])
        ho, wo = (Hin + 2 * (k2 // 2) - k2) // s2 + 1, (Win + 2 * (k2 // 2) - k2) // s2 + 1
    out = torch.full((B, ho, wo, cout), float("nan"), device="cuda")
    rs = res.cuda().contiguous() if res is not None else None
    us = up.cuda().contiguous() if up is not None else None
    L.check(lib.yl_run_op(ctypes.byref(op), dblob.data_ptr(), xin.data_ptr(), rs.data_ptr() if rs is not None else None,
                          us.data_ptr() if us is not None else None, out.data_ptr(), B, Hin, Win,
                          us.shape[1] if us is not None else 0, us.shape[2] if us is not None else 0, use_tc, None))
    torch.cuda.synchronize()
    return out.cpu()


def _ref(x_nchw, w, bias, k, stride, act, groups=1, res=None, up=None):
    y = F.conv2d(x_nchw, w, bias, stride=stride, padding=k // 2, groups=groups)
    if res is not None:
        y = y + res.permute(0, 3, 1, 2)
    if up is not None:
        y = y + F.interpolate(up.permute(0, 3, 1, 2), size=y.shape[-2:], mode="nearest")
    y = F.relu(y) if act == 1 else F.silu(y) if act == 2 else y
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("use_tc", [0, 2])
@pytest.mark.parametrize("cin,cout,hw,act", [(16, 16, 40, 1), (48, 32, 20, 1), (32, 96, 24, 1), (96, 48, 17, 0), (64, 256, 10, 1),
                                              (256, 64, 10, 0), (288, 64, 9, 0), (64, 480, 7, 1), (96, 96, 33, 2), (480, 96, 6, 0),
                                              (244, 244, 5, 1)])
def test_pointwise(use_tc, cin, cout, hw, act):
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    x = torch.randn(3, hw, hw + 1, cin, generator=g)
    w = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
    b = torch.randn(cout, generator=g)
    got = _run(1, x, w, b, 1, 1, act, use_tc=use_tc)
    want = _ref(x.permute(0, 3, 1, 2), w, b, 1, 1, act)
    assert float((got - want).abs().max()) <= TOL


@pytest.mark.parametrize("use_tc", [0, 2])
def test_pointwise_residual_upsample_and_head_layout(use_tc):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 13, 11, 48, generator=g)
    w = torch.randn(96, 48, 1, 1, generator=g) / 7
    b = torch.randn(96, generator=g)
    res = torch.randn(2, 13, 11, 96, generator=g)
    up = torch.randn(2, 7, 6, 96, generator=g)                       # 7x6 -> 13x11 is not an exact 2x
    got = _run(1, x, w, b, 1, 1, 0, res=res, up=up, use_tc=use_tc)
    want = _ref(x.permute(0, 3, 1, 2), w, b, 1, 1, 0, res=res, up=up)
    assert float((got - want).abs().max()) <= TOL
    # head layout: N = A*(5+C) stored as [B,A,H,W,5+C]
    A, D = 2, 9
    w = torch.randn(A * D, 48, 1, 1, generator=g) / 7
    b = torch.randn(A * D, generator=g)
    got = _run(1, x, w, b, 1, 1, 0, anchors=A, use_tc=use_tc).reshape(-1)
    want = _ref(x.permute(0, 3, 1, 2), w, b, 1, 1, 0)               # [B,H,W,A*D]
    want = want.view(2, 13, 11, A, D).permute(0, 3, 1, 2, 4).contiguous().reshape(-1)
    assert float((got - want).abs().max()) <= TOL
    # N = 85 (5+80 classes), A = 1: rows are not 16-byte aligned
    w = torch.randn(85, 48, 1, 1, generator=g) / 7
    b = torch.randn(85, generator=g)
    got = _run(1, x, w, b, 1, 1, 0, anchors=1, use_tc=use_tc)
    assert float((got - _ref(x.permute(0, 3, 1, 2), w, b, 1, 1, 0)).abs().max()) <= TOL


@pytest.mark.parametrize("use_tc", [0, 2])
@pytest.mark.parametrize("cin,cout,stride,hw,act", [(32, 16, 2, 40, 1), (16, 48, 2, 21, 1), (96, 96, 1, 12, 2), (96, 96, 2, 11, 1)])
def test_dense3x3(use_tc, cin, cout, stride, hw, act):
    g = torch.Generator().manual_seed(cin + cout + stride)
    x = torch.randn(2, hw, hw + 2, cin, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    b = torch.randn(cout, generator=g)
    got = _run(1, x, w, b, 3, stride, act, use_tc=use_tc)
    want = _ref(x.permute(0, 3, 1, 2), w, b, 3, stride, act)
    assert float((got - want).abs().max()) <= TOL


@pytest.mark.parametrize("use_tc", [0, 2])
def test_stem_nchw(use_tc):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 45, 52, generator=g)
    w = torch.randn(32, 3, 3, 3, generator=g) / 5
    b = torch.randn(32, generator=g)
    got = _run(0, x, w, b, 3, 2, 1, use_tc=use_tc, nchw_input=True)
    want = _ref(x, w, b, 3, 2, 1)
    assert float((got - want).abs().max()) <= TOL


@pytest.mark.parametrize("k,stride", [(3, 1), (3, 2), (5, 1), (5, 2)])
def test_depthwise(k, stride):
    g = torch.Generator().manual_seed(k * 10 + stride)
    x = torch.randn(2, 19, 23, 96, generator=g)
    w = torch.randn(96, 1, k, k, generator=g) / k
    b = torch.randn(96, generator=g)
    got = _run(2, x, w, b, k, stride, 1)
    want = _ref(x.permute(0, 3, 1, 2), w, b, k, stride, 1, groups=96)
    assert float((got - want).abs().max()) <= TOL


@pytest.mark.parametrize("use_tc", [0, 2])
@pytest.mark.parametrize("c,hw", [(96, 20), (96, 37), (244, 9)])
def test_fused_dw_pw(use_tc, c, hw):
    g = torch.Generator().manual_seed(c + hw)
    x = torch.randn(2, hw, hw - 1, c, generator=g)
    wd = torch.randn(c, 1, 3, 3, generator=g) / 3
    wp = torch.randn(c, c, 1, 1, generator=g) / c ** 0.5
    b = torch.randn(c, generator=g)
    got = _run(3, x, wp, b, 1, 1, 1, w2=wd, use_tc=use_tc)
    mid = F.conv2d(x.permute(0, 3, 1, 2), wd, None, padding=1, groups=c)
    want = _ref(mid, wp, b, 1, 1, 1)
    assert float((got - want).abs().max()) <= TOL


@pytest.mark.parametrize("use_tc", [0, 2])
@pytest.mark.parametrize("cin,cout,k2,h,w,res", [(96, 48, 3, 40, 40, True), (48, 192, 3, 40, 40, False), (256, 64, 5, 20, 20, True),
                                                 (256, 64, 3, 20, 20, True), (192, 64, 5, 20, 20, True), (64, 256, 5, 20, 20, False),
                                                 (32, 96, 5, 80, 80, False), (48, 288, 3, 23, 37, False), (288, 64, 3, 7, 9, True),
                                                 (64, 64, 5, 3, 2, True)])
def test_fused_dw_pw_uir(use_tc, cin, cout, k2, h, w, res):
    """The dw_start -> pw_exp / dw_mid -> pw_proj (+residual) pairs of the backbone's UIR blocks: depthwise 3x3 / 5x5 with
    its own folded-BN bias and ReLU, pointwise with bias; large-K shapes stream the weight slabs (tcgen05 path)."""
    g = torch.Generator().manual_seed(cin + cout + k2)
    x = torch.randn(2, h, w, cin, generator=g)
    wd = torch.randn(cin, 1, k2, k2, generator=g) / k2
    bd = torch.randn(cin, generator=g) * 0.5
    wp = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
    b = torch.randn(cout, generator=g)
    r = torch.randn(2, h, w, cout, generator=g) if res else None
    act2, act = (1, 0) if res else (0, 1)                 # dw_mid: ReLU then linear projection; dw_start: linear then ReLU
    got = _run(3, x, wp, b, 1, 1, act, w2=wd, use_tc=use_tc, b2=bd, act2=act2, res=r)
    mid = F.conv2d(x.permute(0, 3, 1, 2), wd, bd, padding=k2 // 2, groups=cin)
    mid = F.relu(mid) if act2 else mid
    want = _ref(mid, wp, b, 1, 1, act, res=r)
    assert float((got - want).abs().max()) <= TOL


@pytest.mark.parametrize("use_tc", [0, 2])
@pytest.mark.parametrize("cin,cout,k2,h,w", [(96, 48, 5, 80, 80), (288, 64, 3, 40, 40), (96, 48, 5, 21, 37), (64, 96, 3, 9, 6), (32, 32, 5, 3, 3)])
def test_fused_dw_pw_stride2(use_tc, cin, cout, k2, h, w):
    """dw_mid with stride 2 (first block of a stage) fused with pw_proj: depthwise k x k s2 + bias + ReLU -> pointwise + bias."""
    g = torch.Generator().manual_seed(cin + cout + k2 + h)
    x = torch.randn(2, h, w, cin, generator=g)
    wd = torch.randn(cin, 1, k2, k2, generator=g) / k2
    bd = torch.randn(cin, generator=g) * 0.5
    wp = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
    b = torch.randn(cout, generator=g)
    got = _run(3, x, wp, b, 1, 1, 0, w2=wd, use_tc=use_tc, b2=bd, act2=1, stride2=2)
    mid = F.relu(F.conv2d(x.permute(0, 3, 1, 2), wd, bd, stride=2, padding=k2 // 2, groups=cin))
    want = _ref(mid, wp, b, 1, 1, 0)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= TOL


def test_tensor_core_path_is_not_single_pass_tf32():
    """The 3-pass split must recover fp32 accuracy: compare against an fp64 reference on a long-K product."""
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 16, 16, 288, generator=g) * 4
    w = torch.randn(64, 288, 1, 1, generator=g)
    from yololite_b200 import _lib as L
    n0 = L.lib().yl_stat(b"tc_launches")
    got = _run(1, x, w, None, 1, 1, 0, use_tc=2).double()
    assert L.lib().yl_stat(b"tc_launches") == n0 + 1          # the tcgen05 kernel really ran
    want = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double()).permute(0, 2, 3, 1)
    rel = float((got - want).abs().max() / want.abs().max())
    assert rel < 1e-5, rel          # measured 3e-6; single-pass TF32 gives ~5e-4 here
    n1 = L.lib().yl_stat(b"simt_launches")
    _run(1, x, w, None, 1, 1, 0, use_tc=0)
    assert L.lib().yl_stat(b"simt_launches") == n1 + 1


@pytest.mark.parametrize("bf16x3,pw", [(True, False), (False, False), (True, True)])
@pytest.mark.parametrize("hw,n2", [((64, 64), 16), ((45, 52), 16), ((83, 38), 32), ((640, 640), 16), ((96, 100), 12), ((33, 36), 16)])
def test_fused_stem_conv(hw, n2, bf16x3, pw):
    """YL_OP_STEM2: conv_stem (3x3 s2, 3->32, ReLU) -> 3x3 s2 conv (32->n2, ReLU) in one tcgen05 kernel.
    bf16x3 = the fused bf16-triple kernel (csrc/stem_kernel.cu, w3_off image); otherwise (w3_off = -1, or a shape the fused kernel
    cannot take: W % 4 != 0, n2 % 4 != 0) the unfused fallback: SIMT stem -> 3x3 s2 conv -> pointwise as separate launches."""
    from yololite_b200 import _lib as L, packer
    if pw and n2 != 16:
        pytest.skip("the fused pointwise conv is the 16-channel blocks.0.1 of mobilenetv4_conv_small_050")
    g = torch.Generator().manual_seed(hw[0] + n2)
    B = 2 if hw[0] < 600 else 1
    x = torch.randn(B, 3, hw[0], hw[1], generator=g)
    ws = torch.randn(32, 3, 3, 3, generator=g) / 5
    bs = torch.randn(32, generator=g) * 0.3
    w2 = torch.randn(n2, 32, 3, 3, generator=g) / 17
    b2 = torch.randn(n2, generator=g)
    blob, off = [], [0]

    def add(a):
        a = np.ascontiguousarray(a, np.float32).reshape(-1)
        o = off[0]
        blob.append(a)
        pad = (-a.size) % 64
        if pad:
            blob.append(np.zeros(pad, np.float32))
        off[0] += a.size + pad
        return o
    op = L.YlOp()
    op.kind, op.k, op.stride, op.act, op.anchors, op.k2 = L.OP_STEM2, 3, 2, 1, 0, 32
    op.src, op.dst, op.res, op.up = -1, 1, -1, -1
    op.cin, op.cout = 3, n2
    wm = packer._gemm_w(w2.double().numpy())
    op.w_off = add(wm)
    op.wt_off = add(packer.tc_image(wm, n2))
    wsm = np.transpose(ws.double().numpy(), (2, 3, 1, 0)).reshape(27, 32)
    op.w2_off = add(np.concatenate([wsm.reshape(-1), bs.double().numpy()]))
    op.w3_off = add(packer.stem2_image(wm, n2, wsm, bs.double().numpy())) if bf16x3 else -1
    op.b_off = add(packer._pad4(b2.double().numpy()))
    want = F.relu(F.conv2d(F.relu(F.conv2d(x, ws, bs, stride=2, padding=1)), w2, b2, stride=2, padding=1))
    if pw:      # blocks.0.1: 1x1 16 -> 16 + bias + ReLU in the output epilogue
        wq = torch.randn(16, 16, 1, 1, generator=g) / 4
        bq = torch.randn(16, generator=g)
        op.b2_off = add(np.concatenate([wq[:, :, 0, 0].t().double().numpy().reshape(-1), bq.double().numpy()]))
        op.act2 = 1
        want = F.relu(F.conv2d(want, wq, bq))
    want = want.permute(0, 2, 3, 1).contiguous()
    dblob = torch.from_numpy(np.concatenate(blob)).cuda()
    out = torch.full(want.shape, float("nan"), device="cuda")
    xc = x.cuda()
    L.check(L.lib().yl_run_op(ctypes.byref(op), dblob.data_ptr(), xc.data_ptr(), None, None, out.data_ptr(), B, hw[0], hw[1], 0, 0, 1, None))
    torch.cuda.synchronize()
    assert float((out.cpu() - want).abs().max()) <= TOL


@pytest.mark.parametrize("cin,cout,h,w,act,res", [(96, 96, 20, 20, 2, False), (196, 196, 13, 11, 2, False), (328, 328, 16, 24, 2, False),
                                                   (72, 40, 9, 33, 1, True), (128, 256, 8, 16, 0, False)])
def test_dense3x3_tma_tap_path(cin, cout, h, w, act, res):
    """Dense 3x3 stride-1 conv with a long Cin (the YOLOLiteMS FPN smoothing convs, model_v2.py:15-22,201-203): every K-slab of the A
    operand is ONE TMA box of the NHWC input shifted by the tap (zero-filled outside = the padding); weights resident or streamed."""
    from yololite_b200 import _lib as L
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(2, h, w, cin, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    b = torch.randn(cout, generator=g)
    r = torch.randn(2, h, w, cout, generator=g) if res else None
    n0 = L.lib().yl_stat(b"tc_launches")
    got = _run(1, x, wt, b, 3, 1, act, res=r, use_tc=2, tap_layout=True)
    assert L.lib().yl_stat(b"tc_launches") == n0 + 1
    want = _ref(x.permute(0, 3, 1, 2), wt, b, 3, 1, act, res=r)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= TOL
