#!/usr/bin/env python
"""Micro-benchmark / ncu target: one conv op through yl_run_op on synthetic NHWC data.

    python scripts/bench_op.py --kind dwpw --cin 96 --cout 96 --hw 80 --batch 64 --tc 1 --iters 20
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yololite_b200 import _lib as L, packer  # noqa: E402

KINDS = {"stem": 0, "conv": 1, "dw": 2, "dwpw": 3, "stem2": 4}


def build(kind, cin, cout, k, stride, act, up, res, k2=3, act2=0, stride2=0, tap=0):
    g = np.random.RandomState(0)
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
    op.kind, op.k, op.stride, op.act, op.anchors = KINDS[kind], k, stride, act, 0
    op.src, op.dst, op.res, op.up = 0, 1, (2 if res else -1), (3 if up else -1)
    op.k2, op.w2_off, op.wt_off, op.w3_off, op.b2_off, op.act2 = 0, -1, -1, -1, -1, 0
    op.cin, op.cout = cin, cout
    if kind == "dw":
        op.w_off = add(g.randn(k * k, cin) / k)
    elif kind == "stem":
        op.w_off = add(g.randn(27, cout) / 5)
    elif kind == "stem2":
        op.cin, op.k, op.stride, op.k2, op.src = 3, 3, 2, 32, -1
        wm = np.zeros((288, (cout + 3) // 4 * 4))
        wm[:, :cout] = g.randn(288, cout) / 17
        op.w_off = add(wm)
        op.wt_off = add(packer.tc_image(wm, cout))
        wsm = g.randn(27, 32) / 5
        bsv = g.randn(32) * 0.3
        op.w3_off = add(packer.stem2_image(wm, cout, wsm, bsv))
        op.w2_off = add(np.concatenate([wsm.reshape(-1), bsv]))
    else:
        kk = 1 if kind == "dwpw" else k
        wm = np.zeros((kk * kk * cin, (cout + 3) // 4 * 4))
        wm[:, :cout] = g.randn(kk * kk * cin, cout) / np.sqrt(kk * kk * cin)
        op.w_off = add(wm)
        if tap:
            op.wt_off = add(packer.tc_image(packer.tap_padded(wm, kk * kk, cin), cout)); op.wt_layout = 1
        else:
            op.wt_off = add(packer.tc_image(wm, cout))
        if kind == "dwpw":
            op.k, op.k2, op.act2, op.stride2 = 1, k2, act2, stride2
            op.w2_off = add(g.randn(k2 * k2, cin) / k2)
            op.b2_off = add(packer._pad4(g.randn(cin) * 0.3))
    op.b_off = add(packer._pad4(g.randn(cout)))
    return op, torch.from_numpy(np.concatenate(blob)).cuda()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="conv", choices=list(KINDS))
    ap.add_argument("--cin", type=int, default=96)
    ap.add_argument("--cout", type=int, default=96)
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--hw", type=int, default=80)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--act", type=int, default=1)
    ap.add_argument("--up", type=int, default=0)
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--tc", type=int, default=1)
    ap.add_argument("--k2", type=int, default=3)
    ap.add_argument("--act2", type=int, default=0)
    ap.add_argument("--stride2", type=int, default=0)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--tap", type=int, default=0, help="per-tap padded weight image (TMA-fed dense conv)")
    a = ap.parse_args()
    lib = L.lib()
    op, blob = build(a.kind, a.cin, a.cout, a.k, a.stride, a.act, a.up, a.res, a.k2, a.act2, a.stride2, a.tap)
    B, H = a.batch, a.hw
    k = op.k
    ho = (H + 2 * (k // 2) - k) // a.stride + 1
    if a.kind == "dwpw" and a.stride2 == 2:
        ho = (H + 2 * (a.k2 // 2) - a.k2) // 2 + 1
    if a.kind == "stem2":
        ho = ((H + 2 - 3) // 2 + 1 + 2 - 3) // 2 + 1
    x = torch.randn((B, 3, H, H) if a.kind in ("stem", "stem2") else (B, H, H, a.cin), device="cuda")
    out = torch.empty((B, ho, ho, a.cout), device="cuda")
    res = torch.randn_like(out) if a.res else None
    up = torch.randn((B, (ho + 1) // 2, (ho + 1) // 2, a.cout), device="cuda") if a.up else None

    def run():
        L.check(lib.yl_run_op(ctypes.byref(op), blob.data_ptr(), x.data_ptr(), res.data_ptr() if res is not None else None,
                              up.data_ptr() if up is not None else None, out.data_ptr(), B, H, H,
                              up.shape[1] if up is not None else 0, up.shape[2] if up is not None else 0, a.tc, None))
    for _ in range(int(os.environ.get("YL_BENCH_OP_WARMUP", "3"))):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    nbytes = 4 * (x.numel() + out.numel() + (res.numel() if res is not None else 0) + (up.numel() if up is not None else 0))
    macs = B * ho * ho * a.cout * a.cin * a.k * a.k if a.kind == "conv" else 0
    print(json.dumps({"TFLOPs_useful": 2 * macs / ms / 1e9, "kind": a.kind, "cin": a.cin, "cout": a.cout, "k": a.k, "stride": a.stride, "hw": H, "B": B, "tc": a.tc,
                      "up": a.up, "res": a.res, "k2": a.k2, "ms": ms, "GBps": nbytes / ms / 1e6}))


if __name__ == "__main__":
    main()
