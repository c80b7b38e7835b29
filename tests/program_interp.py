"""Test infrastructure: execute a packer.Program with plain torch CPU ops.

Validates the host-side lowering (BN folding, weight layouts, op wiring, buffer reuse) without a GPU: the
interpreter follows the op semantics documented in include/yololite_b200.h, so if interpreter(program) ==
oracle forward, the only thing left to check on the GPU is that each kernel implements its op.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _act(t, act):
    return F.relu(t) if act == 1 else F.silu(t) if act == 2 else t


def run_program(P, x):
    blob = torch.from_numpy(np.concatenate(P.blob).astype(np.float32))
    bufs = {}
    levels = {}
    for op in P.ops:
        cin, cout, k, s = op["cin"], op["cout"], op["k"], op["stride"]
        # NCHW tensors inside the interpreter; x is the image, or the list of backbone features for a from_features program
        src = x[-op["src"] - 2] if op["src"] <= -2 else x if op["src"] < 0 else bufs[op["src"]]
        assert src.shape[1] == cin, (op, src.shape)
        bias = None if op["b_off"] < 0 else blob[op["b_off"]:op["b_off"] + cout]
        if op["kind"] == 4:       # fused stem (3x3 s2, 32 ch, ReLU) -> 3x3 s2 conv
            sc = op["k2"]
            ws = blob[op["w2_off"]:op["w2_off"] + 27 * sc].reshape(3, 3, 3, sc).permute(3, 2, 0, 1)
            bs = blob[op["w2_off"] + 27 * sc:op["w2_off"] + 28 * sc]
            mid = F.relu(F.conv2d(src, ws, bs, stride=2, padding=1))
            ld = (cout + 3) // 4 * 4
            w = blob[op["w_off"]:op["w_off"] + 9 * sc * ld].reshape(9 * sc, ld)[:, :cout].reshape(3, 3, sc, cout).permute(3, 2, 0, 1)
            y = F.conv2d(mid, w, bias, stride=s, padding=1)
            if op.get("b2_off", -1) >= 0:     # fused pointwise conv (same width) after conv2: conv2's act, then 1x1 + bias, act2
                y = _act(y, op["act"])
                wq = blob[op["b2_off"]:op["b2_off"] + cout * cout].reshape(cout, cout).t().reshape(cout, cout, 1, 1)
                bq = blob[op["b2_off"] + cout * cout:op["b2_off"] + cout * cout + cout]
                y = _act(F.conv2d(y, wq, bq), op.get("act2", 0))
                if op["dst"] >= 0:
                    bufs[op["dst"]] = y
                continue
        elif op["kind"] == 0:
            w = blob[op["w_off"]:op["w_off"] + k * k * cin * cout].reshape(k, k, cin, cout).permute(3, 2, 0, 1)
            y = F.conv2d(src, w, bias, stride=s, padding=k // 2)
        elif op["kind"] in (1, 3):
            if op["kind"] == 3:
                k2 = op["k2"]
                w2 = blob[op["w2_off"]:op["w2_off"] + k2 * k2 * cin].reshape(k2, k2, 1, cin).permute(3, 2, 0, 1)
                b2 = None if op.get("b2_off", -1) < 0 else blob[op["b2_off"]:op["b2_off"] + cin]
                src = _act(F.conv2d(src, w2, b2, stride=max(1, op.get("stride2", 0)), padding=k2 // 2, groups=cin), op.get("act2", 0))
            ld = (cout + 3) // 4 * 4
            w = blob[op["w_off"]:op["w_off"] + k * k * cin * ld].reshape(k * k * cin, ld)[:, :cout]
            w = w.reshape(k, k, cin, cout).permute(3, 2, 0, 1)
            y = F.conv2d(src, w, bias, stride=s, padding=k // 2)
        elif op["kind"] == 2:
            w = blob[op["w_off"]:op["w_off"] + k * k * cin].reshape(k, k, 1, cin).permute(3, 2, 0, 1)
            y = F.conv2d(src, w, bias, stride=s, padding=k // 2, groups=cin)
        else:
            raise AssertionError(op)
        if op["res"] >= 0:
            y = y + bufs[op["res"]]
        if op["up"] >= 0:
            y = y + F.interpolate(bufs[op["up"]], size=y.shape[-2:], mode="nearest")
        y = _act(y, op["act"])
        if op["dst"] >= 0:
            bufs[op["dst"]] = y
        else:
            A = op["anchors"]
            B, _, H, W = y.shape
            levels[-op["dst"] - 1] = y.view(B, A, cout // A, H, W).permute(0, 1, 3, 4, 2).contiguous()
    return [levels[i] for i in range(len(levels))]
