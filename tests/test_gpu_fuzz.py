"""Randomised shapes through yl_run_op (tcgen05 path forced) against plain PyTorch fp32 convolutions: partial tiles, tiny
images, channel counts that are not multiples of 32, N chunking, streamed weights, residual / upsample epilogues, stride-2
depthwise.  Same 2e-4 absolute tolerance as test_gpu_ops.py."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from test_gpu_ops import TOL, _ref, _run

pytestmark = pytest.mark.gpu
N_CASES = int(os.environ.get("YL_FUZZ_N", "24"))          # per test; raise for a longer hunt
SEED0 = int(os.environ.get("YL_FUZZ_SEED", "0"))


def _cases(seed, n):
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        out.append(dict(cin=int(rs.choice([16, 32, 48, 64, 96, 128, 192, 244, 256, 384, 480])),
                        cout=int(rs.choice([16, 32, 48, 64, 85, 96, 128, 192, 244, 256, 288])),
                        h=int(rs.randint(1, 42)), w=int(rs.randint(1, 42)), b=int(rs.randint(1, 4)),
                        act=int(rs.randint(0, 3)), extra=int(rs.randint(0, 3)), k2=int(rs.choice([3, 5])),
                        s2=int(rs.choice([1, 1, 2])), seed=int(rs.randint(1 << 30))))
    return out


@pytest.mark.parametrize("c", _cases(1 + SEED0, N_CASES), ids=lambda c: f"pw{c['cin']}-{c['cout']}-{c['h']}x{c['w']}b{c['b']}e{c['extra']}")
def test_pointwise_random(c):
    g = torch.Generator().manual_seed(c["seed"])
    x = torch.randn(c["b"], c["h"], c["w"], c["cin"], generator=g)
    w = torch.randn(c["cout"], c["cin"], 1, 1, generator=g) / c["cin"] ** 0.5
    b = torch.randn(c["cout"], generator=g)
    res = torch.randn(c["b"], c["h"], c["w"], c["cout"], generator=g) if c["extra"] == 1 else None
    up = torch.randn(c["b"], (c["h"] + 1) // 2, (c["w"] + 1) // 2, c["cout"], generator=g) if c["extra"] == 2 else None
    anchors = 1 if c["cout"] == 85 else 0
    got = _run(1, x, w, b, 1, 1, c["act"], res=res, up=up, anchors=anchors, use_tc=2)
    want = _ref(x.permute(0, 3, 1, 2), w, b, 1, 1, c["act"], res=res, up=up)
    assert float((got - want).abs().max()) <= TOL


@pytest.mark.parametrize("c", _cases(2 + SEED0, N_CASES), ids=lambda c: f"dw{c['cin']}-{c['cout']}-k{c['k2']}s{c['s2']}-{c['h']}x{c['w']}b{c['b']}")
def test_fused_dw_pw_random(c):
    cout = c["cout"] if c["cout"] % 4 == 0 else 96              # the fused kernel's vector epilogue needs N % 4 == 0
    g = torch.Generator().manual_seed(c["seed"])
    x = torch.randn(c["b"], c["h"], c["w"], c["cin"], generator=g)
    wd = torch.randn(c["cin"], 1, c["k2"], c["k2"], generator=g) / c["k2"]
    bd = torch.randn(c["cin"], generator=g) * 0.5
    wp = torch.randn(cout, c["cin"], 1, 1, generator=g) / c["cin"] ** 0.5
    b = torch.randn(cout, generator=g)
    mid = F.conv2d(x.permute(0, 3, 1, 2), wd, bd, stride=c["s2"], padding=c["k2"] // 2, groups=c["cin"])
    act2 = c["act"] % 2
    mid = F.relu(mid) if act2 else mid
    res = torch.randn(c["b"], mid.shape[2], mid.shape[3], cout, generator=g) if (c["extra"] == 1 and c["s2"] == 1) else None
    got = _run(3, x, wp, b, 1, 1, c["act"], w2=wd, use_tc=2, b2=bd, act2=act2, stride2=c["s2"], res=res)
    want = _ref(mid, wp, b, 1, 1, c["act"], res=res)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= TOL


@pytest.mark.parametrize("c", _cases(3 + SEED0, max(8, N_CASES // 3)), ids=lambda c: f"d3-{c['cin']}-{c['cout']}-s{c['s2']}-{c['h']}x{c['w']}b{c['b']}")
def test_dense3x3_random(c):
    cin = min(c["cin"], 96)
    cout = c["cout"] if c["cout"] % 4 == 0 else 48
    g = torch.Generator().manual_seed(c["seed"])
    x = torch.randn(c["b"], c["h"] + 2, c["w"] + 2, cin, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    b = torch.randn(cout, generator=g)
    got = _run(1, x, w, b, 3, c["s2"], c["act"], use_tc=2)
    want = _ref(x.permute(0, 3, 1, 2), w, b, 3, c["s2"], c["act"])
    assert float((got - want).abs().max()) <= TOL
