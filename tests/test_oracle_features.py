"""FPN + heads from backbone features (BASELINE config 5 half: YOLOLiteMS, fpn 196 / 328, +P2 / +P6): the oracle restatement and the
packer's from_features program against golden vectors produced by the UNMODIFIED reference classes run on preset feature maps
(oracle/make_golden_features.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
from oracle import model_ref
from program_interp import run_program

with open(os.path.join(GOLDEN, "feat_kat.json")) as _f:
    FEAT_KAT = json.load(_f)
FEAT_CASES = sorted(FEAT_KAT)


def feat_case(name):
    k = FEAT_KAT[name]
    meta = model_ref.make_meta(k["model"], k["nc"], k["img"], use_p2=k["p2"], use_p6=k["p6"], anchors=k["anchors"])
    ck = model_ref.synth_checkpoint(meta, seed=k["seed"], calib_size=k["calib"], feat_chs=k["chs"])
    feats = model_ref.synth_features(k["B"], k["img"], k["chs"], seed=k["feat_seed"])
    return ck, meta, feats, k


@pytest.mark.parametrize("name", FEAT_CASES)
def test_oracle_fpn_heads_match_reference_golden(name):
    ck, meta, feats, k = feat_case(name)
    g = golden(name + ".npz")
    outs = model_ref.forward_ref(ck["state_dict"], meta, None, feats=feats)
    assert [list(o.shape) for o in outs] == g["shapes"].tolist()
    step = int(g["step"])
    for i, o in enumerate(outs):
        f = o.reshape(k["B"], -1, o.shape[-1]).numpy()
        np.testing.assert_allclose(f[:, ::step], g[f"level{i}"], rtol=0, atol=1e-4)
    assert len(ck["state_dict"]) == k["n_keys"]


@pytest.mark.parametrize("name", FEAT_CASES)
def test_from_features_program_reproduces_oracle(name):
    from yololite_b200 import packer
    ck, meta, feats, k = feat_case(name)
    P = packer.lower(ck["state_dict"], meta, from_features=True)
    assert P.feature_channels == k["chs"]
    assert P.strides == golden(name + ".npz")["strides"].tolist()
    want = model_ref.forward_ref(ck["state_dict"], meta, None, feats=feats)
    got = run_program(P, feats)
    for a, b in zip(got, want):
        assert a.shape == b.shape and float((a - b).abs().max()) < 2e-4
    # no backbone op, every external read is a lateral 1x1 conv
    ext = [op for op in P.ops if op["src"] <= -2]
    assert len(ext) == len(k["chs"]) and all(op["k"] == 1 and op["kind"] == 1 for op in ext)


def test_macs_of_config5_fpn_heads_match_survey():
    """yololite_m + P2 @640 nc=80, FPN + heads = 74.717 GMAC (dense 3x3: 65.84 GMAC)."""
    from yololite_b200 import packer
    meta = model_ref.make_meta("yololite_m", 80, 640, use_p2=True)
    chs = model_ref.FEATURE_CHANNELS["tf_efficientnet_lite2"]
    spec = model_ref.state_spec(meta, feat_chs=chs)
    sd = {k: (torch.zeros(s) if kind != "bn_rv" else torch.ones(s)) for k, (s, kind) in spec.items()}
    P = packer.lower(sd, meta, from_features=True, tensor_cores=False)
    px = {4: 160 * 160, 8: 80 * 80, 16: 40 * 40, 32: 20 * 20}
    red = {}
    total = dense = 0
    for op in P.ops:
        r = {-2: 4, -3: 8, -4: 16, -5: 32}[op["src"]] if op["src"] <= -2 else red[op["src"]]
        if op["dst"] >= 0:
            red[op["dst"]] = r
        macs = px[r] * op["cout"] * op["cin"] * op["k"] ** 2 + (px[r] * op["cin"] * op["k2"] ** 2 if op["kind"] == 3 else 0)
        total += macs
        dense += macs if op["k"] == 3 else 0
    assert abs(total / 1e9 - 74.717) < 0.01 and abs(dense / 1e9 - 65.84) < 0.01
