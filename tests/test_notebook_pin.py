"""Pin the oracle's and the packer's backbone / FPN / head tables to a REFERENCE-HELD fixture.

tests/golden/notebook_model_dump.txt is the verbatim `print(model)` of the edge_s model (YOLOLiteMS_CPU over timm 1.0.20's
mobilenetv4_conv_small, fpn 192, depth 2, head_depth 2, P6, 13 classes) that the reference's own training notebook
recorded (YoloLite_custom_training.ipynb:391-1006; extracted by oracle/extract_notebook_dump.py).  Every Conv2d
(in / out / kernel / stride / padding / groups / bias), every BatchNorm (channels, eps) and every activation slot in it is
compared with oracle.model_ref (the parity oracle) and with the product's packer tables.

What a module dump cannot show, and therefore stays unpinned: the residual add of
UniversalInvertedResidual (in == out and stride 1) and which tensors `features_only` returns (block outputs at the last
block before every stride change).  The lateral in-channels 64 / 96 / 960 do pin WHICH blocks are tapped.
"""
import os
import re
from collections import OrderedDict

import pytest

from conftest import GOLDEN
from oracle import model_ref

DUMP = os.path.join(GOLDEN, "notebook_model_dump.txt")


def parse_dump(path=DUMP):
    """-> OrderedDict: dotted module path -> (type name, argument string incl. the continuation line of BatchNormAct2d)."""
    mods = OrderedDict()
    stack = []                                    # (indent, name)
    last = None
    with open(path) as f:
        for raw in f:
            if raw.startswith("#") or not raw.strip():
                continue
            line = raw.rstrip("\n")
            ind = len(line) - len(line.lstrip(" "))
            s = line.strip()
            m = re.match(r"^\((\w+)\): (\w+)\((.*)$", s)
            if m:
                while stack and stack[-1][0] >= ind:
                    stack.pop()
                name, typ, rest = m.groups()
                path_ = ".".join([n for _, n in stack] + [name])
                closed = rest.endswith(")")
                mods[path_] = [typ, rest[:-1] if closed else rest]
                last = path_
                if not closed:
                    stack.append((ind, name))
            elif s == ")":
                while stack and stack[-1][0] >= ind:
                    stack.pop()
            elif s.startswith("YOLOLiteMS_CPU("):
                mods[""] = ["YOLOLiteMS_CPU", ""]
            else:                                  # continuation: "32, eps=1e-05, momentum=0.1, ..."
                mods[last][1] += s
    return OrderedDict((k, tuple(v)) for k, v in mods.items())


def conv_args(argstr):
    m = re.match(r"^(\d+), (\d+), kernel_size=\((\d+), (\d+)\), stride=\((\d+), (\d+)\)(.*)$", argstr)
    assert m, argstr
    cin, cout, kh, kw, sh, sw = (int(v) for v in m.groups()[:6])
    rest = m.group(7)
    pad = re.search(r"padding=\((\d+), (\d+)\)", rest)
    grp = re.search(r"groups=(\d+)", rest)
    assert kh == kw and sh == sw
    return dict(cin=cin, cout=cout, k=kh, stride=sh, pad=int(pad.group(1)) if pad else 0,
                groups=int(grp.group(1)) if grp else 1, bias="bias=False" not in rest)


def bn_args(argstr):
    m = re.match(r"^(\d+), eps=([0-9.e+-]+), momentum=0.1, affine=True, track_running_stats=True", argstr)
    assert m, argstr
    return int(m.group(1)), float(m.group(2))


@pytest.fixture(scope="module")
def dump():
    return parse_dump()


def test_dump_is_the_edge_s_model(dump):
    assert dump[""][0] == "YOLOLiteMS_CPU"
    assert dump["backbone"][0] == "MobileNetV3Features"           # timm's features_only wrapper for mobilenetv4_*
    assert sum(1 for t, _ in dump.values() if t == "Conv2d") == 45 + 3 + 4 * 2 * 2 + 1 + 4 * (2 * 2 + 3)   # backbone, laterals, smooth3-6, p6_down, heads


def test_backbone_convs_bns_and_activations_match_oracle_table(dump):
    blocks, feats = model_ref.backbone_layers("mobilenetv4_conv_small")
    # stem: 3x3 s2, symmetric padding 1 (no TF-"SAME"), BN eps 1e-5, ReLU
    assert conv_args(dump["backbone.conv_stem"][1]) == dict(cin=3, cout=32, k=3, stride=2, pad=1, groups=1, bias=False)
    assert bn_args(dump["backbone.bn1"][1]) == (32, model_ref.BN_EPS)
    assert dump["backbone.act1"][0] == "ReLU"
    seen = {"backbone.conv_stem"}
    for b in blocks:
        bpath = "backbone." + b["key"]
        assert dump[bpath][0] == ("ConvBnAct" if b["type"] == "cn" else "UniversalInvertedResidual"), bpath
        for c in b["convs"]:
            p = "backbone." + c["key"]
            got = conv_args(dump[p][1])
            assert got == dict(cin=c["cin"], cout=c["cout"], k=c["k"], stride=c["stride"], pad=c["k"] // 2, groups=c["groups"],
                               bias=False), (p, got)
            bnp = "backbone." + c["bn"]
            assert dump[bnp][0] == "BatchNormAct2d" and bn_args(dump[bnp][1]) == (c["cout"], model_ref.BN_EPS), bnp
            assert dump[bnp + ".act"][0] == ("ReLU" if c["act"] else "Identity"), bnp
            assert dump[bnp + ".drop"][0] == "Identity"
            seen.add(p)
        # nothing else in the block computes: squeeze-excite, dw_end, layer scale, drop path, anti-aliasing are Identity,
        # and a UIR variant without dw_start / dw_mid has no such child at all
        for extra in ("se", "dw_end", "layer_scale", "drop_path", "aa"):
            if bpath + "." + extra in dump:
                assert dump[bpath + "." + extra][0] == "Identity", (bpath, extra)
        if b["type"] == "uir":
            have = {n for n in ("dw_start", "pw_exp", "dw_mid", "pw_proj") if dump.get(f"{bpath}.{n}", ("",))[0] == "ConvNormAct"}
            want = {c["key"].split(".")[3] for c in b["convs"]}
            assert have == want, bpath
    in_dump = {p for p, (t, _) in dump.items() if t == "Conv2d" and p.startswith("backbone.")}
    assert in_dump == seen                                        # no conv missing, none extra
    # stage structure: 5 stages with 2/2/6/6/1 blocks
    assert [sum(1 for b in blocks if b["key"].startswith(f"blocks.{s}.")) for s in range(5)] == [2, 2, 6, 6, 1]


def test_feature_taps_are_consistent_with_lateral_in_channels(dump):
    _, feats = model_ref.backbone_layers("mobilenetv4_conv_small")
    # model_v2.py:69-74: the FPN takes the last three feature_info entries; their channel counts are the laterals' inputs
    c3, c4, c5 = feats[-3:]
    assert (c3["after"], c4["after"], c5["after"]) == ("blocks.1.1", "blocks.2.5", "blocks.4.0")
    assert (c3["reduction"], c4["reduction"], c5["reduction"]) == (8, 16, 32)
    for name, f in (("lateral3", c3), ("lateral4", c4), ("lateral5", c5)):
        got = conv_args(dump[name][1])
        assert got == dict(cin=f["num_chs"], cout=192, k=1, stride=1, pad=0, groups=1, bias=True), name
    assert [f["num_chs"] for f in feats[-3:]] == [64, 96, 960]


def _edge_s_meta():
    return model_ref.make_meta("edge_s", 13, 640, use_p6=True)


def test_fpn_and_heads_match_oracle_state_spec(dump):
    """Every parameter the oracle's state_spec lists for edge_s (+P6, 13 classes) exists in the dump with that shape, and
    the dump holds no parameterised module the spec does not know."""
    spec = model_ref.state_spec(_edge_s_meta())
    want_w = {k[:-len(".weight")]: s for k, (s, kind) in spec.items() if k.endswith(".weight") and kind in ("conv", "dw", "head_w")}
    got_w = {}
    for p, (t, a) in dump.items():
        if t == "Conv2d":
            c = conv_args(a)
            got_w[p] = (c["cout"], c["cin"] // c["groups"], c["k"], c["k"])
            assert c["bias"] == ((p + ".bias") in spec), p
            assert c["pad"] == c["k"] // 2
    assert got_w == {k: tuple(v) for k, v in want_w.items()}
    want_bn = {k[:-len(".running_mean")]: s[0] for k, (s, kind) in spec.items() if kind == "bn_rm"}
    got_bn = {p: bn_args(a) for p, (t, a) in dump.items() if t in ("BatchNorm2d", "BatchNormAct2d")}
    assert {p: c for p, (c, _) in got_bn.items()} == want_bn
    assert all(eps == model_ref.BN_EPS for _, eps in got_bn.values())
    # DWConvBlock (model_v2.py:23-39) = [dw3x3, pw1x1, BN, ReLU] x n; smooth blocks have n = depth = 2, trunk blocks n = 1
    for name, n in [(f"smooth{l}", 2) for l in (3, 4, 5, 6)] + [(f"head{l}.trunk.{i}", 1) for l in (3, 4, 5, 6) for i in (0, 1)]:
        assert dump[name][0] == "DWConvBlock"
        kinds = [dump[f"{name}.block.{j}"][0] for j in range(4 * n)]
        assert kinds == ["Conv2d", "Conv2d", "BatchNorm2d", "ReLU"] * n, name
        assert f"{name}.block.{4 * n}" not in dump
        for i in range(n):
            assert conv_args(dump[f"{name}.block.{4 * i}"][1])["groups"] == 192 and conv_args(dump[f"{name}.block.{4 * i + 1}"][1])["k"] == 1
    assert conv_args(dump["p6_down"][1]) == dict(cin=192, cout=192, k=3, stride=2, pad=1, groups=1, bias=False)
    assert dump["p6_act"][0] == "ReLU"
    for l in (3, 4, 5, 6):
        assert [conv_args(dump[f"head{l}.out.{n}"][1])["cout"] for n in ("box", "obj", "cls")] == [4, 1, 13]


def test_product_packer_tables_match_the_dump(dump):
    """The PRODUCT's own tables (packer.BACKBONES / synth.state_shapes / the lowered op list) against the same fixture."""
    from yololite_b200 import packer, synth
    meta = synth.make_meta("edge_s", 13, 640, use_p6=True)
    shapes = synth.state_shapes(meta)
    for p, (t, a) in dump.items():
        if t == "Conv2d":
            c = conv_args(a)
            assert tuple(shapes[p + ".weight"]) == (c["cout"], c["cin"] // c["groups"], c["k"], c["k"]), p
            assert ((p + ".bias") in shapes) == c["bias"], p
    n_conv = sum(1 for k in shapes if k.endswith(".weight") and len(shapes[k]) == 4)
    assert n_conv == sum(1 for t, _ in dump.values() if t == "Conv2d")
    # geometry of the lowered program with every fusion off: one op per conv, in execution order
    ck = synth.random_checkpoint(meta, seed=0)
    P = packer.lower(ck["state_dict"], ck["meta"], fuse_dwpw=False, fuse_stem=False, fuse_uir=False, tensor_cores=False)
    blocks, _ = model_ref.backbone_layers("mobilenetv4_conv_small")
    order = ["backbone.conv_stem"] + ["backbone." + c["key"] for b in blocks for c in b["convs"]]
    ops = P.ops[:len(order)]
    for op, p in zip(ops, order):
        c = conv_args(dump[p][1])
        assert (op["cin"], op["cout"], op["k"], op["stride"]) == (c["cin"], c["cout"], c["k"], c["stride"]), p
        assert (op["kind"] == 2) == (c["groups"] > 1), p            # YL_OP_DW exactly for the depthwise convs
    # residual adds sit exactly on the UIR blocks with in == out and stride 1 (timm rule; not visible in the dump)
    res_ops = [i for i, op in enumerate(ops) if op["res"] >= 0]
    want_res = [b["key"] for b in blocks if b["type"] == "uir" and b["cin"] == b["cout"] and b["stride"] == 1]
    assert len(res_ops) == len(want_res) == 10
