"""ORACLE (test infrastructure, not product code).

CPU restatement, in functional PyTorch fp32, of the reference detection forward pass:

  * backbone  -- `timm` `mobilenetv4_conv_small[_050]` `features_only` (un-vendored third-party
                 dependency of the reference: requirements.txt:3 `timm>=0.9`, recorded version 1.0.20 at
                 YoloLite_custom_training.ipynb:101).  Structure restated from the only in-repo pin, the
                 `print(model)` dump at YoloLite_custom_training.ipynb:391-850, plus timm's published
                 channel rounding (`make_divisible`, divisor 8, round_limit 0.9).  Residual rule and feature
                 taps are NOT visible in that dump -> backbone parity is "unpinned" in the
                 strict sense; the FPN/head/decode/NMS parity below IS pinned against the reference code.
  * FPN+heads -- scripts/model/model_v2.py:15-53 (conv_block / DWConvBlock / make_head),
                 :250-377 (YOLOLiteMS_CPU), :77-224 (YOLOLiteMS), :340-350 (_forward_head layout).

Everything here works on a plain ``state_dict`` (same key names as the reference module tree, so the same
checkpoint file loads on both sides) -- there are no nn.Module classes.  Only tests/, bench.py's CPU
baseline leg and __graft_entry__.smoke() may import this file.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5

# --------------------------------------------------------------------------------------------------
# backbone description (YoloLite_custom_training.ipynb:392-850)
# --------------------------------------------------------------------------------------------------
# ("cn", kernel, stride, out_ch)                          ConvBnAct:  conv -> BN -> ReLU, never a skip
# ("uir", dw_start_k, dw_mid_k, stride, expand, out_ch)   UniversalInvertedResidual
_MNV4_CONV_SMALL = [
    [("cn", 3, 2, 32), ("cn", 1, 1, 32)],
    [("cn", 3, 2, 96), ("cn", 1, 1, 64)],
    [("uir", 5, 5, 2, 3.0, 96)] + [("uir", 0, 3, 1, 2.0, 96)] * 4 + [("uir", 3, 0, 1, 4.0, 96)],
    [("uir", 3, 3, 2, 6.0, 128), ("uir", 5, 5, 1, 4.0, 128), ("uir", 0, 5, 1, 4.0, 128),
     ("uir", 0, 5, 1, 3.0, 128), ("uir", 0, 3, 1, 4.0, 128), ("uir", 0, 3, 1, 4.0, 128)],
    [("cn", 1, 1, 960)],
]
_BACKBONES = {
    # name: (stage table, channel multiplier, stem channels)
    "mobilenetv4_conv_small": (_MNV4_CONV_SMALL, 1.0, 32),
    # stem stays at 32 for the 0.5x variant (fix_stem) -- this is what reproduces BENCHMARK.md:353
    # (edge_n 0.553 M params), see SURVEY.md section 8c.
    "mobilenetv4_conv_small_050": (_MNV4_CONV_SMALL, 0.5, 32),
}


def make_divisible(v: float, divisor: int = 8, round_limit: float = 0.9) -> int:
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < round_limit * v:
        new_v += divisor
    return new_v


def backbone_layers(name: str) -> Tuple[List[dict], List[dict]]:
    """Enumerate the backbone as a flat list of conv records plus the feature taps.

    Returns (blocks, feature_info).  Each block is a dict:
      {"key": "blocks.2.0", "type": "cn"|"uir", "cin", "cout", "stride", "convs": [conv records]}
    conv record: {"key", "k", "stride", "groups", "cin", "cout", "bn": key prefix, "act": bool}
    feature_info entries: {"after": block key or "stem", "num_chs", "reduction"}.
    """
    if name not in _BACKBONES:
        raise ValueError(f"oracle has no restatement of timm backbone {name!r}")
    table, mult, stem = _BACKBONES[name]
    blocks: List[dict] = []
    feats: List[dict] = [{"after": "stem", "num_chs": stem, "reduction": 2}]
    cin, red = stem, 2
    for si, stage in enumerate(table):
        for bi, spec in enumerate(stage):
            key = f"blocks.{si}.{bi}"
            if spec[0] == "cn":
                _, k, s, c = spec
                cout = make_divisible(c * mult)
                convs = [dict(key=f"{key}.conv", k=k, stride=s, groups=1, cin=cin, cout=cout,
                              bn=f"{key}.bn1", act=True)]
                blocks.append(dict(key=key, type="cn", cin=cin, cout=cout, stride=s, convs=convs))
            else:
                _, ks, km, s, e, c = spec
                cout = make_divisible(c * mult)
                mid = make_divisible(cin * e)
                convs = []
                if ks:
                    convs.append(dict(key=f"{key}.dw_start.conv", k=ks, stride=(1 if km else s), groups=cin,
                                      cin=cin, cout=cin, bn=f"{key}.dw_start.bn", act=False))
                convs.append(dict(key=f"{key}.pw_exp.conv", k=1, stride=1, groups=1, cin=cin, cout=mid,
                                  bn=f"{key}.pw_exp.bn", act=True))
                if km:
                    convs.append(dict(key=f"{key}.dw_mid.conv", k=km, stride=s, groups=mid, cin=mid, cout=mid,
                                      bn=f"{key}.dw_mid.bn", act=True))
                convs.append(dict(key=f"{key}.pw_proj.conv", k=1, stride=1, groups=1, cin=mid, cout=cout,
                                  bn=f"{key}.pw_proj.bn", act=False))
                blocks.append(dict(key=key, type="uir", cin=cin, cout=cout, stride=s, convs=convs))
            red *= spec[2] if spec[0] == "cn" else spec[3]
            cin = blocks[-1]["cout"]
            blocks[-1]["reduction"] = red
    # timm's builder taps the last block before every stride change, and the very last block.
    flat = blocks
    for i, b in enumerate(flat):
        last = i == len(flat) - 1
        nxt_stride = 1 if last else flat[i + 1]["stride"]
        stage_end = last or flat[i + 1]["key"].split(".")[1] != b["key"].split(".")[1]
        if last or (stage_end and nxt_stride > 1):
            feats.append({"after": b["key"], "num_chs": b["cout"], "reduction": b["reduction"]})
    return blocks, feats


# --------------------------------------------------------------------------------------------------
# model description from a checkpoint "meta" (tools/infer.py:34-77, tools/train.py:62-75)
# --------------------------------------------------------------------------------------------------
def model_cfg_from_meta(meta: dict) -> dict:
    cfg = meta.get("config", {}) or {}
    mcfg = cfg.get("model", {}) or {}
    tcfg = cfg.get("training", {}) or {}
    arch = (meta.get("arch") or mcfg.get("arch") or "YOLOLiteMS").lower()
    if arch not in ("yololitems", "yololitems_cpu"):
        raise ValueError(f"unknown arch {arch}")
    A = tuple(meta.get("num_anchors_per_level") or (1, 1, 1))
    out = dict(
        arch=arch,
        backbone=(meta.get("backbone") or mcfg.get("backbone") or "resnet18"),
        num_classes=int(meta.get("num_classes") or mcfg.get("num_classes") or 80),
        fpn_channels=int(int(mcfg.get("fpn_channels", 128)) * float(mcfg.get("width_multiple", 1.0))),
        depth=max(1, round(2 * float(mcfg.get("depth_multiple", 1.0)))),
        head_depth=int(mcfg.get("head_depth", 1)),
        use_p6=bool(tcfg["use_p6"]),   # hard-indexed in the reference (tools/infer.py:49-50)
        use_p2=bool(tcfg["use_p2"]),
        img_size=int(tcfg.get("img_size", meta.get("img_size", 640))),
    )
    levels = (["p2"] if out["use_p2"] else []) + ["p3", "p4", "p5"] + (["p6"] if out["use_p6"] else [])
    if len(A) >= 3:
        a3, a4, a5 = (int(v) for v in A[:3])
        amap = {"p2": a3, "p3": a3, "p4": a4, "p5": a5, "p6": a5}
    else:
        a = int(A[0]) if len(A) else 1
        amap = {k: a for k in ("p2", "p3", "p4", "p5", "p6")}
    out["levels"] = levels
    out["anchors"] = tuple(amap[l] for l in levels)
    return out


def make_meta(model: str = "edge_n", num_classes: int = 80, img_size: int = 640, use_p2: bool = False,
              use_p6: bool = False, anchors: int = 1, names: Optional[Sequence[str]] = None) -> dict:
    """A checkpoint ``meta`` dict in the layout tools/train.py:62-75 writes, for the model yaml named."""
    yamls = {  # configs/models/*.yaml
        "edge_n": dict(arch="YOLOLiteMS_CPU", backbone="mobilenetv4_conv_small_050", depth_multiple=0.65,
                       width_multiple=0.60, fpn_channels=160, head_depth=1),
        "edge_s": dict(arch="YOLOLiteMS_CPU", backbone="mobilenetv4_conv_small", depth_multiple=0.90,
                       width_multiple=0.75, fpn_channels=256, head_depth=2),
        "edge_m": dict(arch="YOLOLiteMS_CPU", backbone="mobilenetv4_conv_small", depth_multiple=0.95,
                       width_multiple=0.85, fpn_channels=288, head_depth=2),
        "edge_l": dict(arch="YOLOLiteMS_CPU", backbone="mobilenetv4_conv_small", depth_multiple=1.05,
                       width_multiple=1.00, fpn_channels=320, head_depth=3),
        # YOLOLiteMS (dense 3x3 + SiLU FPN) on a backbone the oracle can restate; the reference pairs this
        # arch with tf_efficientnet_* backbones which have no in-repo structural pin.
        "ms_n_mnv4": dict(arch="YOLOLiteMS", backbone="mobilenetv4_conv_small", depth_multiple=1.0,
                          width_multiple=1.0, fpn_channels=196, head_depth=1),
        "ms_m_mnv4": dict(arch="YOLOLiteMS", backbone="mobilenetv4_conv_small", depth_multiple=1.0,
                          width_multiple=1.0, fpn_channels=328, head_depth=2),
        # configs/models/yololite_n.yaml / yololite_m.yaml: the oracle restates their FPN + heads only (forward_ref(feats=...));
        # the tf_efficientnet_lite* backbones are un-vendored timm code with no in-repo structural pin
        "yololite_n": dict(arch="YOLOLiteMS", backbone="tf_efficientnet_lite0", depth_multiple=1.0, width_multiple=1.0,
                           fpn_channels=196, head_depth=1),
        "yololite_m": dict(arch="YOLOLiteMS", backbone="tf_efficientnet_lite2", depth_multiple=1.0, width_multiple=1.0,
                           fpn_channels=328, head_depth=2),
    }
    m = dict(yamls[model])
    m["num_classes"] = num_classes
    n_levels = 3 + int(use_p2) + int(use_p6)
    return {
        "metric_key": "AP50", "metric_value": -1.0,
        "names": list(names) if names else [f"class_{i}" for i in range(num_classes)],
        "num_classes": num_classes, "img_size": img_size, "arch": m["arch"], "backbone": m["backbone"],
        "num_anchors_per_level": tuple([anchors] * n_levels),
        "config": {"model": m, "training": {"img_size": img_size, "use_p6": use_p6, "use_p2": use_p2}},
    }


# channels of the [c2, c3, c4, c5] taps of timm's tf_efficientnet_lite0 / lite2 `features_only` (reductions 4, 8, 16, 32), from
# timm's published arch definitions; used only to give the FPN + head restatement realistic input widths (unpinned)
FEATURE_CHANNELS = {"tf_efficientnet_lite0": (24, 40, 112, 320), "tf_efficientnet_lite2": (24, 48, 120, 352)}


def state_spec(meta: dict, feat_chs: Optional[Sequence[int]] = None) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """Every state_dict entry the reference module tree holds for this meta: key -> (shape, kind).

    kind in {"conv", "dw", "head_w", "bias", "obj_b", "cls_b", "box_b", "bn_w", "bn_b", "bn_rm", "bn_rv",
    "bn_nbt"}.  Mirrors scripts/model/model_v2.py:250-332 (p6_down/p6_bn/smooth6 exist even when unused).
    """
    cfg = model_cfg_from_meta(meta)
    spec: "OrderedDict[str, Tuple[Tuple[int, ...], str]]" = OrderedDict()

    def bn(prefix, c):
        spec[prefix + ".weight"] = ((c,), "bn_w")
        spec[prefix + ".bias"] = ((c,), "bn_b")
        spec[prefix + ".running_mean"] = ((c,), "bn_rm")
        spec[prefix + ".running_var"] = ((c,), "bn_rv")
        spec[prefix + ".num_batches_tracked"] = ((), "bn_nbt")

    take = 4 if cfg["use_p2"] else 3
    if feat_chs is not None:          # FPN + heads only: the backbone (and its parameters) live elsewhere
        chs = list(feat_chs)[-take:]
    else:
        blocks, feats = backbone_layers(cfg["backbone"])
        stem = feats[0]["num_chs"]
        spec["backbone.conv_stem.weight"] = ((stem, 3, 3, 3), "conv")
        bn("backbone.bn1", stem)
        for b in blocks:
            for c in b["convs"]:
                spec[f"backbone.{c['key']}.weight"] = ((c["cout"], c["cin"] // c["groups"], c["k"], c["k"]),
                                                        "dw" if c["groups"] > 1 else "conv")
                bn(f"backbone.{c['bn']}", c["cout"])
        chs = [f["num_chs"] for f in feats[-take:]]
    Fc, d, C = cfg["fpn_channels"], cfg["depth"], cfg["num_classes"]
    cpu = cfg["arch"] == "yololitems_cpu"

    def smooth(name):
        for i in range(d):
            if cpu:
                spec[f"{name}.block.{4*i}.weight"] = ((Fc, 1, 3, 3), "dw")
                spec[f"{name}.block.{4*i+1}.weight"] = ((Fc, Fc, 1, 1), "conv")
                bn(f"{name}.block.{4*i+2}", Fc)
            else:
                spec[f"{name}.{3*i}.weight"] = ((Fc, Fc, 3, 3), "conv")
                bn(f"{name}.{3*i+1}", Fc)

    def head(name, A):
        for i in range(cfg["head_depth"]):
            spec[f"{name}.trunk.{i}.block.0.weight"] = ((Fc, 1, 3, 3), "dw")
            spec[f"{name}.trunk.{i}.block.1.weight"] = ((Fc, Fc, 1, 1), "conv")
            bn(f"{name}.trunk.{i}.block.2", Fc)
        for nm, n, kb in (("box", A * 4, "box_b"), ("obj", A, "obj_b"), ("cls", A * C, "cls_b")):
            spec[f"{name}.out.{nm}.weight"] = ((n, Fc, 1, 1), "head_w")
            spec[f"{name}.out.{nm}.bias"] = ((n,), kb)

    lat = (["lateral2"] if cfg["use_p2"] else []) + ["lateral3", "lateral4", "lateral5"]
    for nm, c in zip(lat, chs):
        spec[f"{nm}.weight"] = ((Fc, c, 1, 1), "conv")
        spec[f"{nm}.bias"] = ((Fc,), "bias")
    for nm in (["smooth2"] if cfg["use_p2"] else []) + ["smooth3", "smooth4", "smooth5"]:
        smooth(nm)
    spec["p6_down.weight"] = ((Fc, Fc, 3, 3), "conv")
    bn("p6_bn", Fc)
    smooth("smooth6")
    for lvl, A in zip(cfg["levels"], cfg["anchors"]):
        head("head" + lvl[1], A)
    return spec


# --------------------------------------------------------------------------------------------------
# functional forward
# --------------------------------------------------------------------------------------------------
class _Ctx:
    """Carries the state dict; in calibration mode BN running stats are (re)written from the batch."""

    def __init__(self, sd: Dict[str, torch.Tensor], calibrate: bool = False):
        self.sd, self.calibrate = sd, calibrate

    def conv(self, x, key, stride=1, groups=1, bias=False):
        w = self.sd[key + ".weight"]
        b = self.sd[key + ".bias"] if bias else None
        return F.conv2d(x, w, b, stride=stride, padding=w.shape[-1] // 2, groups=groups)

    def bn(self, x, key):
        if self.calibrate:
            self.sd[key + ".running_mean"] = x.mean(dim=(0, 2, 3)).detach().clone()
            self.sd[key + ".running_var"] = x.var(dim=(0, 2, 3), unbiased=False).detach().clone() + 1e-3
        return F.batch_norm(x, self.sd[key + ".running_mean"], self.sd[key + ".running_var"],
                            self.sd[key + ".weight"], self.sd[key + ".bias"], False, 0.0, BN_EPS)


def backbone_forward(ctx: _Ctx, x: torch.Tensor, name: str, prefix: str = "backbone.") -> List[torch.Tensor]:
    """All 5 `features_only` taps (reductions 2,4,8,16,32) of mobilenetv4_conv_small[_050]."""
    blocks, feats = backbone_layers(name)
    taps = {f["after"] for f in feats}
    out = []
    x = F.relu(ctx.bn(ctx.conv(x, prefix + "conv_stem", stride=2), prefix + "bn1"))
    if "stem" in taps:
        out.append(x)
    for b in blocks:
        y = x
        for c in b["convs"]:
            y = ctx.bn(ctx.conv(y, prefix + c["key"], stride=c["stride"], groups=c["groups"]), prefix + c["bn"])
            if c["act"]:
                y = F.relu(y)
        if b["type"] == "uir" and b["cin"] == b["cout"] and b["stride"] == 1:
            y = y + x
        x = y
        if b["key"] in taps:
            out.append(x)
    return out


def _dw_block(ctx, x, name, n):          # model_v2.py:23-39
    for i in range(n):
        x = ctx.conv(x, f"{name}.block.{4*i}", groups=x.shape[1])
        x = ctx.conv(x, f"{name}.block.{4*i+1}")
        x = F.relu(ctx.bn(x, f"{name}.block.{4*i+2}"))
    return x


def _dense_block(ctx, x, name, n):       # model_v2.py:15-22
    for i in range(n):
        x = F.silu(ctx.bn(ctx.conv(x, f"{name}.{3*i}"), f"{name}.{3*i+1}"))
    return x


def _head(ctx, p, name, A, C, depth):    # model_v2.py:42-53, :340-350
    for i in range(depth):
        p = _dw_block(ctx, p, f"{name}.trunk.{i}", 1)
    box = ctx.conv(p, f"{name}.out.box", bias=True)
    obj = ctx.conv(p, f"{name}.out.obj", bias=True)
    cls = ctx.conv(p, f"{name}.out.cls", bias=True)
    B, _, H, W = box.shape
    t = torch.cat([box.view(B, A, 4, H, W), obj.view(B, A, 1, H, W), cls.view(B, A, C, H, W)], dim=2)
    return t.permute(0, 1, 3, 4, 2).contiguous()


def forward_ref(sd: Dict[str, torch.Tensor], meta: dict, x: Optional[torch.Tensor], calibrate: bool = False,
                return_feats: bool = False, feats: Optional[Sequence[torch.Tensor]] = None):
    """Reference forward: x [B,3,H,W] fp32 NCHW -> list of [B,A,S,S,5+C] per level (model_v2.py:352-377).  With `feats`
    ([c2,] c3, c4, c5 NCHW, what `self.backbone(x)` returns at model_v2.py:195,353) the backbone is skipped: FPN + heads only."""
    cfg = model_cfg_from_meta(meta)
    ctx = _Ctx(sd, calibrate)
    with torch.no_grad():
        take = 4 if cfg["use_p2"] else 3
        feats = list(feats) if feats is not None else backbone_forward(ctx, x, cfg["backbone"])
        feats = feats[-take:]
        cpu = cfg["arch"] == "yololitems_cpu"
        smooth = _dw_block if cpu else _dense_block
        d, C = cfg["depth"], cfg["num_classes"]

        def up_add(xc, y):
            return F.interpolate(xc, size=y.shape[-2:], mode="nearest") + y

        c5, c4, c3 = feats[-1], feats[-2], feats[-3]
        p5 = smooth(ctx, ctx.conv(c5, "lateral5", bias=True), "smooth5", d)
        p4 = smooth(ctx, up_add(p5, ctx.conv(c4, "lateral4", bias=True)), "smooth4", d)
        p3 = smooth(ctx, up_add(p4, ctx.conv(c3, "lateral3", bias=True)), "smooth3", d)
        pyr = {"p3": p3, "p4": p4, "p5": p5}
        if cfg["use_p2"]:
            pyr["p2"] = smooth(ctx, up_add(p3, ctx.conv(feats[0], "lateral2", bias=True)), "smooth2", d)
        if cfg["use_p6"]:
            t = ctx.bn(ctx.conv(p5, "p6_down", stride=2), "p6_bn")
            pyr["p6"] = smooth(ctx, F.relu(t) if cpu else F.silu(t), "smooth6", d)
        outs = [_head(ctx, pyr[l], "head" + l[1], A, C, cfg["head_depth"])
                for l, A in zip(cfg["levels"], cfg["anchors"])]
    if return_feats:
        return outs, {"c3": c3, "c4": c4, "c5": c5, **pyr}
    return outs


def strides_ref(meta: dict) -> List[int]:
    cfg = model_cfg_from_meta(meta)
    _, feats = backbone_layers(cfg["backbone"])
    base = [f["reduction"] for f in feats[-(4 if cfg["use_p2"] else 3):]]
    return base + ([base[-1] * 2] if cfg["use_p6"] else [])


# --------------------------------------------------------------------------------------------------
# deterministic synthetic checkpoint (no trained weights ship with the reference)
# --------------------------------------------------------------------------------------------------
def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_features(B: int, size: int, chs: Sequence[int], seed: int = 0, reductions: Sequence[int] = (4, 8, 16, 32)):
    """Synthetic backbone taps for an image of `size` px: [B,C_i,size/r_i,size/r_i] NCHW, half-normal (post-activation-like)."""
    reds = list(reductions)[-len(chs):]
    out = []
    for i, (c, r) in enumerate(zip(chs, reds)):
        g = torch.Generator().manual_seed(7000 + 31 * seed + i)
        s = -(-size // r)
        out.append(torch.randn(B, c, s, s, generator=g).abs_().mul_(0.8))
    return out


def synth_checkpoint(meta: dict, seed: int = 0, obj_bias_shift: float = 0.0, calib_size: int = 160,
                     feat_chs: Optional[Sequence[int]] = None) -> dict:
    """{"state_dict", "meta"} with per-key seeded weights and data-calibrated BN statistics.

    Every tensor is drawn from its own generator seeded by crc32(key)^seed, so the result does not depend
    on module construction order.  BN running stats are then set from one forward over a seeded calibration
    batch so activations stay O(1) through the 69-conv stack.
    """
    sd: Dict[str, torch.Tensor] = OrderedDict()
    C = model_cfg_from_meta(meta)["num_classes"]
    for key, (shape, kind) in state_spec(meta, feat_chs).items():
        g = _gen(key, seed)
        if kind in ("conv", "dw"):
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif kind == "head_w":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "box_b":
            t = torch.zeros(shape)
        elif kind == "obj_b":                      # model_v2.py:7-14 init_detect_bias
            t = torch.full(shape, -math.log((1 - 0.01) / 0.01) + obj_bias_shift)
        elif kind == "cls_b":
            t = torch.full(shape, -math.log(C) if C > 1 else 0.0)
        elif kind == "bn_w":
            t = torch.rand(shape, generator=g) * 0.6 + 0.7
        elif kind == "bn_b":
            t = torch.randn(shape, generator=g) * 0.3
        elif kind == "bn_rm":
            t = torch.zeros(shape)
        elif kind == "bn_rv":
            t = torch.ones(shape)
        elif kind == "bn_nbt":
            t = torch.tensor(1, dtype=torch.long)
        else:
            raise AssertionError(kind)
        sd[key] = t.float() if kind != "bn_nbt" else t
    m = dict(meta)
    if feat_chs is not None:
        forward_ref(sd, m, None, calibrate=True, feats=synth_features(2, calib_size, feat_chs, seed=seed + 100))
    else:
        xc = torch.randn(4, 3, calib_size, calib_size, generator=_gen("calib", seed))
        forward_ref(sd, m, xc, calibrate=True)
    # p6 branch statistics when the branch is not part of the graph: leave (0,1).
    return {"state_dict": sd, "meta": meta}


def synth_input(B: int, size: int, seed: int = 0) -> torch.Tensor:
    """Uniform-random RGB through the reference normalisation (tools/infer.py:432-453)."""
    g = torch.Generator().manual_seed(1000 + seed)
    u8 = torch.randint(0, 256, (B, size, size, 3), generator=g, dtype=torch.uint8)
    mean = torch.tensor([0.485, 0.456, 0.406])
    std = torch.tensor([0.229, 0.224, 0.225])
    return ((u8.float() / 255.0 - mean) / std).permute(0, 3, 1, 2).contiguous()
