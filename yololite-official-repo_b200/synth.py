"""Random-init checkpoints in the reference's {"state_dict","meta"} format (tools/train.py:62-75), for benchmarks
and demos -- no trained weights ship with the reference and there is no network.

Weights are He-initialised and BatchNorm is the identity (gamma 1, beta 0, mean 0, var 1), which keeps
activations O(1) through the stack without a calibration pass.  Key names / shapes follow the reference module
tree (scripts/model/model_v2.py:250-332) over timm's mobilenetv4_conv_small naming.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional

import torch

from . import packer

MODEL_YAMLS = {   # configs/models/*.yaml
    "edge_n": dict(arch="YOLOLiteMS_CPU", backbone="mobilenetv4_conv_small_050", depth_multiple=0.65, width_multiple=0.60,
                   fpn_channels=160, head_depth=1),
    "edge_s": dict(arch="YOLOLiteMS_CPU", backbone="mobilenetv4_conv_small", depth_multiple=0.90, width_multiple=0.75,
                   fpn_channels=256, head_depth=2),
    "edge_m": dict(arch="YOLOLiteMS_CPU", backbone="mobilenetv4_conv_small", depth_multiple=0.95, width_multiple=0.85,
                   fpn_channels=288, head_depth=2),
    "edge_l": dict(arch="YOLOLiteMS_CPU", backbone="mobilenetv4_conv_small", depth_multiple=1.05, width_multiple=1.00,
                   fpn_channels=320, head_depth=3),
}


# configs/models/yololite_*.yaml: YOLOLiteMS (dense 3x3 + SiLU FPN) over tf_efficientnet_lite*; only their FPN + heads are lowered
# (engine from_features=True).  FEATURE_CHANNELS: channels of the [c2, c3, c4, c5] taps of those timm backbones.
MODEL_YAMLS.update({
    "yololite_n": dict(arch="YOLOLiteMS", backbone="tf_efficientnet_lite0", depth_multiple=1.0, width_multiple=1.0, fpn_channels=196, head_depth=1),
    "yololite_m": dict(arch="YOLOLiteMS", backbone="tf_efficientnet_lite2", depth_multiple=1.0, width_multiple=1.0, fpn_channels=328, head_depth=2),
})
FEATURE_CHANNELS = {"tf_efficientnet_lite0": (24, 40, 112, 320), "tf_efficientnet_lite2": (24, 48, 120, 352)}


def make_meta(model: str = "edge_n", num_classes: int = 80, img_size: int = 640, use_p2: bool = False,
              use_p6: bool = False, anchors: int = 1) -> dict:
    m = dict(MODEL_YAMLS[model])
    m["num_classes"] = num_classes
    n_levels = 3 + int(use_p2) + int(use_p6)
    return {"metric_key": "AP50", "metric_value": -1.0, "names": [f"class_{i}" for i in range(num_classes)],
            "num_classes": num_classes, "img_size": img_size, "arch": m["arch"], "backbone": m["backbone"],
            "num_anchors_per_level": tuple([anchors] * n_levels),
            "config": {"model": m, "training": {"img_size": img_size, "use_p6": use_p6, "use_p2": use_p2}}}


def state_shapes(meta: dict, feat_chs=None) -> "OrderedDict[str, tuple]":
    """feat_chs: channels of the backbone taps when only the FPN + heads are wanted (no backbone keys)."""
    cfg = packer.parse_meta(meta, from_features=feat_chs is not None)
    out: "OrderedDict[str, tuple]" = OrderedDict()

    def bn(p, c):
        for k in ("weight", "bias", "running_mean", "running_var"):
            out[f"{p}.{k}"] = (c,)
        out[f"{p}.num_batches_tracked"] = ()

    if feat_chs is not None:
        feats = list(feat_chs)
    else:
        table, mult, stem = packer.BACKBONES[cfg.backbone]
        out["backbone.conv_stem.weight"] = (stem, 3, 3, 3)
        bn("backbone.bn1", stem)
        cin = stem
        feats = [stem]
        for si, stage in enumerate(table):
            for bi, spec in enumerate(stage):
                key = f"backbone.blocks.{si}.{bi}"
                if spec[0] == "cn":
                    _, k, s, c = spec
                    cout = packer._round_ch(c * mult)
                    out[key + ".conv.weight"] = (cout, cin, k, k)
                    bn(key + ".bn1", cout)
                else:
                    _, ks, km, s, e, c = spec
                    cout, mid = packer._round_ch(c * mult), packer._round_ch(cin * e)
                    if ks:
                        out[key + ".dw_start.conv.weight"] = (cin, 1, ks, ks); bn(key + ".dw_start.bn", cin)
                    out[key + ".pw_exp.conv.weight"] = (mid, cin, 1, 1); bn(key + ".pw_exp.bn", mid)
                    if km:
                        out[key + ".dw_mid.conv.weight"] = (mid, 1, km, km); bn(key + ".dw_mid.bn", mid)
                    out[key + ".pw_proj.conv.weight"] = (cout, mid, 1, 1); bn(key + ".pw_proj.bn", cout)
                cin = cout
                nxt = table[si + 1][0] if si + 1 < len(table) else None
                if bi == len(stage) - 1 and (nxt is None or (nxt[2] if nxt[0] == "cn" else nxt[3]) > 1):
                    feats.append(cout)
    chs = feats[-(4 if cfg.use_p2 else 3):]
    Fc, d, C = cfg.fpn_channels, cfg.depth, cfg.num_classes
    cpu = cfg.arch == "yololitems_cpu"

    def smooth(name):
        for i in range(d):
            if cpu:
                out[f"{name}.block.{4*i}.weight"] = (Fc, 1, 3, 3)
                out[f"{name}.block.{4*i+1}.weight"] = (Fc, Fc, 1, 1)
                bn(f"{name}.block.{4*i+2}", Fc)
            else:
                out[f"{name}.{3*i}.weight"] = (Fc, Fc, 3, 3)
                bn(f"{name}.{3*i+1}", Fc)

    for nm, c in zip((["lateral2"] if cfg.use_p2 else []) + ["lateral3", "lateral4", "lateral5"], chs):
        out[nm + ".weight"] = (Fc, c, 1, 1)
        out[nm + ".bias"] = (Fc,)
    for nm in (["smooth2"] if cfg.use_p2 else []) + ["smooth3", "smooth4", "smooth5"]:
        smooth(nm)
    out["p6_down.weight"] = (Fc, Fc, 3, 3)
    bn("p6_bn", Fc)
    smooth("smooth6")
    for lvl, A in zip(cfg.levels, cfg.anchors):
        h = "head" + lvl[1]
        for i in range(cfg.head_depth):
            out[f"{h}.trunk.{i}.block.0.weight"] = (Fc, 1, 3, 3)
            out[f"{h}.trunk.{i}.block.1.weight"] = (Fc, Fc, 1, 1)
            bn(f"{h}.trunk.{i}.block.2", Fc)
        for nm, n in (("box", A * 4), ("obj", A), ("cls", A * C)):
            out[f"{h}.out.{nm}.weight"] = (n, Fc, 1, 1)
            out[f"{h}.out.{nm}.bias"] = (n,)
    return out


def random_checkpoint(meta: dict, seed: int = 0, obj_bias: Optional[float] = None, head_std: float = 0.1, feat_chs=None) -> dict:
    g = torch.Generator().manual_seed(seed)
    C = packer.parse_meta(meta, from_features=feat_chs is not None).num_classes
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for k, shp in state_shapes(meta, feat_chs).items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(0, dtype=torch.long)
        elif k.endswith(("running_mean",)) or (k.endswith(".bias") and len(shp) == 1 and ".out." not in k):
            sd[k] = torch.zeros(shp)
        elif k.endswith("running_var") or (len(shp) == 1 and k.endswith(".weight")):
            sd[k] = torch.ones(shp)
        elif ".out." in k and k.endswith(".weight"):
            sd[k] = torch.randn(shp, generator=g) * head_std
        elif k.endswith(".out.obj.bias"):                # model_v2.py:7-14
            sd[k] = torch.full(shp, -math.log(99.0) if obj_bias is None else float(obj_bias))
        elif k.endswith(".out.cls.bias"):
            sd[k] = torch.full(shp, -math.log(C) if C > 1 else 0.0)
        elif k.endswith(".out.box.bias"):
            sd[k] = torch.zeros(shp)
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            # gain 2 where a ReLU/SiLU follows; 1 for linear layers (dw_start, DWConvBlock depthwise, laterals);
            # 0.25 for the residual-branch projection so the 10 skip connections do not double the variance each
            linear = (".dw_start." in k or k.startswith("lateral") or ".block.0." in k or ".block.4." in k)
            gain = 0.25 if ".pw_proj." in k else (1.0 if linear else 2.0)
            sd[k] = torch.randn(shp, generator=g) * math.sqrt(gain / fan_in)
    return {"state_dict": sd, "meta": meta}
