"""ORACLE (test infrastructure).  Minimal stand-in for the un-vendored `timm` package so that the reference's
scripts/model/model_v2.py can be imported UNMODIFIED from /root/reference (it does `import timm` and calls
`timm.create_model(name, features_only=True, pretrained=..., out_indices=...)` at model_v2.py:94-100,266-272).

Only the parameter *container* lives here (nn.Conv2d / nn.BatchNorm2d registered under timm's state-dict key
names: conv_stem, bn1, blocks.S.B.conv|bn1, blocks.S.B.{dw_start,pw_exp,dw_mid,pw_proj}.{conv,bn});
the arithmetic is oracle.model_ref.backbone_forward.  Used by oracle/make_golden.py in the authoring
container only -- it never travels into the product path.
"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import model_ref as _ref  # noqa: E402

__version__ = "0.0-oracle-shim"


class _Holder(nn.Module):
    pass


class _Features(nn.Module):
    def __init__(self, name, out_indices):
        super().__init__()
        blocks, feats = _ref.backbone_layers(name)
        self._name = name
        self._out = list(out_indices) if out_indices is not None else list(range(len(feats)))
        self.feature_info = [dict(num_chs=f["num_chs"], reduction=f["reduction"], module=f["after"]) for f in feats]
        stem = feats[0]["num_chs"]
        self.conv_stem = nn.Conv2d(3, stem, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(stem)
        self.blocks = _Holder()
        for b in blocks:
            _, si, bi = b["key"].split(".")
            if not hasattr(self.blocks, si):
                setattr(self.blocks, si, _Holder())
            blk = _Holder()
            setattr(getattr(self.blocks, si), bi, blk)
            for c in b["convs"]:
                parts = c["key"].split(".")[3:]            # e.g. ["conv"] or ["dw_start","conv"]
                node = blk
                for p in parts[:-1]:
                    if not hasattr(node, p):
                        setattr(node, p, _Holder())
                    node = getattr(node, p)
                setattr(node, parts[-1], nn.Conv2d(c["cin"], c["cout"], c["k"], c["stride"], c["k"] // 2,
                                                   groups=c["groups"], bias=False))
                bparts = c["bn"].split(".")[3:]
                node = blk
                for p in bparts[:-1]:
                    node = getattr(node, p)
                setattr(node, bparts[-1], nn.BatchNorm2d(c["cout"]))

    def forward(self, x):
        if self.training:
            raise RuntimeError("oracle timm shim supports eval() only")
        sd = {k: v for k, v in self.state_dict().items()}
        feats = _ref.backbone_forward(_ref._Ctx(sd), x, self._name, prefix="")
        return [feats[i] for i in self._out]


def create_model(name, features_only=False, pretrained=False, out_indices=None, **kw):
    if not features_only:
        raise NotImplementedError("oracle timm shim: features_only=True only")
    return _Features(name, out_indices)
