"""ORACLE tooling: golden vectors for the FPN + heads half of the models whose backbone is un-vendored and unpinned
(configs/models/yololite_n.yaml, yololite_m.yaml: arch YOLOLiteMS over tf_efficientnet_lite*; BASELINE config 5).

Runs the UNMODIFIED reference classes `YOLOLiteMS` / `YOLOLiteMS_CPU` (scripts/model/model_v2.py:77-224 / :250-377) from
/root/reference.  Their constructor asks `timm.create_model(...)` only for (a) `feature_info` (channels / reductions of the
taps) and (b) a module whose forward returns the taps; here that module is a stub that returns PRESET synthetic feature
maps, so everything after `feats = self.backbone(x)` (model_v2.py:195 / :353) -- laterals, nearest upsample-add, dense 3x3 +
BN + SiLU (or DWConvBlock) smoothing, P6, decoupled heads, output layout -- is the reference's own code and arithmetic.

    python oracle/make_golden_features.py          # authoring container only; writes tests/golden/feat_*.npz + feat_kat.json
"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from oracle import model_ref  # noqa: E402


class _StubBackbone(nn.Module):
    """What model_v2.py needs from timm: `.feature_info` entries with num_chs / reduction, and forward -> list of taps."""

    def __init__(self, chs, out_indices):
        super().__init__()
        reds = [2, 4, 8, 16, 32]
        full = [16] + list(chs)                                     # a stride-2 tap in front, like every timm features_only model
        self.feature_info = [dict(num_chs=c, reduction=r, module=f"tap{i}") for i, (c, r) in enumerate(zip(full, reds))]
        self.out_indices = list(out_indices) if out_indices is not None else list(range(5))
        self.preset = None

    def forward(self, x):
        return [self.preset[i] for i in self.out_indices]


def _install_stub_timm(chs):
    m = types.ModuleType("timm")
    m.__version__ = "0.0-feature-stub"
    m.create_model = lambda name, features_only=False, pretrained=False, out_indices=None, **kw: _StubBackbone(chs, out_indices)
    sys.modules["timm"] = m


CASES = [
    # name, model yaml, nc, img, B, p2, p6, anchors, stored-anchor step
    ("feat_yololite_m_p2_128_nc80", "yololite_m", 80, 128, 1, True, False, 1, 2),
    ("feat_yololite_n_p6_128_nc4_a2", "yololite_n", 4, 128, 2, False, True, 2, 1),
    ("feat_yololite_m_192_nc3", "yololite_m", 3, 192, 1, False, False, 1, 2),
]


def main():
    kat = {}
    for name, mdl, nc, img, B, p2, p6, A, step in CASES:
        meta = model_ref.make_meta(mdl, nc, img, use_p2=p2, use_p6=p6, anchors=A)
        chs4 = model_ref.FEATURE_CHANNELS[meta["backbone"]]
        take = 4 if p2 else 3
        chs = chs4[-take:]
        _install_stub_timm(chs4)
        for k in [k for k in sys.modules if k.startswith("scripts.model")]:
            del sys.modules[k]
        from scripts.model.model_v2 import YOLOLiteMS, YOLOLiteMS_CPU
        cfg = meta["config"]["model"]
        cls = YOLOLiteMS if cfg["arch"] == "YOLOLiteMS" else YOLOLiteMS_CPU
        model = cls(backbone=meta["backbone"], num_classes=nc, fpn_channels=cfg["fpn_channels"],
                    num_anchors_per_level=meta["num_anchors_per_level"], pretrained=False, depth_multiple=cfg["depth_multiple"],
                    width_multiple=cfg["width_multiple"], head_depth=cfg["head_depth"], use_p6=p6, use_p2=p2).eval()
        ck = model_ref.synth_checkpoint(meta, seed=11, calib_size=img, feat_chs=chs)
        spec = model_ref.state_spec(meta, feat_chs=chs)
        ref_sd = model.state_dict()
        assert set(ref_sd.keys()) == set(spec.keys()), sorted(set(ref_sd) ^ set(spec))[:6]
        missing, unexpected = model.load_state_dict(ck["state_dict"], strict=True)
        feats = model_ref.synth_features(B, img, chs, seed=5)
        full = [None] * (5 - len(feats)) + list(feats)
        model.backbone.preset = full
        with torch.no_grad():
            outs = model(torch.zeros(B, 3, img, img))
        mine = model_ref.forward_ref(ck["state_dict"], meta, None, feats=feats)
        err = max(float((a - b).abs().max()) for a, b in zip(outs, mine))
        assert err < 2e-4, (name, err)
        flat = [o.reshape(B, -1, o.shape[-1]).numpy() for o in outs]
        store = {f"level{i}": f[:, ::step].copy() for i, f in enumerate(flat)}
        store["shapes"] = np.array([list(o.shape) for o in outs], np.int64)
        store["strides"] = np.array(model.get_strides(), np.int64)
        store["step"] = np.array(step)
        store["sum"] = np.array([float(f.astype(np.float64).sum()) for f in flat])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **store)
        kat[name] = dict(model=mdl, nc=nc, img=img, B=B, p2=p2, p6=p6, anchors=A, seed=11, feat_seed=5, calib=img, chs=list(chs),
                         n_keys=len(ref_sd), oracle_vs_reference_maxabs=err, logit_absmax=float(max(np.abs(f).max() for f in flat)))
        print(name, kat[name])
    with open(os.path.join(OUT, "feat_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)


if __name__ == "__main__":
    main()
