"""Oracle forward restatement vs golden vectors produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import math

import numpy as np
import pytest
import torch

from conftest import FWD_CASES, case_ckpt, golden, kat
from oracle import model_ref


@pytest.mark.parametrize("name", FWD_CASES)
def test_forward_matches_reference_golden(name):
    ckpt, k = case_ckpt(name)
    g = golden(name + ".npz")
    x = model_ref.synth_input(k["B"], k["img"], seed=k["input_seed"])
    outs = model_ref.forward_ref(ckpt["state_dict"], ckpt["meta"], x)
    assert [list(o.shape) for o in outs] == g["shapes"].tolist()
    assert model_ref.strides_ref(ckpt["meta"]) == g["strides"].tolist()
    step = int(g["step"])
    for i, o in enumerate(outs):
        f = o.reshape(k["B"], -1, o.shape[-1]).numpy()
        np.testing.assert_allclose(f[:, ::step], g[f"level{i}"], rtol=0, atol=1e-4)
        assert abs(float(f.astype(np.float64).sum()) - g["sum"][i]) < 1e-2 + 1e-6 * g["abssum"][i]
    assert len(ckpt["state_dict"]) == k["n_keys"]


def test_param_counts_match_published():
    # BENCHMARK.md:353-355: edge_n 0.553 M, edge_s 2.359 M, edge_m 2.950 M
    assert kat()["params_edge_n"]["n_params"] == 552408
    assert kat()["params_edge_s"]["n_params"] == 2359736
    assert kat()["params_edge_m"]["n_params"] == 2948948
    for mdl, want in (("edge_n", 552408), ("edge_s", 2359736), ("edge_m", 2948948)):
        spec = model_ref.state_spec(model_ref.make_meta(mdl, 3, 640))
        n = sum(int(np.prod(s)) for _, (s, kind) in spec.items() if not kind.startswith("bn_r") and kind != "bn_nbt")
        assert n == want
    assert len(model_ref.state_spec(model_ref.make_meta("edge_n", 3, 640))) == 349
    assert len(model_ref.state_spec(model_ref.make_meta("edge_m", 3, 640))) == 398


def test_shapes_and_strides_known_answers():
    # edge_n 640 -> 80/40/20; +P2 @320 -> 80/40/20/10; +P6 @640 -> 80/40/20/10
    for kw, img, want, strides in ((dict(), 640, [80, 40, 20], [8, 16, 32]),
                                   (dict(use_p2=True), 320, [80, 40, 20, 10], [4, 8, 16, 32]),
                                   (dict(use_p6=True), 640, [80, 40, 20, 10], [8, 16, 32, 64])):
        meta = model_ref.make_meta("edge_n", 3, img, **kw)
        ck = model_ref.synth_checkpoint(meta, seed=1, calib_size=64)
        outs = model_ref.forward_ref(ck["state_dict"], meta, torch.zeros(1, 3, img, img))
        assert [o.shape[2] for o in outs] == want
        assert all(o.shape == (1, 1, s, s, 8) for o, s in zip(outs, want))
        assert model_ref.strides_ref(meta) == strides


def test_head_bias_init_known_answers():
    # model_v2.py:7-14: obj = -log(99), cls = -log(C), box = 0
    ck = model_ref.synth_checkpoint(model_ref.make_meta("edge_n", 80, 64), seed=0, calib_size=64)
    sd = ck["state_dict"]
    assert torch.allclose(sd["head3.out.obj.bias"], torch.tensor(-4.59512), atol=1e-5)
    assert torch.allclose(sd["head4.out.cls.bias"], torch.tensor(-math.log(80.0)), atol=1e-6)
    assert float(sd["head5.out.box.bias"].abs().max()) == 0.0


def test_meta_errors_match_reference():
    meta = model_ref.make_meta("edge_n", 3, 64)
    bad = dict(meta); bad["arch"] = "nope"
    with pytest.raises(ValueError):
        model_ref.model_cfg_from_meta(bad)
    bad = dict(meta); bad["config"] = {"model": meta["config"]["model"], "training": {"img_size": 64}}
    with pytest.raises(KeyError):
        model_ref.model_cfg_from_meta(bad)
