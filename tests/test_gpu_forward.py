"""Parity of the CUDA forward (through the C ABI) against the oracle and the reference-generated golden vectors.

Tolerance: 1e-3 absolute on logits (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from conftest import FWD_CASES, case_ckpt, golden, synth_ckpt
from oracle import model_ref

pytestmark = pytest.mark.gpu


def test_forward_u8_non_square_and_unsupported_shapes():
    """Non-square inputs (even H, W % 16 == 0) take the uint8 entry too; shapes it cannot take are refused with ValueError
    so that callers fall back to preprocess_batch + forward."""
    import yololite_b200 as y
    from conftest import synth_ckpt
    ck = synth_ckpt("edge_n", 3, 64)
    eng = y.YoloLiteB200(ck["state_dict"], ck["meta"], device="cuda:0")
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda")
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda")
    for (H, W) in ((96, 64), (34, 48), (64, 176)):
        img = torch.randint(0, 256, (2, H, W, 3), dtype=torch.uint8, device="cuda", generator=torch.Generator(device="cuda").manual_seed(H + W))
        x = ((img.flip(-1).float() / 255.0 - mean) / std).permute(0, 3, 1, 2).contiguous()       # BGR -> RGB, normalise, CHW
        for g, w in zip(eng.forward_u8(img), eng(x)):
            assert float((g - w).abs().max()) <= 2e-4
    assert not eng.supports_u8(33, 48) and not eng.supports_u8(64, 40)
    with pytest.raises(ValueError):
        eng.forward_u8(torch.zeros((1, 64, 40, 3), dtype=torch.uint8, device="cuda"))


@pytest.mark.parametrize("model,nc,S", [("edge_n", 3, 64), ("edge_n", 80, 320), ("edge_n", 5, 96)])
def test_forward_u8_matches_preprocess_plus_forward(model, nc, S):
    """yl_forward_u8 (normalisation folded into the fused stem kernel, one bf16 split for the integer pixels) against
    yl_preprocess_batch + yl_forward on the same uint8 BGR images, and against the CPU oracle on the normalised tensor."""
    import yololite_b200 as y
    from conftest import synth_ckpt
    from oracle import model_ref, pre_ref
    ck = synth_ckpt(model, nc, S)
    eng = y.YoloLiteB200(ck["state_dict"], ck["meta"], device="cuda:0")
    assert eng.supports_u8(S, S)
    rs = np.random.RandomState(S + nc)
    img = rs.randint(0, 256, (3, S, S, 3)).astype(np.uint8)
    img[0, :2] = 0; img[1, :, :3] = 255; img[2, -2:, -2:] = 0          # exercise the padded borders with extreme values
    d = torch.from_numpy(img).cuda()
    x, _ = y.preprocess_batch(d, S)
    want = eng(x)
    got = eng.forward_u8(d)
    xo = torch.from_numpy(np.concatenate([pre_ref.preprocess_ref(im, S)[0] for im in img]))
    ref = model_ref.forward_ref(ck["state_dict"], ck["meta"], xo)
    for g, w, r in zip(got, want, ref):
        assert g.shape == w.shape
        assert float((g - w).abs().max()) <= 2e-4          # same engine, two input paths
        assert float((g.cpu() - r).abs().max()) <= 1e-3    # the parity gate against the reference forward

LOGIT_TOL = 1e-3


def _engine(ckpt, **kw):
    import yololite_b200 as y
    return y.YoloLiteB200(ckpt["state_dict"], ckpt["meta"], device="cuda:0", **kw)


@pytest.mark.parametrize("name", FWD_CASES)
@pytest.mark.parametrize("fuse,tc", [(True, True), (False, True), (True, False)])
def test_forward_matches_golden_and_oracle(name, fuse, tc):
    ckpt, k = case_ckpt(name)
    g = golden(name + ".npz")
    x = model_ref.synth_input(k["B"], k["img"], seed=k["input_seed"])
    eng = _engine(ckpt, fuse_dwpw=fuse, tensor_cores=tc)
    outs = eng(x.cuda())
    torch.cuda.synchronize()
    want = model_ref.forward_ref(ckpt["state_dict"], ckpt["meta"], x)
    assert [list(o.shape) for o in outs] == g["shapes"].tolist()
    assert eng.get_strides() == g["strides"].tolist()
    step = int(g["step"])
    for i, (o, w) in enumerate(zip(outs, want)):
        assert o.is_contiguous() and o.dtype == torch.float32
        err = float((o.cpu() - w).abs().max())
        assert err <= LOGIT_TOL, (name, i, err)
        f = o.cpu().reshape(k["B"], -1, o.shape[-1]).numpy()
        np.testing.assert_allclose(f[:, ::step], g[f"level{i}"], rtol=0, atol=LOGIT_TOL)


def test_intermediate_taps_match_oracle():
    ckpt, k = case_ckpt("fwd_edge_n_64_nc3")
    x = model_ref.synth_input(2, 64, seed=5)
    eng = _engine(ckpt, fuse_dwpw=False, reuse_buffers=False)
    eng(x.cuda())
    _, feats = model_ref.forward_ref(ckpt["state_dict"], ckpt["meta"], x, return_feats=True)
    for name in ("c3", "c4", "c5", "p5", "p4", "p3"):
        got = eng.read_buffer(name, 2).permute(0, 3, 1, 2).cpu()
        assert got.shape == feats[name].shape
        assert float((got - feats[name]).abs().max()) <= 2e-4, name


def test_edge_n_640_batch_and_batch_invariance():
    ck = synth_ckpt("edge_n", 80, 640)
    eng = _engine(ck)
    x = model_ref.synth_input(3, 640, seed=1)
    outs = eng(x.cuda())
    assert [tuple(o.shape) for o in outs] == [(3, 1, 80, 80, 85), (3, 1, 40, 40, 85), (3, 1, 20, 20, 85)]
    want = model_ref.forward_ref(ck["state_dict"], ck["meta"], x[1:2])
    single = eng(x[1:2].cuda())
    for o, s, w in zip(outs, single, want):
        assert torch.equal(o[1:2], s)                          # an image's logits do not depend on its batch
        assert float((s.cpu() - w).abs().max()) <= LOGIT_TOL
    again = eng(x.cuda())
    for a, b in zip(outs, again):
        assert torch.equal(a, b)                               # deterministic


def test_non_multiple_of_32_input_and_export_concat():
    ck = synth_ckpt("edge_n", 3, 64)
    eng = _engine(ck)
    x = model_ref.synth_input(1, 112, seed=2)[:, :, :104, :88].contiguous()      # 104x88 -> 13x11, 7x6, 4x3
    want = model_ref.forward_ref(ck["state_dict"], ck["meta"], x)
    outs = eng(x.cuda())
    for o, w in zip(outs, want):
        assert o.shape == w.shape
        assert float((o.cpu() - w).abs().max()) <= LOGIT_TOL
    eng.export_concat = True
    cat = eng(x.cuda())
    assert cat.shape == (1, sum(w.shape[2] * w.shape[3] for w in want), 8)


def test_input_validation_mirrors_reference_errors():
    ck = synth_ckpt("edge_n", 3, 64)
    eng = _engine(ck)
    with pytest.raises(ValueError):
        eng(torch.zeros(1, 3, 64, 64))                 # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        eng(torch.zeros(1, 4, 64, 64, device="cuda"))
    x = torch.randn(1, 3, 64, 64, device="cuda")
    x0 = x.clone()
    eng(x)
    assert torch.equal(x, x0)                          # input not modified
