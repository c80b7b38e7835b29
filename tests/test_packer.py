"""Host-side lowering (BN folding, weight layouts, op wiring, buffer reuse) checked on CPU: the op program,
executed by a plain-torch interpreter of the documented op semantics, must reproduce the oracle forward."""
import numpy as np
import pytest
import torch

from conftest import FWD_CASES, case_ckpt
from oracle import model_ref
from program_interp import run_program
from yololite_b200 import packer


@pytest.mark.parametrize("name", FWD_CASES)
@pytest.mark.parametrize("fuse,reuse", [(True, True), (False, False)])
def test_program_reproduces_oracle(name, fuse, reuse):
    ckpt, k = case_ckpt(name)
    P = packer.lower(ckpt["state_dict"], ckpt["meta"], fuse_dwpw=fuse, reuse_buffers=reuse)
    x = model_ref.synth_input(k["B"], k["img"], seed=k["input_seed"])
    want = model_ref.forward_ref(ckpt["state_dict"], ckpt["meta"], x)
    got = run_program(P, x)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) < 2e-4
    assert P.strides == model_ref.strides_ref(ckpt["meta"])


def test_every_checkpoint_tensor_is_consumed_or_known_unused():
    ckpt, _ = case_ckpt("fwd_edge_n_64_nc3")
    sd = packer._SD(ckpt["state_dict"])
    orig = packer._SD
    try:
        packer._SD = lambda _: sd
        packer.lower(ckpt["state_dict"], ckpt["meta"])
    finally:
        packer._SD = orig
    unused = set(ckpt["state_dict"]) - sd.used
    # model_v2.py:297-300: the P6 branch parameters exist in every state_dict even when use_p6 is False
    assert all(k.startswith(("p6_down", "p6_bn", "smooth6")) for k in unused), sorted(unused)[:5]


def test_buffer_reuse_never_aliases_live_tensors():
    ckpt, _ = case_ckpt("fwd_edge_n_96_p2p6_a2")
    P = packer.lower(ckpt["state_dict"], ckpt["meta"], reuse_buffers=True)
    Pn = packer.lower(ckpt["state_dict"], ckpt["meta"], reuse_buffers=False)
    assert P.n_buffers < Pn.n_buffers
    for op in P.ops:
        ins = {op[f] for f in ("src", "res", "up") if op[f] >= 0}
        assert op["dst"] not in ins
    # blob arrays are 16-byte aligned, as the kernels' float4 loads require
    for op in P.ops:
        assert op["w_off"] % 4 == 0 and (op["b_off"] < 0 or op["b_off"] % 4 == 0) and (op["w2_off"] < 0 or op["w2_off"] % 4 == 0)


def test_parse_meta_mirrors_reference_errors():
    meta = model_ref.make_meta("edge_n", 3, 64)
    cfg = packer.parse_meta(meta)
    assert (cfg.fpn_channels, cfg.depth, cfg.head_depth, cfg.anchors) == (96, 1, 1, (1, 1, 1))
    m = packer.parse_meta(model_ref.make_meta("edge_m", 3, 64))
    assert (m.fpn_channels, m.depth, m.head_depth) == (244, 2, 2)
    bad = dict(meta); bad["arch"] = "resnet"
    with pytest.raises(ValueError):
        packer.parse_meta(bad)
    bad = dict(meta); bad["config"] = {"model": meta["config"]["model"], "training": {}}
    with pytest.raises(KeyError):
        packer.parse_meta(bad)      # tools/infer.py:49-50 hard-indexes use_p6/use_p2
    bad = dict(meta); bad["backbone"] = "hgnetv2_b0"
    with pytest.raises(ValueError):
        packer.parse_meta(bad)


def test_product_synth_checkpoint_matches_reference_state_dict_spec():
    from yololite_b200 import synth
    for mdl, kw in (("edge_n", {}), ("edge_m", {}), ("edge_n", dict(use_p2=True, use_p6=True, anchors=2))):
        meta = synth.make_meta(mdl, 5, 64, **kw)
        mine = synth.state_shapes(meta)
        spec = model_ref.state_spec(model_ref.make_meta(mdl, 5, 64, **kw))
        assert set(mine) == set(spec)
        for k, shp in mine.items():
            assert tuple(shp) == tuple(spec[k][0]), k
    ck = synth.random_checkpoint(synth.make_meta("edge_n", 4, 64), seed=0, obj_bias=-1.0)
    x = model_ref.synth_input(1, 64, seed=0)
    outs = run_program(packer.lower(ck["state_dict"], ck["meta"]), x)
    want = model_ref.forward_ref(ck["state_dict"], ck["meta"], x)
    for a, b in zip(outs, want):
        assert torch.isfinite(a).all() and float((a - b).abs().max()) < 2e-4


def test_tc_weight_image_layout():
    """[nslab][3 splits][Npad][32 k] bf16 with the SWIZZLE_64B chunk permutation, checked element by element; the three splits
    reproduce the fp32 weight to 2^-24."""
    rng = np.random.RandomState(0)

    def bf(u16):
        return (np.uint32(u16) << np.uint32(16)).view(np.float32) if isinstance(u16, np.ndarray) else np.array([u16], np.uint32).__lshift__(16).view(np.float32)[0]
    for K, N in ((27, 32), (48, 85), (96, 96), (288, 16)):
        wm = rng.randn(K, (N + 3) // 4 * 4)
        img = packer.tc_image(wm, N).view(np.uint16)
        nslab, npad = (K + 31) // 32, (N + 15) // 16 * 16
        assert img.size == nslab * 3 * npad * 32
        img = img.reshape(nslab, 3, npad, 32)
        for _ in range(300):
            k, n = int(rng.randint(0, nslab * 32)), int(rng.randint(0, npad))
            s, kk = divmod(k, 32)
            pos = (((kk // 8) ^ ((n >> 1) & 3)) * 8) + kk % 8
            want = wm[k, n] if (k < K and n < N) else 0.0
            parts = [float(bf(img[s, q, n, pos])) for q in range(3)]
            assert abs(want - sum(parts)) <= abs(want) * 2.0 ** -23 + 1e-30
            if want:
                assert abs(parts[0] - want) <= abs(want) * 2.0 ** -8 and abs(parts[1]) <= abs(want) * 2.0 ** -8


def test_stem_u8_matrix_reproduces_normalise_then_conv():
    """The uint8 entry folds (u/255 - mean)/std into the stem weights (packer.stem_u8_matrix): checked in float64 against
    normalise -> zero-padded 3x3 s2 conv on an even-sized image, including the top / left border pixels whose padded taps
    the indicator columns take back."""
    rs = np.random.RandomState(0)
    H, W, n = 12, 16, 32
    w = rs.randn(n, 3, 3, 3)                                   # [n][ci][ky][kx]
    b0 = rs.randn(n)
    ws = np.transpose(w, (2, 3, 1, 0)).reshape(27, n)          # k = (ky*3+kx)*3 + ci
    su = packer.stem_u8_matrix(ws, b0)
    img = rs.randint(0, 256, (H, W, 3)).astype(np.float64)     # RGB order, integer values
    mean, std = np.array(packer.IMAGENET_MEAN), np.array(packer.IMAGENET_STD)
    x = (img / 255.0 - mean) / std
    want = torch.nn.functional.conv2d(torch.from_numpy(x).permute(2, 0, 1)[None], torch.from_numpy(w), torch.from_numpy(b0),
                                      stride=2, padding=1)[0].numpy()           # [n][H/2][W/2]
    pad = np.zeros((H + 2, W + 2, 3))
    pad[1:-1, 1:-1] = img                                       # raw bytes, zero outside (what the kernel's patch holds)
    got = np.zeros_like(want)
    for sy in range(H // 2):
        for sx in range(W // 2):
            a = np.zeros(32)
            a[:27] = pad[2 * sy:2 * sy + 3, 2 * sx:2 * sx + 3, :].reshape(27)
            a[27], a[28], a[29] = 1.0, float(sy == 0), float(sx == 0)
            a[30] = a[28] * a[29]
            got[:, sy, sx] = su @ a
    assert np.abs(got - want).max() < 1e-10


def test_edge_n_program_shape():
    """The fusions the kernels rely on: edge_n lowers to 40 ops -- one fused stem op (conv_stem + blocks.0.0 + blocks.0.1), every
    depthwise conv of the UIR blocks riding in a YL_OP_DWPW (stride 2 included), no standalone depthwise op left."""
    from yololite_b200 import _lib as L
    ck, _ = case_ckpt("fwd_edge_n_320_nc80")
    P = packer.lower(ck["state_dict"], ck["meta"])
    kinds = [op["kind"] for op in P.ops]
    assert len(P.ops) == 40 and kinds[0] == L.OP_STEM2 and P.ops[0]["b2_off"] >= 0 and P.ops[0]["w3_off"] >= 0
    assert L.OP_DW not in kinds and L.OP_STEM not in kinds
    dwpw = [op for op in P.ops if op["kind"] == L.OP_DWPW]
    assert len(dwpw) == 21 and sum(op["stride2"] == 2 for op in dwpw) == 2 and {op["k2"] for op in dwpw} == {3, 5}
    assert sum(op["res"] >= 0 for op in P.ops) == 10          # blocks 2.1-2.5 and 3.1-3.5 carry the skip connection
    # unfused lowering keeps the reference's layer list: 45 backbone convs + 9 FPN + 12 head convs, head outputs merged per level
    Pn = packer.lower(ck["state_dict"], ck["meta"], fuse_dwpw=False, fuse_stem=False)
    assert len(Pn.ops) == 45 + 9 + 6 + 3
