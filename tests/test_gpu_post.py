"""Parity of the fused postprocess kernel (through the C ABI) against the oracle and reference golden vectors.

Bar: identical kept anchor indices / classes / order; boxes <= 1e-3 px; scores <= 1e-5."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import post_ref

pytestmark = pytest.mark.gpu
SETTINGS = [(0.25, 0.5, 300), (0.4, 0.5, 300), (0.001, 0.65, 0)]


def _run(levels_np, img, conf, iou, max_det, cap=None):
    import yololite_b200 as y
    lv = [torch.from_numpy(np.ascontiguousarray(l)).cuda() for l in levels_np]
    return y.detect(lv, img, conf, iou, max_det, cap)


def _check(got, want, tag=""):
    assert len(got) == len(want)
    for b, (g, w) in enumerate(zip(got, want)):
        gi = g["index"].cpu().numpy()
        assert gi.shape == w["index"].shape, (tag, b, gi.shape, w["index"].shape)
        if not np.array_equal(gi, w["index"]):
            # the only tolerated difference: two candidates of one class whose scores differ by <= 1 ulp between
            # CUDA expf and numpy exp swap places (the reference's own CPU and CUDA paths differ the same way)
            np.testing.assert_array_equal(np.sort(gi), np.sort(w["index"]), err_msg=f"{tag} image {b}")
            bad = np.nonzero(gi != w["index"])[0]
            assert len(bad) <= 4 and np.all(np.abs(w["scores"][bad] - g["scores"].cpu().numpy()[bad]) <= 2e-7), (tag, b, bad)
        np.testing.assert_array_equal(g["classes"].cpu().numpy(), w["classes"])
        np.testing.assert_allclose(g["scores"].cpu().numpy(), w["scores"], rtol=0, atol=1e-5)
        if np.array_equal(gi, w["index"]):
            np.testing.assert_allclose(g["boxes"].cpu().numpy(), w["boxes"], rtol=0, atol=1e-3)
        assert g["classes"].dtype == torch.int64 and g["boxes"].dtype == torch.float32


def _glevels(g):
    return [g[k] for k in sorted(k for k in g.files if k.startswith("level"))]


@pytest.mark.parametrize("name", ["post_c3", "post_c1", "post_c7_a2"])
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_matches_reference_golden(name, tag):
    g = golden(name + ".npz")
    conf, iou = (float(v) for v in g[f"det_{tag}_conf_iou"])
    got = _run(_glevels(g), int(g["img"]), conf, iou, 300)
    want = [{k: g[f"det_{tag}_{b}_{k}"] for k in ("boxes", "scores", "classes", "index")} for b in range(2)]
    _check(got, want, f"{name}/{tag}")


def _random_levels(rng, B, C, sizes, A=1, obj_shift=-3.0, wh_shift=2.0):
    out = []
    for S in sizes:
        t = (rng.randn(B, A, S, S, 5 + C) * 2.0).astype(np.float32)
        t[..., 2:4] += wh_shift
        t[..., 4] += obj_shift
        out.append(t)
    return out


@pytest.mark.parametrize("conf,iou,max_det", SETTINGS)
def test_edge_n_640_shape_random_logits(conf, iou, max_det):
    rng = np.random.RandomState(17)
    lv = _random_levels(rng, 3, 80, (80, 40, 20), obj_shift=-2.0 if conf > 0.01 else -6.0)
    want = post_ref.detect_ref(lv, 640, conf, iou, max_det)
    assert sum(len(w["index"]) for w in want) > 50
    _check(_run(lv, 640, conf, iou, max_det), want, f"{conf}/{iou}")


def test_single_class_heavy_overlap_long_segments():
    # C == 1: score = sigmoid(obj) only; every candidate is in one NMS segment (> 128 -> memory path)
    rng = np.random.RandomState(5)
    lv = _random_levels(rng, 2, 1, (40, 20, 10), obj_shift=0.0, wh_shift=4.0)
    for conf, iou, md in ((0.25, 0.5, 300), (0.001, 0.65, 0), (0.25, 0.5, 7)):
        want = post_ref.detect_ref(lv, 320, conf, iou, md)
        assert len(want[0]["index"]) > 0
        _check(_run(lv, 320, conf, iou, md), want, f"c1/{conf}/{md}")


def test_more_than_8192_candidates_uses_global_sort():
    rng = np.random.RandomState(9)
    lv = _random_levels(rng, 1, 3, (96, 48, 24), obj_shift=3.0, wh_shift=0.0)      # N = 12096, nearly all pass
    want = post_ref.detect_ref(lv, 768, 0.001, 0.65, 0)
    assert len(want[0]["index"]) > 8192
    _check(_run(lv, 768, 0.001, 0.65, 0), want, "big")


def test_max_det_per_class_and_capacity_flag():
    import yololite_b200 as y
    rng = np.random.RandomState(2)
    lv = _random_levels(rng, 2, 2, (16, 8, 4), obj_shift=2.0, wh_shift=-1.0)
    want = post_ref.detect_ref(lv, 128, 0.1, 0.9, 5)
    assert max(np.bincount(w["classes"]).max() for w in want) == 5
    _check(_run(lv, 128, 0.1, 0.9, 5), want, "maxdet")
    pp = y.PostProcessor()
    d = pp([torch.from_numpy(l).cuda() for l in lv], 128, 0.1, 0.9, 0, cap=4)
    assert all(int(c) & (1 << 30) for c in d.counts.cpu())
    with pytest.raises(RuntimeError):
        d.to_list()


def test_empty_result_ties_and_unaligned_levels():
    rng = np.random.RandomState(4)
    lv = _random_levels(rng, 2, 5, (7, 3, 1), obj_shift=-30.0)          # 49+9+1 anchors, D=10: unaligned tiles
    got = _run(lv, 56, 0.25, 0.5, 300)
    assert all(g["index"].numel() == 0 for g in got)
    lv = _random_levels(rng, 2, 5, (7, 3, 1), obj_shift=1.0, wh_shift=3.0)
    lv[0][0, 0, 2, 3] = lv[0][0, 0, 2, 2]                               # identical logits -> score tie, idx order
    lv[0][1, 0, 1, 1, 2:4] = -60.0                                      # zero-area boxes
    lv[0][1, 0, 1, 2, 2:4] = -60.0
    for conf, iou, md in SETTINGS:
        _check(_run(lv, 56, conf, iou, md), post_ref.detect_ref(lv, 56, conf, iou, md), "small")


def test_first_max_class_on_saturated_sigmoid_ties():
    # two class logits that both saturate sigmoid to 1.0f: torch.max over sigmoid values returns the FIRST
    lv = [np.full((1, 1, 2, 2, 5 + 4), -2.0, np.float32)]
    lv[0][..., 4] = 3.0
    lv[0][0, 0, 0, 0, 5:] = [1.0, 20.0, 30.0, 25.0]     # sigmoid(20)=sigmoid(30)=1.0f -> class 1
    lv[0][0, 0, 0, 1, 5:] = [18.0, 1.0, 40.0, 17.5]     # sigmoid(18) == 1.0f -> class 0
    want = post_ref.detect_ref(lv, 16, 0.25, 0.5)
    assert sorted(want[0]["classes"].tolist())[:2] == [0, 1]
    _check(_run(lv, 16, 0.25, 0.5, 300), want, "sat")


def test_decode_kernel_matches_oracle():
    import yololite_b200 as y
    rng = np.random.RandomState(8)
    lv = _random_levels(rng, 2, 6, (12, 6, 3), A=2)
    lv[0][0, 0, 0, 0, 2] = 25.0          # softplus threshold branch
    lv[0][0, 0, 0, 1, 0] = -40.0
    want = post_ref.decode_ref(lv, 96)
    got = y.decode_preds_anchorfree([torch.from_numpy(l).cuda() for l in lv], img_size=96, center_mode="v8", wh_mode="softplus")
    assert got["box"].shape == want["box"].shape and got["obj"].shape == want["obj"].shape and got["cls"].shape == want["cls"].shape
    np.testing.assert_allclose(got["box"].cpu().numpy(), want["box"], rtol=0, atol=1e-3)
    np.testing.assert_array_equal(got["obj"].cpu().numpy(), want["obj"])
    np.testing.assert_array_equal(got["cls"].cpu().numpy(), want["cls"])
    with pytest.raises(AssertionError):
        y.decode_preds_anchorfree([torch.zeros(2, 1, 4, 4, 8).cuda(), torch.zeros(1, 1, 2, 2, 8).cuda()], 32)


def test_full_batch_properties_at_bench_size():
    """B=64, N=8400, C=80 (the bench workload): size-independent properties instead of a CPU comparison."""
    import yololite_b200 as y
    g = torch.Generator(device="cuda").manual_seed(0)
    lv = [torch.randn(64, 1, s, s, 85, device="cuda", generator=g) * 2 for s in (80, 40, 20)]
    for l in lv:
        l[..., 4] -= 2.0
        l[..., 2:4] += 2.0
    pp = y.PostProcessor()
    d1 = pp(lv, 640, 0.25, 0.5, 300)
    r1 = [{k: v.clone() for k, v in x.items()} for x in d1.to_list()]
    r2 = pp(lv, 640, 0.25, 0.5, 300).to_list()
    for a, b in zip(r1, r2):                                   # idempotent / deterministic
        assert torch.equal(a["index"], b["index"]) and torch.equal(a["boxes"], b["boxes"])
    sub = pp([l[5:6].contiguous() for l in lv], 640, 0.25, 0.5, 300).to_list()[0]
    assert torch.equal(sub["index"], r1[5]["index"])           # per-image independence
    for r in r1[:8]:
        c, s = r["classes"].cpu().numpy(), r["scores"].cpu().numpy()
        assert np.all(np.diff(c) >= 0) and np.all(s > 0.25)
        for cc in np.unique(c):
            assert np.all(np.diff(s[c == cc]) <= 0)
        b = r["boxes"].cpu().numpy()
        assert b.min() >= 0 and b.max() <= 639
    want = post_ref.detect_ref([l[:2].cpu().numpy() for l in lv], 640, 0.25, 0.5, 300)
    _check(r1[:2], want, "bench-size")


@pytest.mark.parametrize("name", ["post_c3", "post_c1", "post_c7_a2"])
def test_coco_dets_wrapper_matches_reference_golden(name):
    """The PRODUCT's evaluation wrapper post.decode_batch_to_coco_dets (conf 0.001, iou 0.65, no cap, (cx,cy,w,h) boxes,
    category_id = class + 1) against the dets the unmodified scripts/helpers/helpers.py:86-153 produced for the same
    logits (tests/golden/post_coco.json, written by oracle/make_golden.py)."""
    import json
    import os
    import yololite_b200 as y
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "post_coco.json")) as f:
        want = json.load(f)[name]
    g = golden(name + ".npz")
    lv = [torch.from_numpy(np.ascontiguousarray(l)).cuda() for l in _glevels(g)]
    got = y.decode_batch_to_coco_dets(lv, int(g["img"]))
    assert [len(x) for x in got] == [len(x) for x in want]
    for b, (gi, wi) in enumerate(zip(got, want)):
        assert [d["category_id"] for d in gi] == [d["category_id"] for d in wi], (name, b)
        assert all(set(d) == {"category_id", "bbox", "score"} and isinstance(d["category_id"], int) for d in gi)
        np.testing.assert_allclose([d["score"] for d in gi], [d["score"] for d in wi], rtol=0, atol=1e-5)
        np.testing.assert_allclose(np.array([d["bbox"] for d in gi]).reshape(-1, 4), np.array([d["bbox"] for d in wi]).reshape(-1, 4),
                                   rtol=0, atol=1e-3)
    # add_one=False keeps the raw class index (helpers.py:86 signature)
    raw = y.decode_batch_to_coco_dets(lv, int(g["img"]), add_one=False)
    assert [d["category_id"] + 1 for d in raw[0]] == [d["category_id"] for d in want[0]]
