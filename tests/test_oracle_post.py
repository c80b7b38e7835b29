"""Oracle postprocess restatement vs (a) golden vectors from the reference's own functions, (b) torchvision."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
from oracle import post_ref

POST_CASES = ["post_c3", "post_c1", "post_c7_a2"]


def _levels(g):
    return [g[k] for k in sorted(k for k in g.files if k.startswith("level"))]


@pytest.mark.parametrize("name", POST_CASES)
def test_decode_matches_reference(name):
    g = golden(name + ".npz")
    dec = post_ref.decode_ref(_levels(g), int(g["img"]))
    np.testing.assert_allclose(dec["box"], g["box"], rtol=0, atol=1e-4)      # px
    np.testing.assert_array_equal(dec["obj"], g["obj"])
    np.testing.assert_array_equal(dec["cls"], g["cls"])


@pytest.mark.parametrize("name", POST_CASES)
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_detections_match_reference_loop(name, tag):
    g = golden(name + ".npz")
    conf, iou = g[f"det_{tag}_conf_iou"]
    dets = post_ref.detect_ref(_levels(g), int(g["img"]), float(conf), float(iou), 300)
    for b, d in enumerate(dets):
        np.testing.assert_array_equal(d["index"], g[f"det_{tag}_{b}_index"])
        np.testing.assert_array_equal(d["classes"], g[f"det_{tag}_{b}_classes"])
        np.testing.assert_allclose(d["scores"], g[f"det_{tag}_{b}_scores"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(d["boxes"], g[f"det_{tag}_{b}_boxes"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("name", POST_CASES)
def test_coco_dets_match_reference(name):
    with open(os.path.join(GOLDEN, "post_coco.json")) as f:
        want = json.load(f)[name]
    g = golden(name + ".npz")
    got = post_ref.coco_dets_ref(_levels(g), int(g["img"]))
    assert [len(x) for x in got] == [len(x) for x in want]
    for gi, wi in zip(got, want):
        for a, b in zip(gi, wi):
            assert a["category_id"] == b["category_id"]
            assert abs(a["score"] - b["score"]) < 1e-6
            np.testing.assert_allclose(a["bbox"], b["bbox"], rtol=0, atol=1e-4)


def test_nms_ref_matches_torchvision():
    from torchvision.ops import nms
    rng = np.random.RandomState(0)
    for trial in range(60):
        n = int(rng.randint(1, 200))
        xy = rng.rand(n, 2).astype(np.float32) * 50
        wh = rng.rand(n, 2).astype(np.float32) * 30
        boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        scores = rng.rand(n).astype(np.float32)
        if trial % 3 == 0:                       # exact score ties and duplicate boxes
            scores[: n // 2] = scores[0]
            boxes[1::4] = boxes[0]
        if trial % 5 == 0:                       # zero-area boxes -> NaN IoU keeps
            boxes[::7, 2:] = boxes[::7, :2]
        for thr in (0.3, 0.5, 0.65):
            want = nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
            got = post_ref.nms_ref(boxes, scores, thr)
            np.testing.assert_array_equal(got, want)


def test_decode_closed_form_at_zero_logits():
    # logits 0 -> px = (0.5+gx)*stride, pw = ln2*stride
    S, img = 4, 64
    lv = [np.zeros((1, 1, S, S, 6), np.float32)]
    box = post_ref.decode_ref(lv, img)["box"][0].reshape(S, S, 4)
    stride = img / S
    for gy in range(S):
        for gx in range(S):
            cx, cy, w = (0.5 + gx) * stride, (0.5 + gy) * stride, np.log(2.0) * stride
            np.testing.assert_allclose(box[gy, gx], [cx - w / 2, cy - w / 2, cx + w / 2, cy + w / 2], atol=1e-4)


def test_output_order_and_single_class_rule():
    rng = np.random.RandomState(3)
    lv = [rng.randn(1, 1, 6, 6, 5 + 4).astype(np.float32) * 2 + 1]
    d = post_ref.detect_ref(lv, 48, 0.05, 0.5)[0]
    assert np.all(np.diff(d["classes"]) >= 0)                       # classes ascending
    for c in np.unique(d["classes"]):
        s = d["scores"][d["classes"] == c]
        assert np.all(np.diff(s) <= 0)                              # score-descending within a class
    lv1 = [rng.randn(1, 1, 6, 6, 6).astype(np.float32)]
    d1 = post_ref.detect_ref(lv1, 48, 0.0, 1.0)[0]                  # C == 1 -> score = sigmoid(obj) only
    obj = 1.0 / (1.0 + np.exp(-lv1[0][0, 0, :, :, 4].reshape(-1)))
    np.testing.assert_allclose(np.sort(d1["scores"]), np.sort(obj.astype(np.float32)), atol=1e-6)
