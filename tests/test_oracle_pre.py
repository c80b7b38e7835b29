"""Oracle preprocessing restatement vs cv2 itself, and the whole CLI path (config 1) vs the reference's JSON."""
import numpy as np
import pytest
import torch

from conftest import golden, synth_ckpt
from oracle import model_ref, post_ref, pre_ref

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("shape", [(360, 500, 230, 320), (100, 37, 320, 118), (33, 77, 200, 467), (500, 360, 640, 461),
                                   (720, 1280, 360, 640), (7, 5, 64, 46), (64, 64, 64, 64)])
def test_resize_bit_exact_vs_cv2(shape):
    h, w, nh, nw = shape
    img = np.random.RandomState(h + w).randint(0, 256, (h, w, 3)).astype(np.uint8)
    np.testing.assert_array_equal(pre_ref.resize_linear_u8(img, nw, nh), cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR))


def _cli_image(g):
    return cv2.resize(g["small"], (500, 360), interpolation=cv2.INTER_CUBIC)


def test_cli_path_matches_reference_json():
    """tools/infer.py main() on one synthetic image, edge_n 320 px, CPU (BASELINE config 1)."""
    g = golden("cli_edge_n_320.npz")
    ck = synth_ckpt("edge_n", 80, 320, seed=int(g["seed"]), obj_shift=float(g["obj_bias_shift"]))
    img = _cli_image(g)
    x, scale, left, top = pre_ref.preprocess_ref(img, 320)
    levels = [o.numpy() for o in model_ref.forward_ref(ck["state_dict"], ck["meta"], torch.from_numpy(x))]
    conf, iou = g["conf_iou"]
    d = post_ref.detect_ref(levels, 320, float(conf), float(iou), 300)[0]
    boxes = post_ref.backmap_ref(d["boxes"], scale, left, top, img.shape[0], img.shape[1])
    assert len(d["scores"]) == len(g["scores"]) > 0
    np.testing.assert_array_equal(d["classes"], g["classes"])
    np.testing.assert_allclose(d["scores"], g["scores"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(boxes, g["boxes"], rtol=0, atol=2e-3)
