"""GPU preprocessing kernel (letterbox + BGR->RGB + normalise + CHW) vs the oracle (bit-exact), and the predict() facade
(checkpoint file -> detections in original-image coordinates) vs the reference CLI golden (BASELINE config 1)."""
import numpy as np
import pytest
import torch

from conftest import golden, synth_ckpt
from oracle import pre_ref

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("shape,S", [((360, 500), 320), ((100, 37), 320), ((33, 77), 96), ((500, 360), 640), ((640, 640), 640),
                                     ((720, 1280), 640), ((400, 640), 640), ((640, 302), 640), ((96, 96), 96)])
def test_preprocess_bit_exact(shape, S):
    import yololite_b200 as y
    img = np.random.RandomState(shape[0]).randint(0, 256, (shape[0], shape[1], 3)).astype(np.uint8)
    want, scale, left, top = pre_ref.preprocess_ref(img, S)
    x, geo = y.preprocess([img], S, "cuda:0")
    assert geo[0][1:3] == (left, top) and abs(geo[0][0] - scale) < 1e-12
    np.testing.assert_array_equal(x.cpu().numpy(), want)
    xb, g2 = y.preprocess_batch(torch.from_numpy(np.stack([img, img[::-1].copy()])).cuda(), S)
    np.testing.assert_array_equal(xb[0].cpu().numpy(), want[0])
    np.testing.assert_array_equal(xb[1].cpu().numpy(), pre_ref.preprocess_ref(img[::-1].copy(), S)[0][0])


def test_predict_matches_reference_cli(tmp_path):
    import yololite_b200 as y
    g = golden("cli_edge_n_320.npz")
    ck = synth_ckpt("edge_n", 80, 320, seed=int(g["seed"]), obj_shift=float(g["obj_bias_shift"]))
    path = str(tmp_path / "edge_n_320.pt")
    torch.save(ck, path)
    img = cv2.resize(g["small"], (500, 360), interpolation=cv2.INTER_CUBIC)
    ipath = str(tmp_path / "synth.png")
    cv2.imwrite(ipath, img)
    model = y.YoloLite(path)
    conf, iou = (float(v) for v in g["conf_iou"])
    res = model.predict(ipath, conf=conf, iou=iou)[0]
    assert set(res) >= {"boxes", "scores", "classes", "masks", "speed"} and res["masks"] is None
    assert res["speed"]["total_ms"] > 0
    assert len(res["scores"]) == len(g["scores"]) > 0
    np.testing.assert_array_equal(res["classes"], g["classes"])
    np.testing.assert_allclose(res["scores"], g["scores"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(res["boxes"], g["boxes"], rtol=0, atol=5e-2)
    js = model.to_json(res)
    assert js["detections"][0].keys() == {"bbox_xyxy", "score", "class_id", "class_name"}
    # batched device-side path gives the same detections (letterboxed coordinates back-mapped)
    det, geo = model.predict_batch(torch.from_numpy(img[None]).cuda(), conf=conf, iou=iou)
    d0 = det.to_list()[0]
    bm = y.backmap(d0["boxes"], geo[0], geo[1], geo[2], geo[3], geo[4]).cpu().numpy()
    np.testing.assert_allclose(bm, res["boxes"], rtol=0, atol=1e-4)


def test_loader_errors_mirror_reference(tmp_path):
    import yololite_b200 as y
    p = str(tmp_path / "bad.pt")
    torch.save({"weights": 1}, p)
    with pytest.raises(RuntimeError):
        y.load_model_names_imgsize_from_ckpt(p, torch.device("cuda:0"))
    ck = synth_ckpt("edge_n", 3, 64)
    bad = {"state_dict": ck["state_dict"], "meta": dict(ck["meta"], arch="nope")}
    torch.save(bad, p)
    with pytest.raises(ValueError):
        y.load_model_names_imgsize_from_ckpt(p, torch.device("cuda:0"))
    model, names, img_size = (lambda q: (torch.save(ck, q), y.load_model_names_imgsize_from_ckpt(q, torch.device("cuda:0")))[1])(p)
    assert names == ck["meta"]["names"] and img_size == 64 and model.get_strides() == [8, 16, 32]
