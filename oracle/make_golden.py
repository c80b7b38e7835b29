"""ORACLE tooling: generate tests/golden/*.npz|json by running the UNMODIFIED reference from /root/reference.

Run in the authoring container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

What is taken from the reference, unmodified:
  scripts/model/model_v2.py   YOLOLiteMS_CPU / YOLOLiteMS module trees (FPN + heads), through
  tools/infer.py              load_model_names_imgsize_from_ckpt(), nms(), and main() (CLI, config 1)
  scripts/helpers/utils_ms.py decode_preds_anchorfree()
  scripts/helpers/helpers.py  _decode_batch_to_coco_dets()
What is NOT the reference: `timm` (absent) is replaced by oracle/timm_shim (backbone restated from the
notebook dump) -- so backbone numerics are oracle-vs-oracle; everything downstream is reference-vs-oracle.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, os.path.join(HERE, "timm_shim"))
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from oracle import model_ref  # noqa: E402


def _ref_model(ckpt, tmp):
    import tools.infer as ref_infer
    path = os.path.join(tmp, "ckpt.pt")
    torch.save(ckpt, path)
    model, names, img_size = ref_infer.load_model_names_imgsize_from_ckpt(path, torch.device("cpu"))
    return model, names, img_size, path


def gen_forward_cases(tmp):
    cases = [
        # name, model, nc, img, B, p2, p6, anchors, stride of stored anchors
        ("fwd_edge_n_64_nc3", "edge_n", 3, 64, 2, False, False, 1, 1),
        ("fwd_edge_n_320_nc80", "edge_n", 80, 320, 1, False, False, 1, 8),
        ("fwd_edge_n_96_p2p6_a2", "edge_n", 5, 96, 2, True, True, 2, 1),
        ("fwd_edge_m_64_nc3", "edge_m", 3, 64, 1, False, False, 1, 1),
        ("fwd_ms_n_64_nc4_p6", "ms_n_mnv4", 4, 64, 1, False, True, 1, 1),
    ]
    kat = {}
    for name, mdl, nc, img, B, p2, p6, A, step in cases:
        meta = model_ref.make_meta(mdl, nc, img, use_p2=p2, use_p6=p6, anchors=A)
        ckpt = model_ref.synth_checkpoint(meta, seed=7)
        model, names, img_size, _ = _ref_model(ckpt, tmp)
        ref_sd = model.state_dict()
        spec = model_ref.state_spec(meta)
        assert set(ref_sd.keys()) == set(spec.keys()), (set(ref_sd) ^ set(spec))
        for k, (shape, _) in spec.items():
            assert tuple(ref_sd[k].shape) == tuple(shape), (k, ref_sd[k].shape, shape)
        x = model_ref.synth_input(B, img, seed=3)
        with torch.no_grad():
            outs = model(x)
        mine = model_ref.forward_ref(ckpt["state_dict"], meta, x)
        err = max(float((a - b).abs().max()) for a, b in zip(outs, mine))
        assert err < 1e-4, (name, err)
        flat = [o.reshape(B, -1, o.shape[-1]).numpy() for o in outs]
        store = {f"level{i}": f[:, ::step].copy() for i, f in enumerate(flat)}
        store["sum"] = np.array([float(f.astype(np.float64).sum()) for f in flat])
        store["abssum"] = np.array([float(np.abs(f.astype(np.float64)).sum()) for f in flat])
        store["shapes"] = np.array([list(o.shape) for o in outs], np.int64)
        store["strides"] = np.array(model.get_strides(), np.int64)
        store["step"] = np.array(step)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **store)
        kat[name] = dict(model=mdl, nc=nc, img=img, B=B, p2=p2, p6=p6, anchors=A, seed=7, input_seed=3,
                         n_keys=len(ref_sd), n_params=int(sum(p.numel() for p in model.parameters())),
                         oracle_vs_reference_maxabs=err, logit_absmax=float(max(np.abs(f).max() for f in flat)))
        print(name, kat[name])
    # parameter-count known answers (BENCHMARK.md:353-355, nc=3, no P2/P6 heads)
    for mdl in ("edge_n", "edge_s", "edge_m"):
        meta = model_ref.make_meta(mdl, 3, 640)
        spec = model_ref.state_spec(meta)
        n = sum(int(np.prod(s)) for k, (s, kind) in spec.items() if not kind.startswith("bn_r") and kind != "bn_nbt")
        kat["params_" + mdl] = dict(n_params=n, n_keys=len(spec))
        print(mdl, kat["params_" + mdl])
    return kat


def gen_post_cases():
    """Random logits -> reference decode / NMS loop / coco dets."""
    from scripts.helpers.utils_ms import decode_preds_anchorfree
    from scripts.helpers.helpers import _decode_batch_to_coco_dets
    import tools.infer as ref_infer
    out = {}
    g = torch.Generator().manual_seed(11)
    for name, C, A, sizes, img in (("post_c3", 3, 1, (8, 4, 2), 64), ("post_c1", 1, 1, (8, 4, 2), 64),
                                   ("post_c7_a2", 7, 2, (12, 6, 3, 2), 96)):
        B = 2
        levels = []
        for S in sizes:
            t = torch.randn(B, A, S, S, 5 + C, generator=g) * 2.0
            t[..., 2:4] += (3.0 if C != 7 else 1.5)   # big boxes -> many real overlaps
            t[..., 4] += 1.0
            levels.append(t)
        # a few exact duplicates / zero-area / tie cases
        levels[0][0, 0, 0, 1] = levels[0][0, 0, 0, 0]
        levels[0][1, 0, 1, 1, 2:4] = -60.0     # softplus -> ~0 : zero-area box
        levels[0][1, 0, 1, 2, 2:4] = -60.0
        dec = decode_preds_anchorfree(levels, img_size=img, center_mode="v8", wh_mode="softplus")
        rec = {f"level{i}": l.numpy() for i, l in enumerate(levels)}
        rec.update(box=dec["box"].numpy(), obj=dec["obj"].numpy(), cls=dec["cls"].numpy(), img=np.array(img))
        for tag, conf, iou in (("a", 0.25, 0.5), ("b", 0.4, 0.5), ("c", 0.001, 0.65)):
            # the tools/infer.py:466-493 loop, driven through the reference's own nms() wrapper
            for b in range(B):
                boxes_t = dec["box"][b]; obj = dec["obj"][b].squeeze(-1).sigmoid(); cls_log = dec["cls"][b]
                if cls_log.shape[-1] > 1:
                    confs, cls_idx = cls_log.sigmoid().max(dim=-1); scores_t = obj * confs
                else:
                    cls_idx = torch.zeros_like(obj, dtype=torch.long); scores_t = obj
                m0 = scores_t > conf
                idx0 = torch.nonzero(m0).squeeze(-1)
                bt, st, ct = boxes_t[m0], scores_t[m0], cls_idx[m0]
                fb, fs, fc, fi = [], [], [], []
                for c in ct.unique():
                    mc = ct == c
                    keep = ref_infer.nms(bt[mc], st[mc], iou)
                    fb.append(bt[mc][keep]); fs.append(st[mc][keep]); fi.append(idx0[mc][keep])
                    fc.append(torch.full((keep.numel(),), int(c), dtype=torch.long))
                cat = lambda l, dt: (torch.cat(l).numpy() if l else np.zeros((0,), dt))
                rec[f"det_{tag}_{b}_boxes"] = torch.cat(fb).numpy() if fb else np.zeros((0, 4), np.float32)
                rec[f"det_{tag}_{b}_scores"] = cat(fs, np.float32)
                rec[f"det_{tag}_{b}_classes"] = cat(fc, np.int64)
                rec[f"det_{tag}_{b}_index"] = cat(fi, np.int64)
            rec[f"det_{tag}_conf_iou"] = np.array([conf, iou])
        coco = _decode_batch_to_coco_dets(levels, img, conf_th=0.001, iou_th=0.65)
        out[name] = coco
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "N =", dec["box"].shape[1], "coco dets/img =", [len(c) for c in coco],
              "kept@0.25/0.5 =", [len(rec[f"det_a_{b}_index"]) for b in range(B)])
    with open(os.path.join(OUT, "post_coco.json"), "w") as f:
        json.dump(out, f)


def gen_cli_case(tmp):
    """Config 1: edge_n 320 px, batch 1, CPU, through tools/infer.py main() on one synthetic image."""
    import cv2
    import tools.infer as ref_infer
    meta = model_ref.make_meta("edge_n", 80, 320)
    ckpt = model_ref.synth_checkpoint(meta, seed=7, obj_bias_shift=1.5)
    path = os.path.join(tmp, "edge_n_320.pt")
    torch.save(ckpt, path)
    rng = np.random.RandomState(5)
    # smooth-ish synthetic photo: low-res noise upsampled, 360x500 BGR (non-square -> exercises letterbox)
    small = rng.randint(0, 256, (23, 32, 3)).astype(np.uint8)
    img = cv2.resize(small, (500, 360), interpolation=cv2.INTER_CUBIC)
    ipath = os.path.join(tmp, "synth.png")
    cv2.imwrite(ipath, img)
    cwd = os.getcwd()
    os.chdir(tmp)
    argv = sys.argv
    try:
        sys.argv = ["infer.py", "--weights", path, "--img", ipath, "--device", "cpu", "--conf", "0.25", "--iou", "0.5"]
        ref_infer.main()
    finally:
        sys.argv = argv
        os.chdir(cwd)
    with open(os.path.join(tmp, "runs", "infer", "1", "json", "synth.json")) as f:
        dets = json.load(f)["detections"]
    np.savez_compressed(os.path.join(OUT, "cli_edge_n_320.npz"), small=small,
                        boxes=np.array([d["bbox_xyxy"] for d in dets], np.float32).reshape(-1, 4),
                        scores=np.array([d["score"] for d in dets], np.float32),
                        classes=np.array([d["class_id"] for d in dets], np.int64),
                        conf_iou=np.array([0.25, 0.5]), seed=np.array(7), obj_bias_shift=np.array(1.5))
    print("cli case:", len(dets), "detections")
    return len(dets)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    with tempfile.TemporaryDirectory() as tmp:
        kat = gen_forward_cases(tmp)
        gen_post_cases()
        kat["cli_n_dets"] = gen_cli_case(tmp)
    kat["torch"] = torch.__version__
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
