"""ORACLE (test infrastructure, not product code).

numpy fp32 restatement of the reference per-image postprocess:

  decode_ref        scripts/helpers/utils_ms.py:25-123   (center_mode="v8", wh_mode="softplus" -- the only
                    modes any reference caller passes: tools/infer.py:462, scripts/helpers/helpers.py:96)
  score_ref         tools/infer.py:466-475               (sigmoid, class max, obj*conf, C==1 -> obj only)
  nms_ref           torchvision.ops.nms CPU semantics as the reference reaches them through
                    tools/infer.py:134-152 (stable descending sort, IoU = inter/(a+b-inter) in fp32,
                    suppress iff IoU > thr compared in double, NaN keeps)
  detect_ref        tools/infer.py:476-493               (classes ascending, per-class NMS, per-class max_det)
  coco_dets_ref     scripts/helpers/helpers.py:86-153    (eval variant: conf 0.001 / iou 0.65, xywh, cat+1)

Pinned in tests/test_oracle_post.py against torchvision.ops.nms itself and against golden vectors produced
by the reference functions imported from /root/reference (oracle/make_golden.py).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

f32 = np.float32


def _sigmoid(x: np.ndarray) -> np.ndarray:
    x = x.astype(f32, copy=False)
    with np.errstate(over="ignore"):
        return (f32(1.0) / (f32(1.0) + np.exp(-x, dtype=f32))).astype(f32)


def _softplus(x: np.ndarray) -> np.ndarray:
    x = x.astype(f32, copy=False)
    with np.errstate(over="ignore"):
        sp = np.log1p(np.exp(x, dtype=f32), dtype=f32)
    return np.where(x > f32(20.0), x, sp).astype(f32)      # torch softplus threshold=20, beta=1


def decode_ref(levels: Sequence[np.ndarray], img_size: int) -> Dict[str, np.ndarray]:
    """levels: list of [B,A,S,S,5+C] (or [B,S,S,5+C]) fp32 -> {"box":[B,N,4], "obj":[B,N,1], "cls":[B,N,C]}."""
    boxes, objs, clss = [], [], []
    B = levels[0].shape[0]
    lim = f32(img_size - 1)
    for p in levels:
        p = np.asarray(p, dtype=f32)
        if p.ndim == 4:
            p = p[:, None]
        B_, A, S, S2, D = p.shape
        assert B_ == B, "batch mismatch between levels"
        stride = f32(img_size / float(S))
        gx = np.arange(S2, dtype=f32).reshape(1, 1, 1, S2)
        gy = np.arange(S, dtype=f32).reshape(1, 1, S, 1)
        px = ((_sigmoid(p[..., 0]) * f32(2.0) - f32(0.5)) + gx) * stride
        py = ((_sigmoid(p[..., 1]) * f32(2.0) - f32(0.5)) + gy) * stride
        pw = _softplus(p[..., 2]) * stride
        ph = _softplus(p[..., 3]) * stride
        x1 = np.clip(px - pw * f32(0.5), f32(0), lim)
        y1 = np.clip(py - ph * f32(0.5), f32(0), lim)
        x2 = np.clip(px + pw * f32(0.5), f32(0), lim)
        y2 = np.clip(py + ph * f32(0.5), f32(0), lim)
        n = A * S * S2
        boxes.append(np.stack([x1, y1, x2, y2], -1).reshape(B, n, 4).astype(f32))
        objs.append(p[..., 4].reshape(B, n, 1))
        clss.append(p[..., 5:].reshape(B, n, D - 5))
    return {"box": np.concatenate(boxes, 1), "obj": np.concatenate(objs, 1), "cls": np.concatenate(clss, 1)}


def score_ref(obj_logit: np.ndarray, cls_logit: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """[N], [N,C] -> (score [N] f32, class [N] i64); first-max class on ties (torch CPU behaviour)."""
    obj = _sigmoid(obj_logit)
    if cls_logit.shape[-1] > 1:
        p = _sigmoid(cls_logit)
        ci = p.argmax(-1)                       # numpy argmax returns the first maximum
        return (obj * p[np.arange(p.shape[0]), ci]).astype(f32), ci.astype(np.int64)
    return obj, np.zeros(obj.shape, np.int64)


def nms_ref(boxes: np.ndarray, scores: np.ndarray, iou_thr: float) -> np.ndarray:
    """Indices kept, in descending-score order (ties: lower index first)."""
    boxes = np.asarray(boxes, f32).reshape(-1, 4)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    order = np.argsort(-np.asarray(scores, f32), kind="stable")
    x1, y1, x2, y2 = (boxes[:, i] for i in range(4))
    area = ((x2 - x1) * (y2 - y1)).astype(f32)
    dead = np.zeros(n, bool)
    keep = []
    thr = float(iou_thr)
    for pos in range(n):
        i = order[pos]
        if dead[i]:
            continue
        keep.append(i)
        rest = order[pos + 1:]
        if rest.size == 0:
            break
        w = np.maximum(f32(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest])).astype(f32)
        h = np.maximum(f32(0), np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest])).astype(f32)
        inter = (w * h).astype(f32)
        with np.errstate(invalid="ignore", divide="ignore"):
            ovr = (inter / ((area[i] + area[rest]).astype(f32) - inter).astype(f32)).astype(f32)
        dead[rest[ovr.astype(np.float64) > thr]] = True          # NaN compares False -> kept
    return np.asarray(keep, np.int64)


def detect_ref(levels: Sequence[np.ndarray], img_size: int, conf: float, iou: float,
               max_det_per_class: int = 300) -> List[Dict[str, np.ndarray]]:
    """Per image: {"boxes":[K,4] f32, "scores":[K] f32, "classes":[K] i64, "index":[K] i64 anchor index}."""
    dec = decode_ref(levels, img_size)
    out = []
    conf32 = f32(conf)
    for b in range(dec["box"].shape[0]):
        score, cls = score_ref(dec["obj"][b, :, 0], dec["cls"][b])
        m = np.nonzero(score > conf32)[0]
        bx, sc, cl = dec["box"][b][m], score[m], cls[m]
        fb, fs, fc, fi = [], [], [], []
        for c in np.unique(cl):
            mc = np.nonzero(cl == c)[0]
            k = nms_ref(bx[mc], sc[mc], iou)
            if max_det_per_class and max_det_per_class > 0:
                k = k[:max_det_per_class]
            fb.append(bx[mc][k]); fs.append(sc[mc][k]); fc.append(np.full(k.size, c, np.int64)); fi.append(m[mc][k])
        if fb:
            out.append({"boxes": np.concatenate(fb).astype(f32), "scores": np.concatenate(fs).astype(f32),
                        "classes": np.concatenate(fc), "index": np.concatenate(fi).astype(np.int64)})
        else:
            out.append({"boxes": np.zeros((0, 4), f32), "scores": np.zeros((0,), f32),
                        "classes": np.zeros((0,), np.int64), "index": np.zeros((0,), np.int64)})
    return out


def coco_dets_ref(levels: Sequence[np.ndarray], img_size: int, conf: float = 0.001, iou: float = 0.65,
                  add_one: bool = True) -> List[List[dict]]:
    """scripts/helpers/helpers.py:86-153 : no max_det, xyxy->xywh (helpers.py:58-83), category_id = cls+1."""
    res = []
    for d in detect_ref(levels, img_size, conf, iou, max_det_per_class=0):
        b = d["boxes"]
        w = np.maximum(b[:, 2] - b[:, 0], f32(0))
        h = np.maximum(b[:, 3] - b[:, 1], f32(0))
        cx = b[:, 0] + f32(0.5) * w
        cy = b[:, 1] + f32(0.5) * h
        res.append([{"category_id": int(c) + (1 if add_one else 0), "bbox": [float(a), float(bb), float(cc), float(dd)],
                     "score": float(s)} for a, bb, cc, dd, s, c in zip(cx, cy, w, h, d["scores"], d["classes"])])
    return res


# letterbox geometry + back-map (tools/infer.py:121-131, :507-516); image resampling itself is cv2's.
def letterbox_geometry(h: int, w: int, new_size: int) -> Tuple[float, int, int, int, int]:
    scale = min(new_size / h, new_size / w)
    nh, nw = int(round(h * scale)), int(round(w * scale))
    top = (new_size - nh) // 2
    left = (new_size - nw) // 2
    return scale, nh, nw, left, top


def backmap_ref(boxes: np.ndarray, scale: float, padx: int, pady: int, h0: int, w0: int) -> np.ndarray:
    b = np.array(boxes, f32, copy=True).reshape(-1, 4)
    b[:, [0, 2]] -= f32(padx)
    b[:, [1, 3]] -= f32(pady)
    b /= f32(max(scale, 1e-6))
    b[:, [0, 2]] = np.clip(b[:, [0, 2]], 0, w0 - 1)
    b[:, [1, 3]] = np.clip(b[:, [1, 3]], 0, h0 - 1)
    return b
