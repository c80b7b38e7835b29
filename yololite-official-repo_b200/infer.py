"""Checkpoint loading and the user-facing predict() facade.

  load_model_names_imgsize_from_ckpt   tools/infer.py:80-102 (same name, arguments, return triple, errors)
  letterbox_geometry / preprocess      tools/infer.py:121-131, :432-453 (GPU kernel, bit-exact to cv2 8-bit
                                       INTER_LINEAR + the numpy normalisation)
  YoloLite(...).predict()              README.md:22-42, benchmark.py:73-82,127-129 (results-dict layout:
                                       'boxes' xyxy ndarray, 'scores', 'classes', 'masks' None, 'speed' dict)
"""
from __future__ import annotations

import ctypes
import json
import os
import time
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib as L
from .engine import YoloLiteB200
from .post import PostProcessor, backmap


def load_model_names_imgsize_from_ckpt(weights: str, device) -> Tuple[YoloLiteB200, List[str], int]:
    ckpt = torch.load(weights, map_location="cpu", weights_only=False)
    if not (isinstance(ckpt, dict) and "state_dict" in ckpt and "meta" in ckpt):
        raise RuntimeError("Checkpoint saknar 'state_dict'/'meta'. Spara vikter via save_checkpoint_state(...).")
    meta = ckpt["meta"] or {}
    model = YoloLiteB200(ckpt["state_dict"], meta, device=device)
    names = meta.get("names") or [str(i) for i in range(int(meta.get("num_classes", 80)))]
    return model, names, int(meta.get("img_size", 640))


def letterbox_geometry(h: int, w: int, new_size: int):
    """(scale, nh, nw, left, top) exactly as tools/infer.py:121-131 computes them."""
    scale = min(new_size / h, new_size / w)
    nh, nw = int(round(h * scale)), int(round(w * scale))
    return scale, nh, nw, (new_size - nw) // 2, (new_size - nh) // 2


def preprocess(images_bgr: Sequence[Union[np.ndarray, torch.Tensor]], img_size: int, device, letterbox: bool = True,
               out: Optional[torch.Tensor] = None):
    """uint8 HWC BGR images -> normalised fp32 [B,3,S,S] on `device` + per-image (scale, padx, pady, h0, w0)."""
    dev = torch.device(device)
    B = len(images_bgr)
    x = out if out is not None else torch.empty((B, 3, img_size, img_size), device=dev, dtype=torch.float32)
    geo = []
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    keep = []
    for i, im in enumerate(images_bgr):
        t = torch.from_numpy(np.ascontiguousarray(im)) if isinstance(im, np.ndarray) else im
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise ValueError("images must be uint8 HxWx3 (BGR)")
        h0, w0 = int(t.shape[0]), int(t.shape[1])
        if letterbox:
            scale, nh, nw, left, top = letterbox_geometry(h0, w0, img_size)
        else:   # --no_letterbox: plain resize to the square (tools/infer.py:443-445)
            scale, nh, nw, left, top = min(img_size / h0, img_size / w0), img_size, img_size, 0, 0
        td = t.to(dev, non_blocking=True).contiguous()
        keep.append(td)
        L.check(L.lib().yl_preprocess(td.data_ptr(), h0, w0, w0 * 3, x[i].data_ptr(), img_size, nh, nw, left, top, stream))
        geo.append((scale, left, top, h0, w0))
    return x, geo


def preprocess_batch(images_u8: torch.Tensor, img_size: int, out: Optional[torch.Tensor] = None, letterbox: bool = True):
    """A CUDA uint8 [B,H,W,3] BGR batch (equal sizes) -> normalised fp32 [B,3,S,S] in ONE launch + the shared geometry."""
    if not (images_u8.is_cuda and images_u8.dtype == torch.uint8 and images_u8.dim() == 4 and images_u8.shape[3] == 3):
        raise ValueError("expected a CUDA uint8 tensor [B,H,W,3]")
    images_u8 = images_u8.contiguous()
    B, h0, w0, _ = images_u8.shape
    if letterbox:
        scale, nh, nw, left, top = letterbox_geometry(h0, w0, img_size)
    else:
        scale, nh, nw, left, top = min(img_size / h0, img_size / w0), img_size, img_size, 0, 0
    x = out if out is not None else torch.empty((B, 3, img_size, img_size), device=images_u8.device, dtype=torch.float32)
    stream = ctypes.c_void_p(torch.cuda.current_stream(images_u8.device).cuda_stream)
    L.check(L.lib().yl_preprocess_batch(images_u8.data_ptr(), B, h0, w0, x.data_ptr(), img_size, nh, nw, left, top, stream))
    return x, (scale, left, top, h0, w0)


class YoloLite:
    """`YoloLite(weights).predict(source)` -> list of result dicts (README.md:22-42)."""

    def __init__(self, weights: str, device="cuda:0", graph: bool = True):
        self.model, self.names, self.img_size = load_model_names_imgsize_from_ckpt(weights, torch.device(device))
        self.model.set_option("graph", 1 if graph else 0)
        self.device = torch.device(device)
        self._post = PostProcessor()

    def predict(self, source, device=None, draw: bool = False, conf: float = 0.4, iou: float = 0.5, max_det: int = 300,
                img_size: int = 0) -> List[dict]:
        """max_det is the PER-CLASS cap (`keep[:max_det]` inside the class loop of tools/infer.py:134-152,476-493); the reference
        CLI always uses 300 there whatever its --max_det flag says, which is this default."""
        import cv2
        if device is not None and torch.device(device).type != "cuda":
            raise RuntimeError("yololite_b200 has no CPU path (device must be a CUDA device)")
        S = int(img_size) if img_size else self.img_size
        t0 = time.perf_counter()
        if isinstance(source, (str, os.PathLike)):
            paths = [str(source)]
            imgs = [cv2.imread(paths[0])]
            if imgs[0] is None:
                raise ValueError(f"could not read {source}")
        elif isinstance(source, np.ndarray):
            paths, imgs = [None], [source]
        else:
            paths, imgs = [], []
            for s in source:
                if isinstance(s, np.ndarray):
                    paths.append(None); imgs.append(s)
                    continue
                im = cv2.imread(str(s))
                if im is None:                          # tools/infer.py:437-440: warn and skip unreadable images
                    print(f"⚠️  Kunde inte läsa {s}")
                    continue
                paths.append(str(s)); imgs.append(im)
            if not imgs:
                return []
        x, geo = preprocess(imgs, S, self.device)
        torch.cuda.synchronize(self.device)
        t1 = time.perf_counter()
        levels = self.model(x)
        torch.cuda.synchronize(self.device)
        t2 = time.perf_counter()
        dets = self._post(levels, S, conf, iou, max_det).to_list()
        results = []
        for d, (scale, padx, pady, h0, w0), pth in zip(dets, geo, paths):
            boxes = backmap(d["boxes"], scale, padx, pady, h0, w0)
            results.append({"boxes": boxes.cpu().numpy(), "scores": d["scores"].cpu().numpy(),
                            "classes": d["classes"].cpu().numpy(), "masks": None, "path": pth})
        t3 = time.perf_counter()
        n = max(1, len(results))
        speed = {"pre_ms": (t1 - t0) * 1e3 / n, "infer_ms": (t2 - t1) * 1e3 / n, "post_ms": (t3 - t2) * 1e3 / n,
                 "total_ms": (t3 - t0) * 1e3 / n}
        for r in results:
            r["speed"] = dict(speed)
        return results

    def predict_batch(self, images_u8: torch.Tensor, conf: float = 0.4, iou: float = 0.5, max_det: int = 300, img_size: int = 0,
                      cap: Optional[int] = None, packed: Optional[torch.Tensor] = None):
        """Batched device-side predict: uint8 [B,H,W,3] BGR (CUDA) -> Detections (fixed capacity, letterboxed coordinates)
        + the letterbox geometry.  No host synchronisation; call `.to_list()` / `backmap` on the result when needed.  The
        Detections alias buffers owned by this object: valid until the next predict_batch call with the same shape.
        packed: optional [B, cap+1, 6] tensor the kernel fills as well (the multi-GPU gather payload, dist.gather_packed)."""
        from .post import Detections
        S = int(img_size) if img_size else self.img_size
        B, h0, w0 = images_u8.shape[0], images_u8.shape[1], images_u8.shape[2]
        direct = h0 == S and w0 == S and self.model.supports_u8(S, S)
        if direct:
            # no letterbox resize / padding needed: the stem kernel reads the uint8 image itself (normalisation folded in)
            x, geo = images_u8, (1.0, 0, 0, h0, w0)
        else:
            key = (tuple(images_u8.shape), S)
            if getattr(self, "_xbuf_key", None) != key:
                self._xbuf = torch.empty((images_u8.shape[0], 3, S, S), device=images_u8.device, dtype=torch.float32)
                self._xbuf_key = key
            x, geo = preprocess_batch(images_u8, S, out=self._xbuf)
        cap = int(cap) if cap else 1024
        okey = (B, cap, images_u8.device)
        if getattr(self, "_obuf_key", None) != okey:
            dev = images_u8.device
            self._obuf = (torch.empty((B, cap, 4), device=dev), torch.empty((B, cap), device=dev),
                          torch.empty((B, cap), device=dev, dtype=torch.int64), torch.empty((B, cap), device=dev, dtype=torch.int64),
                          torch.zeros((B,), device=dev, dtype=torch.int32))
            self._obuf_key = okey
        # forward + postprocess as ONE C call (one CUDA graph launch once the engine has captured it)
        bx, sc, cl, ix, cn = self.model.detect(x, S, conf, iou, max_det, cap, outputs=self._obuf, packed=packed)
        return Detections(bx, sc, cl, ix, cn), geo

    def to_json(self, result: dict) -> dict:
        """The per-image JSON record tools/infer.py:540-549 writes."""
        rec = [{"bbox_xyxy": [float(v) for v in b], "score": float(s), "class_id": int(c),
                "class_name": self.names[int(c)] if int(c) < len(self.names) else str(int(c))}
               for b, s, c in zip(result["boxes"].tolist(), result["scores"].tolist(), result["classes"].tolist())]
        return {"image": result.get("path"), "detections": rec}
