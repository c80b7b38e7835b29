"""Host-side mirror of the reference postprocess functions on top of the fused CUDA kernel.

  decode_preds_anchorfree   scripts/helpers/utils_ms.py:25-123  (same name, arguments and result dict)
  detect                    tools/infer.py:460-493              (score, threshold, per-class NMS)
  decode_batch_to_coco_dets scripts/helpers/helpers.py:86-153   (evaluation variant -> COCO det dicts)
  backmap                   tools/infer.py:507-516
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L

OVERFLOW_BIT = 1 << 30


def _levels5(levels) -> List[torch.Tensor]:
    levels = list(levels) if isinstance(levels, (list, tuple)) else [levels]
    out = []
    for p in levels:
        if not (isinstance(p, torch.Tensor) and p.is_cuda and p.dtype == torch.float32):
            raise ValueError("levels must be CUDA float32 tensors")
        if p.dim() == 4:
            p = p.unsqueeze(1)
        if p.dim() != 5:
            raise ValueError("each level must be [B,A,S,S,5+C] or [B,S,S,5+C]")
        out.append(p.contiguous())
    B = out[0].shape[0]
    for p in out:
        assert p.shape[0] == B, "Batch mismatch mellan nivåer"
        if p.shape[-1] != out[0].shape[-1]:
            raise ValueError("all levels must share 5+C")
    return out


def _level_args(levels: List[torch.Tensor]):
    n = len(levels)
    ptrs = (ctypes.c_void_p * n)(*[p.data_ptr() for p in levels])
    dims = (ctypes.c_int32 * (3 * n))(*[int(v) for p in levels for v in (p.shape[1], p.shape[2], p.shape[3])])
    return ptrs, dims


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


@torch.no_grad()
def decode_preds_anchorfree(preds_levels, img_size: int, center_mode: str = "v8", wh_mode: str = "softplus") -> Dict[str, torch.Tensor]:
    if center_mode != "v8" or wh_mode != "softplus":
        raise ValueError("only center_mode='v8', wh_mode='softplus' (the modes every reference caller uses) are lowered")
    lv = _levels5(preds_levels)
    B, D = lv[0].shape[0], lv[0].shape[-1]
    N = sum(p.shape[1] * p.shape[2] * p.shape[3] for p in lv)
    dev = lv[0].device
    box = torch.empty((B, N, 4), device=dev, dtype=torch.float32)
    obj = torch.empty((B, N, 1), device=dev, dtype=torch.float32)
    cls = torch.empty((B, N, D - 5), device=dev, dtype=torch.float32)
    ptrs, dims = _level_args(lv)
    L.check(L.lib().yl_decode(ptrs, dims, len(lv), B, D, int(img_size), box.data_ptr(), obj.data_ptr(),
                              cls.data_ptr() if D > 5 else None, _stream(dev)))
    return {"box": box, "obj": obj, "cls": cls}


class Detections:
    """Fixed-capacity device-side result of one postprocess launch.  The tensors belong to the PostProcessor / caller that
    produced them and are overwritten by its next call; `to_list()` returns independent copies."""

    def __init__(self, boxes, scores, classes, index, counts):
        self.boxes, self.scores, self.classes, self.index, self.counts = boxes, scores, classes, index, counts

    def to_list(self, copy: bool = True) -> List[Dict[str, torch.Tensor]]:
        """Per image (boxes [K,4] f32, scores [K] f32, classes [K] i64, index [K] i64) sliced by the counts (one D2H sync).
        copy=True (default) clones the slices, like the fresh tensors the reference returns (tools/infer.py:476-493);
        copy=False returns views that the next call on the same PostProcessor overwrites."""
        cnt = self.counts.cpu().tolist()
        out = []
        for b, c in enumerate(cnt):
            if c & OVERFLOW_BIT:
                raise RuntimeError(f"image {b}: more detections than the output capacity {self.boxes.shape[1]}")
            d = {"boxes": self.boxes[b, :c], "scores": self.scores[b, :c], "classes": self.classes[b, :c], "index": self.index[b, :c]}
            out.append({k: v.clone() for k, v in d.items()} if copy else d)
        return out


def unpack(packed: torch.Tensor) -> List[Dict[str, torch.Tensor]]:
    """[B, cap+1, 6] payload written by the kernel (row 0 = count, overflow flag, K; rows 1.. = x1,y1,x2,y2,score,class)
    -> per-image dicts like tools/infer.py produces (one D2H sync for the header rows)."""
    head = packed[:, 0, :3].cpu()
    res = []
    for b in range(packed.shape[0]):
        if head[b, 1] != 0:
            raise RuntimeError(f"image {b}: {int(head[b, 2])} detections exceed the capacity {packed.shape[1] - 1}")
        c = int(head[b, 0])
        p = packed[b, 1:1 + c]
        res.append({"boxes": p[:, :4].clone(), "scores": p[:, 4].clone(), "classes": p[:, 5].to(torch.int64)})
    return res


class PostProcessor:
    """Owns the scratch + output buffers for a fixed (B, N, cap) so repeated calls allocate nothing.  The returned Detections
    alias those buffers: they are valid until the next call on this object (use one PostProcessor per stream / thread)."""

    def __init__(self):
        self._key = None

    def _ensure(self, B, N, cap, dev):
        key = (B, N, cap, dev)
        if key != self._key:
            nbytes = L.lib().yl_postprocess_scratch_bytes(B, N)
            self.scratch = torch.empty((nbytes + 255) // 256 * 256, device=dev, dtype=torch.uint8)
            self.scratch_bytes = nbytes
            self.boxes = torch.empty((B, cap, 4), device=dev, dtype=torch.float32)
            self.scores = torch.empty((B, cap), device=dev, dtype=torch.float32)
            self.classes = torch.empty((B, cap), device=dev, dtype=torch.int64)
            self.index = torch.empty((B, cap), device=dev, dtype=torch.int64)
            self.counts = torch.zeros((B,), device=dev, dtype=torch.int32)
            self._key = key

    @torch.no_grad()
    def __call__(self, preds_levels, img_size: int, conf: float = 0.4, iou: float = 0.5, max_det: int = 300,
                 cap: Optional[int] = None, packed: Optional[torch.Tensor] = None) -> Detections:
        """max_det is the PER-CLASS cap of tools/infer.py:134-152 (`keep[:max_det]` inside the class loop; the reference CLI
        always passes 300 there, its --max_det flag is not forwarded on this branch), 0 = unlimited (helpers.py:86-153).
        packed: optional [B, cap+1, 6] fp32 tensor the kernel additionally fills (see `unpack`)."""
        lv = _levels5(preds_levels)
        B, D = lv[0].shape[0], lv[0].shape[-1]
        N = sum(p.shape[1] * p.shape[2] * p.shape[3] for p in lv)
        cap = int(cap) if cap else N
        dev = lv[0].device
        self._ensure(B, N, cap, dev)
        if packed is not None and not (packed.is_cuda and packed.dtype == torch.float32 and tuple(packed.shape) == (B, cap + 1, 6)
                                       and packed.is_contiguous()):
            raise ValueError(f"packed must be a contiguous CUDA float32 tensor of shape {(B, cap + 1, 6)}")
        ptrs, dims = _level_args(lv)
        L.check(L.lib().yl_postprocess_ex(ptrs, dims, len(lv), B, D, int(img_size), float(conf), float(iou), int(max_det or 0),
                                          cap, self.boxes.data_ptr(), self.scores.data_ptr(), self.classes.data_ptr(),
                                          self.index.data_ptr(), self.counts.data_ptr(),
                                          packed.data_ptr() if packed is not None else None, self.scratch.data_ptr(),
                                          self.scratch_bytes, _stream(dev)))
        return Detections(self.boxes, self.scores, self.classes, self.index, self.counts)


_default = PostProcessor()


def detect(preds_levels, img_size: int, conf: float = 0.4, iou: float = 0.5, max_det: int = 300,
           cap: Optional[int] = None) -> List[Dict[str, torch.Tensor]]:
    """tools/infer.py:460-493 for a whole batch; defaults are the CLI's (conf 0.4, iou 0.5, 300 per class).  Returns fresh
    tensors (the shared PostProcessor's buffers are copied out)."""
    return _default(preds_levels, img_size, conf, iou, max_det, cap).to_list(copy=True)


def decode_batch_to_coco_dets(preds, img_size, conf_th=0.001, iou_th=0.65, add_one=True) -> List[List[dict]]:
    """scripts/helpers/helpers.py:86-153: no max_det, xyxy -> xywh (helpers.py:58-83), category_id = cls + 1."""
    res = []
    for d in detect(preds, img_size, conf_th, iou_th, max_det=0):
        b = d["boxes"]
        w = (b[:, 2] - b[:, 0]).clamp_min(0)
        h = (b[:, 3] - b[:, 1]).clamp_min(0)
        xywh = torch.stack([b[:, 0] + 0.5 * w, b[:, 1] + 0.5 * h, w, h], dim=-1).cpu().tolist()
        sc = d["scores"].cpu().tolist()
        cc = (d["classes"] + (1 if add_one else 0)).cpu().tolist()
        res.append([{"category_id": int(c), "bbox": [float(v) for v in bx], "score": float(s)}
                    for bx, s, c in zip(xywh, sc, cc)])
    return res


def backmap(boxes: torch.Tensor, scale: float, padx: int, pady: int, h0: int, w0: int) -> torch.Tensor:
    """tools/infer.py:507-516: undo the letterbox and clip to the original image."""
    b = boxes.clone()
    b[:, [0, 2]] -= padx
    b[:, [1, 3]] -= pady
    b /= max(scale, 1e-6)
    b[:, [0, 2]] = b[:, [0, 2]].clamp(0, w0 - 1)
    b[:, [1, 3]] = b[:, [1, 3]].clamp(0, h0 - 1)
    return b
