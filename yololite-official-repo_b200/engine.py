"""Host-side mirror of the reference model object (scripts/model/model_v2.py:250-399) on top of the C ABI.

``YoloLiteB200`` is duck-type compatible with ``YOLOLiteMS_CPU`` / ``YOLOLiteMS`` for inference callers
(tools/infer.py:456, scripts/helpers/evaluate.py:273,290,423): ``model(x)`` takes fp32 ``[B,3,H,W]`` NCHW on
the model's CUDA device and returns a list of new contiguous ``[B,A,S,S,5+C]`` tensors (or one ``[B,N,5+C]``
when ``export_concat`` is set, model_v2.py:57-64).  PyTorch only supplies device memory and the stream.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import _lib as L
from . import packer


def _stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class YoloLiteB200:
    def __init__(self, state_dict: dict, meta: dict, device="cuda:0", fuse_dwpw: bool = True,
                 reuse_buffers: bool = True, tensor_cores: bool = True, fuse_stem: bool = True):
        lib = L.lib()                                   # raises ImportError if the extension is not built
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("yololite_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device visible: yololite_b200 has no CPU fallback")
        self.meta = meta
        self.program = packer.lower(state_dict, meta, fuse_dwpw=fuse_dwpw, reuse_buffers=reuse_buffers,
                                    tensor_cores=tensor_cores, fuse_stem=fuse_stem)
        cfg = self.program.cfg
        self.cfg = cfg
        self.num_classes = cfg.num_classes
        self.use_p2, self.use_p6 = cfg.use_p2, cfg.use_p6
        self.num_anchors_per_level = tuple(cfg.anchors)
        self.fpn_strides = list(self.program.strides)
        self.export_concat = False
        self.export_decode = False
        self.names = cfg.names
        self.training = False
        ops, blob = packer.to_c(self.program)
        self._n_levels = len(cfg.levels)
        h = ctypes.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        L.check(lib.yl_engine_create(ops, len(ops), blob.ctypes.data_as(ctypes.c_void_p), blob.size,
                                     self.program.n_buffers, self._n_levels, idx, ctypes.byref(h)))
        self._h = h
        self._dev_index = idx
        L.check(lib.yl_engine_set_option(h, b"tensor_cores", 1 if tensor_cores else 0))
        self._shape_cache = {}

    # ---- reference-compatible surface
    def eval(self):
        return self

    def to(self, device):
        if torch.device(device) != self.device and torch.device(device).index not in (None, self._dev_index):
            raise RuntimeError("a YoloLiteB200 engine is bound to its device; build another one for " + str(device))
        return self

    def get_strides(self) -> List[int]:
        return list(self.fpn_strides)

    def get_num_anchors_per_level(self):
        return tuple(self.num_anchors_per_level)

    def level_shapes(self, B: int, H: int, W: int):
        key = (B, H, W)
        if key not in self._shape_cache:
            shp = (ctypes.c_int32 * (4 * self._n_levels))()
            L.check(L.lib().yl_engine_plan(self._h, B, H, W, shp))
            self._shape_cache[key] = [tuple(shp[l * 4:l * 4 + 4]) for l in range(self._n_levels)]
        return self._shape_cache[key]

    def forward(self, x: torch.Tensor, out: Optional[Sequence[torch.Tensor]] = None):
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
            raise ValueError("expected a CUDA float32 tensor of shape [B,3,H,W]")
        if x.device.index != self._dev_index:
            raise ValueError(f"input is on {x.device}, engine on cuda:{self._dev_index}")
        x = x.contiguous()
        B, _, H, W = x.shape
        shapes = self.level_shapes(B, H, W)
        if out is None:
            out = [torch.empty((B, A, sh, sw, D), device=x.device, dtype=torch.float32) for (A, sh, sw, D) in shapes]
        ptrs = (ctypes.c_void_p * self._n_levels)(*[o.data_ptr() for o in out])
        L.check(L.lib().yl_forward(self._h, ctypes.c_void_p(x.data_ptr()), B, H, W, ptrs, _stream_ptr(x.device)))
        out = list(out)
        if self.export_concat:
            return torch.cat([o.view(B, -1, o.shape[-1]) for o in out], dim=1)
        return out

    __call__ = forward

    def forward_u8(self, images: torch.Tensor, out: Optional[Sequence[torch.Tensor]] = None):
        """model(normalise(images)) for a CUDA uint8 batch [B,S,S,3] in BGR (cv2 order) that needs no letterbox resize:
        BGR->RGB, /255, (x-mean)/std (tools/infer.py:442-453) are folded into the stem kernel, the fp32 image never
        exists.  Raises ValueError when the engine / shape has no uint8 entry (use preprocess_batch + forward then)."""
        if not (isinstance(images, torch.Tensor) and images.is_cuda and images.dtype == torch.uint8 and images.dim() == 4
                and images.shape[3] == 3):
            raise ValueError("expected a CUDA uint8 tensor of shape [B,H,W,3] (BGR)")
        images = images.contiguous()
        B, H, W, _ = images.shape
        shapes = self.level_shapes(B, H, W)
        if out is None:
            out = [torch.empty((B, A, sh, sw, D), device=images.device, dtype=torch.float32) for (A, sh, sw, D) in shapes]
        ptrs = (ctypes.c_void_p * self._n_levels)(*[o.data_ptr() for o in out])
        L.check(L.lib().yl_forward_u8(self._h, ctypes.c_void_p(images.data_ptr()), B, H, W, ptrs, _stream_ptr(images.device)))
        out = list(out)
        if self.export_concat:
            return torch.cat([o.view(B, -1, o.shape[-1]) for o in out], dim=1)
        return out

    def supports_u8(self, H: int, W: int) -> bool:
        op0 = self.program.ops[0]
        return bool(op0["kind"] == L.OP_STEM2 and op0["w3_off"] >= 0 and H % 2 == 0 and W % 16 == 0)

    def profile_ops(self, x: torch.Tensor):
        """Per-op device milliseconds of one forward (CUDA events between launches) -> list of (op dict, ms)."""
        x = x.contiguous()
        B, _, H, W = x.shape
        out = [torch.empty((B, A, sh, sw, D), device=x.device, dtype=torch.float32) for (A, sh, sw, D) in self.level_shapes(B, H, W)]
        ptrs = (ctypes.c_void_p * self._n_levels)(*[o.data_ptr() for o in out])
        n = len(self.program.ops)
        ms = (ctypes.c_float * n)()
        L.check(L.lib().yl_forward_profile(self._h, ctypes.c_void_p(x.data_ptr()), B, H, W, ptrs, _stream_ptr(x.device), ms, n))
        return [(op, float(t)) for op, t in zip(self.program.ops, ms)]

    def read_buffer(self, name_or_id, B: int) -> torch.Tensor:
        """Parity tap: NHWC activation of the last forward (build the engine with reuse_buffers=False)."""
        bid = self.program.taps[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
        dims = (ctypes.c_int32 * 3)()
        L.check(L.lib().yl_engine_read_buffer(self._h, bid, None, dims, None))
        t = torch.empty((B, dims[0], dims[1], dims[2]), device=self.device, dtype=torch.float32)
        L.check(L.lib().yl_engine_read_buffer(self._h, bid, ctypes.c_void_p(t.data_ptr()), dims, _stream_ptr(self.device)))
        return t

    def close(self):
        if getattr(self, "_h", None):
            L.lib().yl_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
