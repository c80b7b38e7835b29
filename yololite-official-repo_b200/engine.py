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


class _DevMem:
    """A raw device pointer as a __cuda_array_interface__ object (fp32), so torch can view engine-owned memory."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 3}


class YoloLiteB200:
    def __init__(self, state_dict: dict, meta: dict, device="cuda:0", fuse_dwpw: bool = True,
                 reuse_buffers: bool = True, tensor_cores: bool = True, fuse_stem: bool = True, from_features: bool = False,
                 graph: bool = False, pdl: bool = True):
        """from_features: build only the FPN + heads; ``forward_features([c2,] c3, c4, c5)`` then takes the backbone's
        feature maps (any timm backbone run elsewhere).  graph: replay each distinct call as one CUDA graph launch.
        pdl: programmatic dependent launch between the tcgen05 kernels."""
        lib = L.lib()                                   # raises ImportError if the extension is not built
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("yololite_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device visible: yololite_b200 has no CPU fallback")
        self.meta = meta
        self.from_features = bool(from_features)
        self.program = packer.lower(state_dict, meta, fuse_dwpw=fuse_dwpw, reuse_buffers=reuse_buffers,
                                    tensor_cores=tensor_cores, fuse_stem=fuse_stem, from_features=from_features)
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
        self.device = torch.device("cuda", idx)
        L.check(lib.yl_engine_set_option(h, b"tensor_cores", 1 if tensor_cores else 0))
        L.check(lib.yl_engine_set_option(h, b"pdl", 1 if pdl else 0))
        L.check(lib.yl_engine_set_option(h, b"graph", 1 if graph else 0))
        self._shape_cache = {}

    def set_option(self, key: str, value: int):
        L.check(L.lib().yl_engine_set_option(self._h, key.encode(), int(value)))

    # ---- reference-compatible surface
    def eval(self):
        return self

    def to(self, device):
        d = torch.device(device)
        if d.type != "cuda":
            raise RuntimeError("yololite_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if d.index not in (None, self._dev_index):
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

    def _outputs(self, shapes, B, out, dev):
        """Fresh level tensors, or the caller's after checking that the raw pointers handed to the kernels are what they expect."""
        if out is None:
            return [torch.empty((B, A, sh, sw, D), device=dev, dtype=torch.float32) for (A, sh, sw, D) in shapes]
        out = list(out)
        if len(out) != len(shapes):
            raise ValueError(f"out must hold {len(shapes)} level tensors")
        for o, (A, sh, sw, D) in zip(out, shapes):
            if not (isinstance(o, torch.Tensor) and o.is_cuda and o.device.index == self._dev_index and o.dtype == torch.float32
                    and tuple(o.shape) == (B, A, sh, sw, D) and o.is_contiguous()):
                raise ValueError(f"out tensors must be contiguous CUDA float32 of shape {(B, A, sh, sw, D)} on cuda:{self._dev_index}")
        return out

    def _finish(self, out, B):
        if self.export_concat:
            return torch.cat([o.view(B, -1, o.shape[-1]) for o in out], dim=1)
        return out

    def forward(self, x: torch.Tensor, out: Optional[Sequence[torch.Tensor]] = None):
        if self.from_features:
            raise RuntimeError("this engine was built from_features: call forward_features([c2,] c3, c4, c5)")
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
            raise ValueError("expected a CUDA float32 tensor of shape [B,3,H,W]")
        if x.device.index != self._dev_index:
            raise ValueError(f"input is on {x.device}, engine on cuda:{self._dev_index}")
        x = x.contiguous()
        B, _, H, W = x.shape
        out = self._outputs(self.level_shapes(B, H, W), B, out, x.device)
        ptrs = (ctypes.c_void_p * self._n_levels)(*[o.data_ptr() for o in out])
        L.check(L.lib().yl_forward(self._h, ctypes.c_void_p(x.data_ptr()), B, H, W, ptrs, _stream_ptr(x.device)))
        return self._finish(out, B)

    __call__ = forward

    def forward_u8(self, images: torch.Tensor, out: Optional[Sequence[torch.Tensor]] = None):
        """model(normalise(images)) for a CUDA uint8 batch [B,S,S,3] in BGR (cv2 order) that needs no letterbox resize:
        BGR->RGB, /255, (x-mean)/std (tools/infer.py:442-453) are folded into the stem kernel, the fp32 image never
        exists.  Raises ValueError when the engine / shape has no uint8 entry (use preprocess_batch + forward then)."""
        if not (isinstance(images, torch.Tensor) and images.is_cuda and images.dtype == torch.uint8 and images.dim() == 4
                and images.shape[3] == 3):
            raise ValueError("expected a CUDA uint8 tensor of shape [B,H,W,3] (BGR)")
        if images.device.index != self._dev_index:
            raise ValueError(f"input is on {images.device}, engine on cuda:{self._dev_index}")
        images = images.contiguous()
        B, H, W, _ = images.shape
        out = self._outputs(self.level_shapes(B, H, W), B, out, images.device)
        ptrs = (ctypes.c_void_p * self._n_levels)(*[o.data_ptr() for o in out])
        L.check(L.lib().yl_forward_u8(self._h, ctypes.c_void_p(images.data_ptr()), B, H, W, ptrs, _stream_ptr(images.device)))
        return self._finish(out, B)

    def supports_u8(self, H: int, W: int) -> bool:
        if self.from_features:
            return False
        op0 = self.program.ops[0]
        return bool(op0["kind"] == L.OP_STEM2 and op0["w3_off"] >= 0 and H % 2 == 0 and W % 16 == 0)

    # ---- FPN + heads on backbone features (model_v2.py:195-224 / :353-377 without the timm call)
    def forward_features(self, feats: Sequence[torch.Tensor], out: Optional[Sequence[torch.Tensor]] = None):
        """feats: what the reference's ``self.backbone(x)`` returns -- [c2,] c3, c4, c5 as [B,C,H,W] fp32 CUDA tensors.  Tensors in
        torch.channels_last memory format are read in place (that IS the engine's NHWC layout); others are converted once."""
        if not self.from_features:
            raise RuntimeError("build the engine with from_features=True to feed backbone features")
        chs = self.program.feature_channels
        if len(feats) != len(chs):
            raise ValueError(f"expected {len(chs)} feature maps")
        fl = []
        for f, c in zip(feats, chs):
            if not (isinstance(f, torch.Tensor) and f.is_cuda and f.dtype == torch.float32 and f.dim() == 4 and f.shape[1] == c
                    and f.device.index == self._dev_index):
                raise ValueError(f"features must be CUDA float32 [B,{c},H,W] tensors on cuda:{self._dev_index}")
            fl.append(f.contiguous(memory_format=torch.channels_last))
        B = int(fl[0].shape[0])
        dims = (ctypes.c_int32 * (3 * len(fl)))(*[int(v) for f in fl for v in (f.shape[2], f.shape[3], f.shape[1])])
        key = ("feat", B) + tuple(dims)
        if key not in self._shape_cache:
            shp = (ctypes.c_int32 * (4 * self._n_levels))()
            L.check(L.lib().yl_engine_plan_features(self._h, B, dims, len(fl), shp))
            self._shape_cache[key] = [tuple(shp[l * 4:l * 4 + 4]) for l in range(self._n_levels)]
        out = self._outputs(self._shape_cache[key], B, out, fl[0].device)
        fptrs = (ctypes.c_void_p * len(fl))(*[f.data_ptr() for f in fl])
        ptrs = (ctypes.c_void_p * self._n_levels)(*[o.data_ptr() for o in out])
        L.check(L.lib().yl_forward_features(self._h, fptrs, dims, len(fl), B, ptrs, _stream_ptr(fl[0].device)))
        self._keep = fl                                  # converted copies stay alive until the next call
        return self._finish(out, B)

    # ---- model(x) + postprocess as ONE C call (one CUDA graph launch with graph=True)
    def detect(self, x: torch.Tensor, img_size: int, conf: float = 0.4, iou: float = 0.5, max_det: int = 300, cap: int = 1024,
               outputs=None, packed: Optional[torch.Tensor] = None):
        """tools/infer.py:456-493 for a batch.  x: fp32 [B,3,H,W] normalised, or uint8 [B,H,W,3] BGR (no-resize image entry).
        Returns (boxes [B,cap,4], scores [B,cap], classes i64, index i64, counts i32) -- pass `outputs` (same tuple) to reuse
        buffers.  `packed` [B,cap+1,6] (see yl_postprocess_ex) is filled too when given; with `packed` and no `outputs` only the
        payload is written and returned."""
        if self.from_features:
            raise RuntimeError("detect() needs the full network")
        u8 = x.dtype == torch.uint8
        if not (x.is_cuda and x.dim() == 4 and x.device.index == self._dev_index and (x.shape[3] == 3 if u8 else (x.dtype == torch.float32 and x.shape[1] == 3))):
            raise ValueError("expected CUDA float32 [B,3,H,W] or uint8 [B,H,W,3] on the engine's device")
        x = x.contiguous()
        B = int(x.shape[0])
        H, W = (int(x.shape[1]), int(x.shape[2])) if u8 else (int(x.shape[2]), int(x.shape[3]))
        dev = x.device
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        bx = sc = cl = ix = cn = None
        if packed is not None:
            if not (packed.is_cuda and packed.dtype == torch.float32 and tuple(packed.shape) == (B, cap + 1, 6) and packed.is_contiguous()):
                raise ValueError(f"packed must be a contiguous CUDA float32 tensor of shape {(B, cap + 1, 6)}")
        if outputs is not None:
            bx, sc, cl, ix, cn = outputs
            want = (((B, cap, 4), torch.float32), ((B, cap), torch.float32), ((B, cap), torch.int64), ((B, cap), torch.int64), ((B,), torch.int32))
            for t, (shp, dt) in zip(outputs, want):
                if not (t.is_cuda and t.dtype == dt and tuple(t.shape) == shp and t.is_contiguous() and t.device.index == self._dev_index):
                    raise ValueError(f"output buffer must be contiguous CUDA {dt} of shape {shp}")
        elif packed is None:
            bx = torch.empty((B, cap, 4), device=dev, dtype=torch.float32)
            sc = torch.empty((B, cap), device=dev, dtype=torch.float32)
            cl = torch.empty((B, cap), device=dev, dtype=torch.int64)
            ix = torch.empty((B, cap), device=dev, dtype=torch.int64)
            cn = torch.empty((B,), device=dev, dtype=torch.int32)
        L.check(L.lib().yl_engine_detect(self._h, None if u8 else ptr(x), ptr(x) if u8 else None, B, H, W, int(img_size), float(conf),
                                         float(iou), int(max_det or 0), int(cap), ptr(bx), ptr(sc), ptr(cl), ptr(ix), ptr(cn), ptr(packed),
                                         _stream_ptr(dev)))
        self._last_detect_B = B
        return (bx, sc, cl, ix, cn) if bx is not None else packed

    def last_levels(self) -> List[torch.Tensor]:
        """Copies of the engine-owned logits written by the last detect() (one [B,A,S,S,5+C] tensor per level)."""
        n = self._n_levels
        ptrs = (ctypes.c_void_p * n)()
        shp = (ctypes.c_int32 * (4 * n))()
        L.check(L.lib().yl_engine_levels(self._h, ptrs, shp))
        B = self._last_detect_B
        out = []
        for l in range(n):
            A, sh, sw, D = (int(v) for v in shp[l * 4:l * 4 + 4])
            out.append(torch.as_tensor(_DevMem(ptrs[l], (B, A, sh, sw, D)), device=self.device).clone())
        return out

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
