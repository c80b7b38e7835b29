"""Lower a reference checkpoint {"state_dict", "meta"} into the engine's layer program.

Reads the same files the reference reads:
  * meta layout                    tools/train.py:62-75 (save_checkpoint_state), tools/infer.py:34-77
  * FPN / head module tree         scripts/model/model_v2.py:250-377 (YOLOLiteMS_CPU), :77-224 (YOLOLiteMS)
  * backbone                       timm `mobilenetv4_conv_small[_050]` as called at model_v2.py:266-272; block
                                   structure as dumped at YoloLite_custom_training.ipynb:392-850

and emits (a) a flat op list over numbered NHWC activation buffers, (b) one fp32 blob with BatchNorm folded
into the preceding conv (done in float64, rounded once), GEMM weights stored [K][N] with N padded to 4.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib as L

BN_EPS = 1e-5

# ("cn", k, stride, c) | ("uir", dw_start_k, dw_mid_k, stride, expand, c)
_MNV4_SMALL = (
    (("cn", 3, 2, 32), ("cn", 1, 1, 32)),
    (("cn", 3, 2, 96), ("cn", 1, 1, 64)),
    (("uir", 5, 5, 2, 3.0, 96),) + (("uir", 0, 3, 1, 2.0, 96),) * 4 + (("uir", 3, 0, 1, 4.0, 96),),
    (("uir", 3, 3, 2, 6.0, 128), ("uir", 5, 5, 1, 4.0, 128), ("uir", 0, 5, 1, 4.0, 128),
     ("uir", 0, 5, 1, 3.0, 128), ("uir", 0, 3, 1, 4.0, 128), ("uir", 0, 3, 1, 4.0, 128)),
    (("cn", 1, 1, 960),),
)
BACKBONES = {"mobilenetv4_conv_small": (_MNV4_SMALL, 1.0, 32), "mobilenetv4_conv_small_050": (_MNV4_SMALL, 0.5, 32)}


def _round_ch(v: float, div: int = 8) -> int:
    n = max(div, int(v + div / 2) // div * div)
    return n + div if n < 0.9 * v else n


@dataclass
class ModelCfg:
    arch: str
    backbone: str
    num_classes: int
    fpn_channels: int
    depth: int
    head_depth: int
    use_p6: bool
    use_p2: bool
    img_size: int
    levels: Tuple[str, ...]
    anchors: Tuple[int, ...]
    names: List[str]


def parse_meta(meta: dict, from_features: bool = False) -> ModelCfg:
    """Same field resolution (and the same KeyError / ValueError behaviour) as tools/infer.py:34-77.  from_features: the
    layer program starts at the FPN (the backbone runs elsewhere), so any backbone name is accepted."""
    cfg = meta.get("config", {}) or {}
    mcfg = cfg.get("model", {}) or {}
    tcfg = cfg.get("training", {}) or {}
    arch = (meta.get("arch") or mcfg.get("arch") or "YOLOLiteMS").lower()
    backbone = meta.get("backbone") or mcfg.get("backbone") or "resnet18"
    nc = int(meta.get("num_classes") or mcfg.get("num_classes") or 80)
    apl = tuple(meta.get("num_anchors_per_level") or (1, 1, 1))
    use_p6 = cfg["training"]["use_p6"]
    use_p2 = cfg["training"]["use_p2"]
    if arch not in ("yololitems", "yololitems_cpu"):
        raise ValueError(f"Okänd arch i meta/config: {arch}")
    if backbone not in BACKBONES and not from_features:
        raise ValueError(f"backbone {backbone!r} has no sm_100a lowering yet (supported: {sorted(BACKBONES)})")
    levels = (("p2",) if use_p2 else ()) + ("p3", "p4", "p5") + (("p6",) if use_p6 else ())
    if len(apl) >= 3:
        a3, a4, a5 = (int(v) for v in apl[:3])
        amap = {"p2": a3, "p3": a3, "p4": a4, "p5": a5, "p6": a5}
    else:
        a = int(apl[0]) if len(apl) else 1
        amap = dict.fromkeys(("p2", "p3", "p4", "p5", "p6"), a)
    names = meta.get("names") or [str(i) for i in range(int(meta.get("num_classes", 80)))]
    return ModelCfg(arch=arch, backbone=backbone, num_classes=nc,
                    fpn_channels=int(int(mcfg.get("fpn_channels", 128)) * float(mcfg.get("width_multiple", 1.0))),
                    depth=max(1, round(2 * float(mcfg.get("depth_multiple", 1.0)))),
                    head_depth=int(mcfg.get("head_depth", 1)), use_p6=bool(use_p6), use_p2=bool(use_p2),
                    img_size=int(tcfg.get("img_size", meta.get("img_size", 640))), levels=levels,
                    anchors=tuple(amap[l] for l in levels), names=list(names))


@dataclass
class _T:
    """A virtual activation tensor."""
    vid: int
    C: int
    red: int            # reduction w.r.t. the network input (bookkeeping only)


@dataclass
class Program:
    ops: List[dict] = field(default_factory=list)
    blob: List[np.ndarray] = field(default_factory=list)
    blob_len: int = 0
    n_virtual: int = 0
    taps: Dict[str, int] = field(default_factory=dict)    # name -> virtual tensor id (debug/parity taps)
    strides: List[int] = field(default_factory=list)
    cfg: Optional[ModelCfg] = None
    n_buffers: int = 0
    vmap: Dict[int, int] = field(default_factory=dict)    # virtual id -> physical buffer id
    feature_channels: Optional[List[int]] = None          # from_features programs: channels of the external inputs

    def add_blob(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)      # float32 views (bf16 pairs) keep their bits
        off = self.blob_len
        pad = (-arr.size) % 64                        # keep every array 256-byte aligned
        self.blob.append(arr)
        if pad:
            self.blob.append(np.zeros(pad, np.float32))
        self.blob_len += arr.size + pad
        return off

    def new(self, C: int, red: int) -> _T:
        t = _T(self.n_virtual, C, red)
        self.n_virtual += 1
        return t


class _SD:
    def __init__(self, sd):
        self.sd = sd
        self.used = set()

    def get(self, key) -> np.ndarray:
        if key not in self.sd:
            raise KeyError(f"checkpoint state_dict has no {key!r}")
        self.used.add(key)
        v = self.sd[key]
        return (v.detach().cpu().double().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, np.float64))

    def bn(self, prefix) -> Tuple[np.ndarray, np.ndarray]:
        g, b = self.get(prefix + ".weight"), self.get(prefix + ".bias")
        m, v = self.get(prefix + ".running_mean"), self.get(prefix + ".running_var")
        self.used.add(prefix + ".num_batches_tracked")
        s = g / np.sqrt(v + BN_EPS)
        return s, b - m * s


def _gemm_w(w: np.ndarray) -> np.ndarray:
    """[Cout,Cin,k,k] -> [k*k*Cin][Cout padded to 4], k index = (ky*k+kx)*Cin+ci."""
    cout = w.shape[0]
    m = np.transpose(w, (2, 3, 1, 0)).reshape(-1, cout)
    ld = (cout + 3) // 4 * 4
    out = np.zeros((m.shape[0], ld), np.float64)
    out[:, :cout] = m
    return out


def bf16_split3(x: np.ndarray):
    """x (fp32) -> three uint16 arrays of bf16 bit patterns with x ~= b1 + b2 + b3 (each round-to-nearest-even), the
    error-compensated operand format of the fused stem kernel (csrc/stem_kernel.cu)."""
    def rn(v):
        u = np.ascontiguousarray(v, np.float32).view(np.uint32).astype(np.uint64)
        r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32)
        return r.astype(np.uint16), (r << 16).astype(np.uint32).view(np.float32)
    x = np.ascontiguousarray(x, np.float32)
    b1, f1 = rn(x)
    r1 = (x - f1).astype(np.float32)
    b2, f2 = rn(r1)
    b3, _ = rn((r1 - f2).astype(np.float32))
    return b1, b2, b3


IMAGENET_MEAN = (0.485, 0.456, 0.406)      # tools/infer.py:432-433
IMAGENET_STD = (0.229, 0.224, 0.225)


def bf16_split3_f64(x: np.ndarray):
    """float64 x -> three bf16 bit patterns with x ~= b1 + b2 + b3 (24 bits kept): like bf16_split3 but the residuals are
    taken in float64, so nothing is lost to an intermediate fp32 rounding."""
    def rn(v):
        f = np.asarray(v, np.float64).astype(np.float32)
        u = f.view(np.uint32).astype(np.uint64)
        r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32)
        return r.astype(np.uint16), (r << 16).astype(np.uint32).view(np.float32).astype(np.float64)
    x = np.asarray(x, np.float64)
    b1, f1 = rn(x)
    b2, f2 = rn(x - f1)
    b3, _ = rn(x - f1 - f2)
    return b1, b2, b3


def _sw64_rows(m: np.ndarray) -> np.ndarray:
    """[rows][32] uint16 -> the K-major SWIZZLE_64B image: row r holds four 16 B chunks, chunk c at position c ^ ((r >> 1) & 3)."""
    rows = m.shape[0]
    t = m.reshape(rows, 4, 8)
    out = np.empty_like(t)
    r = np.arange(rows)
    for c in range(4):
        out[r, c ^ ((r >> 1) & 3), :] = t[r, c, :]
    return out.reshape(rows, 32)


def tc_image(wm: np.ndarray, n_out: int) -> np.ndarray:
    """tcgen05 weight image of a GEMM matrix wm [K][>=n_out] (fp64, BN folded) for csrc/tc_gemm.cu, as float32 words (two bf16
    per word, bits preserved): [ceil(K/32) slabs][3 splits][ceil16(N) rows][32 k].

    w = w1 + w2 + w3 with bf16 splits taken in float64 (24 bits kept).  Each (slab, split) block is the K-major SWIZZLE_64B
    canonical UMMA layout: row n holds 32 consecutive k as four 16 B chunks, chunk c stored at position c ^ ((n >> 1) & 3);
    the three splits of a slab are adjacent so that [W1 | W2] is one B operand of N = 2 * rows."""
    K = wm.shape[0]
    nslab, npad = (K + 31) // 32, (n_out + 15) // 16 * 16
    w = np.zeros((nslab * 32, npad), np.float64)
    w[:K, :n_out] = wm[:, :n_out]
    out = np.empty((nslab, 3, npad, 32), np.uint16)
    for q, sp in enumerate(bf16_split3_f64(w)):
        t = sp.reshape(nslab, 32, npad).transpose(0, 2, 1)               # [slab][n][k]
        for sl in range(nslab):
            out[sl, q] = _sw64_rows(np.ascontiguousarray(t[sl]))
    return out.reshape(-1).view(np.float32)


def tap_padded(wm: np.ndarray, taps: int, cin: int) -> np.ndarray:
    """[taps*cin][N] GEMM matrix -> [taps*ceil32(cin)][N] with every tap's channel block zero-padded to a multiple of 32
    (yl_op.wt_layout = 1): a K-slab of 32 then never straddles two taps."""
    cp = (cin + 31) // 32 * 32
    out = np.zeros((taps * cp, wm.shape[1]), np.float64)
    for t in range(taps):
        out[t * cp:t * cp + cin] = wm[t * cin:(t + 1) * cin]
    return out


def stem_u8_matrix(ws: np.ndarray, b0: np.ndarray) -> np.ndarray:
    """Stem weights for uint8 input as a float64 [32 out][32 k] matrix.  ws: [27][32] (k = (ky*3+kx)*3 + ci, BN folded),
    b0: [32].  Columns 0..26 multiply the raw pixel bytes (channel ci of the RGB tensor), 27 a constant 1, 28 / 29 / 30
    the indicators "stem output row 0" / "column 0" / both (the taps in the zero padding there), 31 is unused."""
    wsd = np.asarray(ws, np.float64).reshape(3, 3, 3, -1)                 # [ky][kx][ci][n]
    n = wsd.shape[-1]
    a = 1.0 / (255.0 * np.asarray(IMAGENET_STD, np.float64))
    b = -np.asarray(IMAGENET_MEAN, np.float64) / np.asarray(IMAGENET_STD, np.float64)
    su = np.zeros((n, 32), np.float64)
    su[:, :27] = (wsd * a[None, None, :, None]).reshape(27, n).T
    wb = wsd * b[None, None, :, None]
    su[:, 27] = np.asarray(b0, np.float64) + wb.sum(axis=(0, 1, 2))
    su[:, 28] = -wb[0].sum(axis=(0, 1))
    su[:, 29] = -wb[:, 0].sum(axis=(0, 1))
    su[:, 30] = wb[0, 0].sum(axis=0)
    return su


def stem2_image(w2m: np.ndarray, n_out: int, ws: np.ndarray, b0: np.ndarray) -> np.ndarray:
    """Weight image of the fused stem kernel as float32 words (two bf16 per word, bits preserved):
    [9 taps][3 splits][ceil16(n_out)][32 ch] conv2 weights (the three splits of a tap are adjacent row blocks, so one
    MMA can take [W1|W2|W3], [W1|W2] or [W1] as its B operand), then [3 splits][32 stem ch][32 k] stem weights where
    k = (ky*3 + kx)*3 + ci for k < 27, k = 27 is the folded BN bias (multiplied by a constant-1 column), k > 27 zero.
    w2m: [288][>= n_out] with k = (ky*3 + kx)*32 + ci; ws: [27][32]; b0: [32]."""
    n2 = (n_out + 15) // 16 * 16
    w2 = np.zeros((9, n2, 32), np.float32)
    w2[:, :n_out, :] = np.asarray(w2m, np.float64)[:, :n_out].reshape(9, 32, n_out).transpose(0, 2, 1).astype(np.float32)
    st = np.zeros((32, 32), np.float32)
    st[:, :27] = np.asarray(ws, np.float64).T.astype(np.float32)
    st[:, 27] = np.asarray(b0, np.float64).astype(np.float32)
    parts = []
    sp2 = bf16_split3(w2)
    for t in range(9):
        for sp in sp2:
            parts.append(_sw64_rows(sp[t]).reshape(-1))
    for sp in bf16_split3(st):
        parts.append(_sw64_rows(sp).reshape(-1))
    # second stem image, for uint8 BGR input (yl_forward_u8): x = a_c*u + b_c with a_c = 1/(255 std_c), b_c = -mean_c/std_c
    # (tools/infer.py:432-433,449-451) is affine in the integer u, so  conv(x) = sum_k (w_k a_c) u_k + sum_{valid k} w_k b_c.
    # Row 27 (times 1) carries bias + the full b-term; rows 28 / 29 / 30 (times the top-row / left-column / corner
    # indicators) take back the taps that fall into the zero padding there.  Computed in float64, then split.
    su = stem_u8_matrix(ws, b0)
    for sp in bf16_split3_f64(su):
        parts.append(_sw64_rows(sp).reshape(-1))
    return np.concatenate(parts).astype(np.uint16).view(np.float32)


def _pad4(b: np.ndarray) -> np.ndarray:
    out = np.zeros(((b.size + 3) // 4 * 4,), np.float64)
    out[:b.size] = b
    return out


def lower(state_dict: dict, meta: dict, fuse_dwpw: bool = True, reuse_buffers: bool = True,
          tensor_cores: bool = True, fuse_stem: bool = True, fuse_uir: Optional[bool] = None, fuse_pw01: bool = True,
          from_features: bool = False) -> Program:
    """from_features: lower only the FPN + heads (model_v2.py:124-133,201-224 / :289-294,359-377); the program reads the backbone's
    feature maps [c2,] c3, c4, c5 as external NHWC inputs (op.src = YL_SRC_FEATURE(i)), their channel counts are taken from
    the lateral convs' weights, their reductions are assumed to be [4,] 8, 16, 32 (what every timm backbone the reference
    configures returns).  fuse_dwpw: DWConvBlock (FPN smooth / head trunk) as one op; fuse_uir (default = fuse_dwpw): the stride-1
    depthwise convs of the backbone's UIR blocks ride in the producer stage of the pointwise conv that follows them
    (dw_start -> pw_exp, dw_mid -> pw_proj + residual), so their outputs never reach HBM."""
    if fuse_uir is None:
        fuse_uir = fuse_dwpw
    cfg = parse_meta(meta, from_features)
    sd = _SD(state_dict)
    P = Program(cfg=cfg)

    def emit(kind, src: Optional[_T], cout, red, k=1, stride=1, act=L.ACT_NONE, w=None, b=None, res: Optional[_T] = None,
             up: Optional[_T] = None, anchors=0, level=None, w2=None, k2=0, w3=None, b2=None, act2=L.ACT_NONE, stride2=0) -> Optional[_T]:
        dst = None if level is not None else P.new(cout, red)
        cin_ = 3 if src is None else src.C
        # dense k x k, stride 1, long cin (the 3x3 convs of the YOLOLiteMS FPN): per-tap padded K axis -> every K-slab is one TMA box
        tap_layout = bool(tensor_cores and kind == L.OP_CONV and k > 1 and stride == 1 and cin_ > 64 and cin_ % 4 == 0)
        wmat = np.asarray(w, np.float64).reshape(-1, w.shape[-1])
        if tap_layout:
            wmat = tap_padded(wmat, k * k, cin_)
        P.ops.append(dict(kind=kind, src=(-1 if src is None else src.vid), dst=(-(1 + level) if level is not None else dst.vid),
                          res=(-1 if res is None else res.vid), up=(-1 if up is None else up.vid),
                          cin=(3 if src is None else src.C), cout=cout, k=k, stride=stride, act=act, anchors=anchors, k2=k2,
                          w_off=P.add_blob(w), b_off=(-1 if b is None else P.add_blob(_pad4(b))),
                          w2_off=(-1 if w2 is None else P.add_blob(w2)), w3_off=(-1 if w3 is None else P.add_blob(w3)),
                          b2_off=(-1 if b2 is None else P.add_blob(_pad4(b2))), act2=act2, stride2=stride2,
                          wt_layout=(1 if tap_layout else 0),
                          wt_off=(P.add_blob(tc_image(wmat, cout))
                                  if tensor_cores and kind in (L.OP_CONV, L.OP_DWPW, L.OP_STEM2) and cout >= 8 and w.shape[0] >= 8 else -1)))
        return dst

    def conv_bn(x: Optional[_T], wkey, bnkey, k, stride, act, red, res=None) -> _T:
        w = sd.get(wkey + ".weight")
        s, b = sd.bn(bnkey)
        w = w * s[:, None, None, None]
        cout = w.shape[0]
        if x is None:      # stem on the NCHW input: [k*k*3][Cout]
            return emit(L.OP_STEM, None, cout, red, k=k, stride=stride, act=act,
                        w=np.transpose(w, (2, 3, 1, 0)).reshape(-1, cout), b=b)
        return emit(L.OP_CONV, x, cout, red, k=k, stride=stride, act=act, w=_gemm_w(w), b=b, res=res)

    def dw_bn(x: _T, wkey, bnkey, k, stride, act, red) -> _T:
        w = sd.get(wkey + ".weight")                       # [C,1,k,k]
        s, b = sd.bn(bnkey)
        w = w * s[:, None, None, None]
        return emit(L.OP_DW, x, x.C, red, k=k, stride=stride, act=act, w=np.transpose(w, (2, 3, 1, 0)).reshape(k * k, -1), b=b)

    def dw_pw_bn(x: _T, dwkey, dwbn, k, act2, pwkey, pwbn, act, red, res=None, stride=1) -> _T:
        """depthwise k x k (stride 1 or 2) + BN (+act2) -> pointwise + BN (+res) (+act) as ONE op (YL_OP_DWPW)."""
        wd = sd.get(dwkey + ".weight")
        sdw, bdw = sd.bn(dwbn)
        wd = wd * sdw[:, None, None, None]
        wp = sd.get(pwkey + ".weight")
        sp, bp = sd.bn(pwbn)
        wp = wp * sp[:, None, None, None]
        return emit(L.OP_DWPW, x, wp.shape[0], red, k=1, act=act, w=_gemm_w(wp), b=bp, res=res,
                    w2=np.transpose(wd, (2, 3, 1, 0)).reshape(k * k, -1), k2=k, b2=bdw, act2=act2, stride2=stride)

    take = 4 if cfg.use_p2 else 3
    if from_features:
        lat = (["lateral2"] if cfg.use_p2 else []) + ["lateral3", "lateral4", "lateral5"]
        reds = [4, 8, 16, 32][-take:]
        feats = [_T(L.src_feature(i), int(sd.get(nm + ".weight").shape[1]), r) for i, (nm, r) in enumerate(zip(lat, reds))]
        P.feature_channels = [f.C for f in feats]
    else:
        # ---------------- backbone
        table, mult, stem_c = BACKBONES[cfg.backbone]
        bb = "backbone."
        first = table[0][0]
        # the stem feature itself is never tapped (the FPN takes the last 3-4 taps), so conv_stem can be fused with
        # blocks.0.0 when both are 3x3 s2 and the stem has 32 channels
        fused_stem = bool(fuse_stem and tensor_cores and stem_c == 32 and first[0] == "cn" and first[1] == 3 and first[2] == 2)
        pw_blob = None
        if fused_stem:
            ws = sd.get(bb + "conv_stem.weight")
            s0, b0 = sd.bn(bb + "bn1")
            ws = np.transpose(ws * s0[:, None, None, None], (2, 3, 1, 0)).reshape(27, stem_c)
            key = bb + "blocks.0.0"
            w1 = sd.get(key + ".conv.weight")
            s1, b1 = sd.bn(key + ".bn1")
            w1m = _gemm_w(w1 * s1[:, None, None, None])
            # blocks.0.1 (1x1, same width, BN + ReLU) rides in the output epilogue of the fused kernel when conv2 has 16 channels
            nxt01 = table[0][1] if len(table[0]) > 1 else None
            c1 = int(w1.shape[0])
            pw_blob = None
            if fuse_pw01 and c1 == 16 and nxt01 is not None and nxt01[0] == "cn" and nxt01[1] == 1 and nxt01[2] == 1 and _round_ch(nxt01[3] * mult) == 16:
                wq = sd.get(bb + "blocks.0.1.conv.weight")
                sq, bq = sd.bn(bb + "blocks.0.1.bn1")
                wq = (wq * sq[:, None, None, None])[:, :, 0, 0]                       # [n][k]
                pw_blob = np.concatenate([wq.T.reshape(-1), bq.reshape(-1)])          # [k][n] then bias[n]
            x = emit(L.OP_STEM2, None, c1, 4, k=3, stride=2, act=L.ACT_RELU, w=w1m, b2=pw_blob, act2=(L.ACT_RELU if pw_blob is not None else L.ACT_NONE),
                     w3=stem2_image(w1m, int(w1.shape[0]), ws, b0) if int(w1.shape[0]) <= 32 else None,
                     b=b1, w2=np.concatenate([ws.reshape(-1), b0.reshape(-1)]), k2=stem_c)
            feats = [_T(-1, stem_c, 2)]
            red = 4
        else:
            x = conv_bn(None, bb + "conv_stem", bb + "bn1", 3, 2, L.ACT_RELU, 2)
            feats = [x]
            red = 2
        for si, stage in enumerate(table):
            for bi, spec in enumerate(stage):
                key = f"{bb}blocks.{si}.{bi}"
                if fused_stem and si == 0 and (bi == 0 or (bi == 1 and pw_blob is not None)):
                    pass
                elif spec[0] == "cn":
                    _, k, s, c = spec
                    red *= s
                    x = conv_bn(x, key + ".conv", key + ".bn1", k, s, L.ACT_RELU, red)
                else:
                    _, ks, km, s, e, c = spec
                    cout = _round_ch(c * mult)
                    skip = x if (x.C == cout and s == 1) else None
                    y = x
                    s_start = 1 if km else s
                    fuse_start = bool(ks and fuse_uir and s_start == 1 and x.C % 4 == 0)
                    fuse_mid = bool(km and fuse_uir and s in (1, 2))
                    if fuse_start:
                        y = dw_pw_bn(y, key + ".dw_start.conv", key + ".dw_start.bn", ks, L.ACT_NONE,
                                     key + ".pw_exp.conv", key + ".pw_exp.bn", L.ACT_RELU, red)
                    else:
                        if ks:
                            y = dw_bn(y, key + ".dw_start.conv", key + ".dw_start.bn", ks, s_start, L.ACT_NONE, red * s_start)
                        y = conv_bn(y, key + ".pw_exp.conv", key + ".pw_exp.bn", 1, 1, L.ACT_RELU, y.red)
                    if fuse_mid:
                        red *= s
                        x = dw_pw_bn(y, key + ".dw_mid.conv", key + ".dw_mid.bn", km, L.ACT_RELU,
                                     key + ".pw_proj.conv", key + ".pw_proj.bn", L.ACT_NONE, red, res=skip, stride=s)
                    else:
                        if km:
                            y = dw_bn(y, key + ".dw_mid.conv", key + ".dw_mid.bn", km, s, L.ACT_RELU, red * s)
                        red *= s
                        x = conv_bn(y, key + ".pw_proj.conv", key + ".pw_proj.bn", 1, 1, L.ACT_NONE, red, res=skip)
                last_of_stage = bi == len(stage) - 1
                nxt = table[si + 1][0] if si + 1 < len(table) else None
                nxt_stride = None if nxt is None else (nxt[2] if nxt[0] == "cn" else nxt[3])
                if last_of_stage and (nxt is None or nxt_stride > 1):
                    feats.append(x)
    feats = feats[-take:]
    P.strides = [f.red for f in feats] + ([feats[-1].red * 2] if cfg.use_p6 else [])

    # ---------------- FPN
    Fc, d, C = cfg.fpn_channels, cfg.depth, cfg.num_classes
    cpu = cfg.arch == "yololitems_cpu"

    def dw_block(x: _T, name: str, n: int) -> _T:           # model_v2.py:23-39
        for i in range(n):
            wd = sd.get(f"{name}.block.{4*i}.weight")        # [C,1,3,3], no BN, no act
            wp = sd.get(f"{name}.block.{4*i+1}.weight")
            s, b = sd.bn(f"{name}.block.{4*i+2}")
            wp = wp * s[:, None, None, None]
            wd9 = np.transpose(wd, (2, 3, 1, 0)).reshape(9, -1)
            if fuse_dwpw and x.C <= 384:
                x = emit(L.OP_DWPW, x, wp.shape[0], x.red, k=1, act=L.ACT_RELU, w=_gemm_w(wp), b=b, w2=wd9, k2=3)
            else:
                t = emit(L.OP_DW, x, x.C, x.red, k=3, act=L.ACT_NONE, w=wd9)
                x = emit(L.OP_CONV, t, wp.shape[0], x.red, k=1, act=L.ACT_RELU, w=_gemm_w(wp), b=b)
        return x

    def dense_block(x: _T, name: str, n: int) -> _T:        # model_v2.py:15-22
        for i in range(n):
            x = conv_bn(x, f"{name}.{3*i}", f"{name}.{3*i+1}", 3, 1, L.ACT_SILU, x.red)
        return x

    smooth = dw_block if cpu else dense_block

    def lateral(c: _T, name: str, up: Optional[_T]) -> _T:
        w = sd.get(name + ".weight")
        return emit(L.OP_CONV, c, w.shape[0], c.red, k=1, w=_gemm_w(w), b=sd.get(name + ".bias"), up=up)

    c5, c4, c3 = feats[-1], feats[-2], feats[-3]
    pyr: Dict[str, _T] = {}
    pyr["p5"] = smooth(lateral(c5, "lateral5", None), "smooth5", d)
    pyr["p4"] = smooth(lateral(c4, "lateral4", pyr["p5"]), "smooth4", d)
    pyr["p3"] = smooth(lateral(c3, "lateral3", pyr["p4"]), "smooth3", d)
    if cfg.use_p2:
        pyr["p2"] = smooth(lateral(feats[0], "lateral2", pyr["p3"]), "smooth2", d)
    if cfg.use_p6:
        t = conv_bn(pyr["p5"], "p6_down", "p6_bn", 3, 2, L.ACT_RELU if cpu else L.ACT_SILU, pyr["p5"].red * 2)
        pyr["p6"] = smooth(t, "smooth6", d)
    P.taps.update({"c3": c3.vid, "c4": c4.vid, "c5": c5.vid, **{k: v.vid for k, v in pyr.items()}})

    # ---------------- heads (model_v2.py:42-53, :340-350): one GEMM with N = A*(5+C), rows ordered a*(5+C)+d
    for li, (lvl, A) in enumerate(zip(cfg.levels, cfg.anchors)):
        name = "head" + lvl[1]
        p = pyr[lvl]
        for i in range(cfg.head_depth):
            p = dw_block(p, f"{name}.trunk.{i}", 1)
        wb, bb_ = sd.get(f"{name}.out.box.weight")[:, :, 0, 0], sd.get(f"{name}.out.box.bias")
        wo, bo = sd.get(f"{name}.out.obj.weight")[:, :, 0, 0], sd.get(f"{name}.out.obj.bias")
        wc, bc = sd.get(f"{name}.out.cls.weight")[:, :, 0, 0], sd.get(f"{name}.out.cls.bias")
        D = 5 + C
        W = np.zeros((A * D, Fc), np.float64)
        Bv = np.zeros((A * D,), np.float64)
        for a in range(A):
            W[a * D:a * D + 4] = wb[a * 4:a * 4 + 4]; Bv[a * D:a * D + 4] = bb_[a * 4:a * 4 + 4]
            W[a * D + 4] = wo[a]; Bv[a * D + 4] = bo[a]
            W[a * D + 5:(a + 1) * D] = wc[a * C:(a + 1) * C]; Bv[a * D + 5:(a + 1) * D] = bc[a * C:(a + 1) * C]
        emit(L.OP_CONV, p, A * D, p.red, k=1, w=_gemm_w(W[:, :, None, None]), b=Bv, anchors=A, level=li)

    _assign_buffers(P, reuse_buffers)
    return P


def _assign_buffers(P: Program, reuse: bool) -> None:
    """Map virtual tensors to physical buffers; with reuse, a buffer is recycled after its last reader."""
    last_use: Dict[int, int] = {}
    for i, op in enumerate(P.ops):
        for f in ("src", "res", "up"):
            if op[f] >= 0:
                last_use[op[f]] = i
    if not reuse:
        P.vmap = {v: v for v in range(P.n_virtual)}
        P.n_buffers = P.n_virtual
    else:
        free: List[int] = []
        vmap: Dict[int, int] = {}
        n = 0
        for i, op in enumerate(P.ops):
            if op["dst"] >= 0:
                if free:
                    pid = free.pop()
                else:
                    pid = n
                    n += 1
                vmap[op["dst"]] = pid
            # release inputs whose last reader is this op (after allocating dst: never alias in/out)
            for f in ("src", "res", "up"):
                v = op[f]
                if v >= 0 and last_use.get(v) == i and vmap[v] not in free:
                    free.append(vmap[v])
            if op["dst"] >= 0 and op["dst"] not in last_use:      # dead tensor
                free.append(vmap[op["dst"]])
        P.vmap, P.n_buffers = vmap, max(n, 1)
    for op in P.ops:
        for f in ("src", "res", "up", "dst"):
            if op[f] >= 0:
                op[f] = P.vmap[op[f]]
    P.taps = {k: P.vmap[v] for k, v in P.taps.items() if v >= 0}


def to_c(P: Program):
    """(ctypes op array, contiguous fp32 blob)."""
    arr = (L.YlOp * len(P.ops))()
    for i, op in enumerate(P.ops):
        o = arr[i]
        for f in ("kind", "src", "dst", "res", "up", "cin", "cout", "k", "stride", "act", "anchors", "k2", "w_off", "b_off",
                  "w2_off", "wt_off", "w3_off", "b2_off", "act2", "stride2", "wt_layout"):
            setattr(o, f, int(op[f]))
    blob = np.concatenate(P.blob).astype(np.float32, copy=False)
    assert blob.size == P.blob_len
    return arr, blob
