"""ctypes binding of libyololite_b200.so (include/yololite_b200.h).  No fallback: if the shared library is
missing the import fails loudly and tells the user how to build it."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libyololite_b200.so")


class YlOp(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("src", ctypes.c_int32), ("dst", ctypes.c_int32), ("res", ctypes.c_int32),
                ("up", ctypes.c_int32), ("cin", ctypes.c_int32), ("cout", ctypes.c_int32), ("k", ctypes.c_int32),
                ("stride", ctypes.c_int32), ("act", ctypes.c_int32), ("anchors", ctypes.c_int32),
                ("k2", ctypes.c_int32), ("w_off", ctypes.c_int64), ("b_off", ctypes.c_int64),
                ("w2_off", ctypes.c_int64), ("wt_off", ctypes.c_int64), ("w3_off", ctypes.c_int64),
                ("b2_off", ctypes.c_int64), ("act2", ctypes.c_int32), ("stride2", ctypes.c_int32),
                ("wt_layout", ctypes.c_int32), ("reserved0", ctypes.c_int32)]

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        if not args:                      # optional blob offsets default to "absent", not to offset 0
            for f in ("b_off", "w2_off", "wt_off", "w3_off", "b2_off"):
                if f not in kw:
                    setattr(self, f, -1)


OP_STEM, OP_CONV, OP_DW, OP_DWPW, OP_STEM2 = 0, 1, 2, 3, 4
ACT_NONE, ACT_RELU, ACT_SILU = 0, 1, 2
SRC_INPUT = -1
ABI_VERSION = 4


def src_feature(i: int) -> int:
    """op.src of externally supplied backbone feature i (YL_SRC_FEATURE in the header)."""
    return -(2 + i)

_lib = None

_P = ctypes.c_void_p
_SIGNATURES = {
    "yl_engine_create": (ctypes.c_int, [ctypes.POINTER(YlOp), ctypes.c_int32, _P, ctypes.c_size_t, ctypes.c_int32,
                                        ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(_P)]),
    "yl_engine_destroy": (ctypes.c_int, [_P]),
    "yl_engine_set_option": (ctypes.c_int, [_P, ctypes.c_char_p, ctypes.c_int32]),
    "yl_run_op": (ctypes.c_int, [ctypes.POINTER(YlOp), _P, _P, _P, _P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                 ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P]),
    "yl_engine_plan": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]),
    "yl_forward": (ctypes.c_int, [_P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(_P), _P]),
    "yl_forward_u8": (ctypes.c_int, [_P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(_P), _P]),
    "yl_forward_profile": (ctypes.c_int, [_P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(_P), _P,
                                          ctypes.POINTER(ctypes.c_float), ctypes.c_int32]),
    "yl_engine_read_buffer": (ctypes.c_int, [_P, ctypes.c_int32, _P, ctypes.POINTER(ctypes.c_int32), _P]),
    "yl_postprocess_scratch_bytes": (ctypes.c_size_t, [ctypes.c_int32, ctypes.c_int64]),
    "yl_postprocess": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int32), ctypes.c_int32, ctypes.c_int32,
                                      ctypes.c_int32, ctypes.c_int32, ctypes.c_float, ctypes.c_double, ctypes.c_int32,
                                      ctypes.c_int32, _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _P]),
    "yl_postprocess_ex": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int32), ctypes.c_int32, ctypes.c_int32,
                                         ctypes.c_int32, ctypes.c_int32, ctypes.c_float, ctypes.c_double, ctypes.c_int32,
                                         ctypes.c_int32, _P, _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _P]),
    "yl_engine_plan_features": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.c_int32,
                                               ctypes.POINTER(ctypes.c_int32)]),
    "yl_forward_features": (ctypes.c_int, [_P, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int32), ctypes.c_int32, ctypes.c_int32,
                                           ctypes.POINTER(_P), _P]),
    "yl_engine_detect": (ctypes.c_int, [_P, _P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_float,
                                        ctypes.c_double, ctypes.c_int32, ctypes.c_int32, _P, _P, _P, _P, _P, _P, _P]),
    "yl_engine_levels": (ctypes.c_int, [_P, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int32)]),
    "yl_decode": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int32), ctypes.c_int32, ctypes.c_int32,
                                 ctypes.c_int32, ctypes.c_int32, _P, _P, _P, _P]),
    "yl_preprocess": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P, ctypes.c_int32,
                                     ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P]),
    "yl_preprocess_batch": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P]),
    "yl_stat": (ctypes.c_longlong, [ctypes.c_char_p]),
    "yl_last_error": (ctypes.c_char_p, []),
    "yl_abi_version": (ctypes.c_int, []),
    "yl_device_count": (ctypes.c_int, []),
}
EXPORTS = tuple(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension is not built and there is no fallback path. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` or "
                "`python yololite-official-repo_b200/build.py` (needs nvcc, targets sm_100a).")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().yl_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise RuntimeError(msg)
