"""The C-ABI library loads and exports every symbol include/yololite_b200.h declares; no compute without a GPU."""
import ctypes
import os
import re

import pytest
import torch

from conftest import REPO


def _declared():
    with open(os.path.join(REPO, "include", "yololite_b200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(yl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import yololite_b200 as y
    lib = y.lib()
    decl = _declared()
    assert len(decl) >= 12
    for name in decl:
        assert hasattr(lib, name), name
    assert sorted(y.EXPORTS) == decl
    assert lib.yl_abi_version() == 4


def test_op_struct_layout_matches_header():
    from yololite_b200 import _lib as L
    assert ctypes.sizeof(L.YlOp) == 12 * 4 + 6 * 8 + 4 * 4
    assert L.YlOp.wt_layout.offset == 104
    assert L.YlOp.b2_off.offset == 88 and L.YlOp.act2.offset == 96
    assert L.YlOp.w_off.offset == 48


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    import yololite_b200 as y
    from conftest import synth_ckpt
    from yololite_b200 import _lib as L, packer
    ck = synth_ckpt("edge_n", 3, 64)
    with pytest.raises(RuntimeError):
        y.YoloLiteB200(ck["state_dict"], ck["meta"])
    P = packer.lower(ck["state_dict"], ck["meta"])
    ops, blob = packer.to_c(P)
    h = ctypes.c_void_p()
    rc = y.lib().yl_engine_create(ops, len(ops), blob.ctypes.data_as(ctypes.c_void_p), blob.size, P.n_buffers, 3, 0, ctypes.byref(h))
    assert rc != 0 and len(y.lib().yl_last_error()) > 0
    with pytest.raises((RuntimeError, ValueError)):
        L.check(rc)


def test_missing_library_is_an_import_error(monkeypatch):
    from yololite_b200 import _lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", os.path.join(REPO, "does_not_exist.so"))
    with pytest.raises(ImportError):
        L.lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "yololite-official-repo_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(root, f)) as fh:
                    assert not re.search(r"^\s*(from|import)\s+oracle", fh.read(), flags=re.M), f


def test_fused_stem_row_order_enumerates_the_halo_plane_by_plane():
    """GEMM1 rows of the fused stem kernel are in PARITY-PLANE order (csrc/stem_kernel.cu: s2_row_pixel): row q is pixel q of the
    four planes (row parity, column parity) of the 33 x 17 stem-output halo, pitches 9 / 8, offsets 0 / 153 / 289 / 433 -- the
    layout conv2's shifted SWIZZLE_64B descriptors address.  Host-side view through yl_stat; no GPU needed."""
    import yololite_b200 as y
    lib = y.lib()
    seen = {}
    for q in range(640):
        v = lib.yl_stat(b"stem_row_pixel%d" % q)
        if q >= 561:
            assert v == 0xFFFF, (q, v)
            continue
        hy, hx = v & 0xFF, v >> 8
        assert 0 <= hy < 33 and 0 <= hx < 17
        off = {(0, 0): 0, (0, 1): 153, (1, 0): 289, (1, 1): 433}[(hy & 1, hx & 1)]
        assert q == off + (hy >> 1) * (8 if hx & 1 else 9) + (hx >> 1), (q, hy, hx)
        seen[(hy, hx)] = q
    assert len(seen) == 33 * 17
    assert lib.yl_stat(b"stem_row_pixel640") == -1
