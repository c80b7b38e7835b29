"""Engine-level behaviour on the GPU: the one-call detect entry, CUDA graph replay, programmatic dependent launch, the packed
multi-GPU payload, launch-record caching, device handling and the unfused stem fallback -- every variant must give the same
bits as the plain yl_forward + yl_postprocess path."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import synth_ckpt
from oracle import model_ref, post_ref

pytestmark = pytest.mark.gpu


def _engine(ck, dev="cuda:0", **kw):
    import yololite_b200 as y
    return y.YoloLiteB200(ck["state_dict"], ck["meta"], device=dev, **kw)


def _stat(key):
    import yololite_b200 as y
    return y.lib().yl_stat(key.encode())


def test_detect_one_call_equals_forward_plus_postprocess_and_graph_replay():
    import yololite_b200 as y
    ck = synth_ckpt("edge_n", 80, 320, obj_shift=2.0)
    x = model_ref.synth_input(3, 320, seed=4).cuda()
    ref_eng = _engine(ck, pdl=False)
    levels = ref_eng(x)
    want = y.PostProcessor()(levels, 320, 0.25, 0.5, 300, cap=512)
    wl = want.to_list()
    assert sum(len(d["index"]) for d in wl) > 10
    for graph in (False, True):
        eng = _engine(ck, graph=graph)
        cap0 = _stat("graph_captures")
        bufs = (torch.empty((3, 512, 4), device="cuda"), torch.empty((3, 512), device="cuda"),
                torch.empty((3, 512), device="cuda", dtype=torch.int64), torch.empty((3, 512), device="cuda", dtype=torch.int64),
                torch.zeros((3,), device="cuda", dtype=torch.int32))
        for rep in range(4):           # graph mode: 1st call eager, 2nd captures, later ones replay (same pointers every call)
            bx, sc, cl, ix, cn = eng.detect(x, 320, 0.25, 0.5, 300, cap=512, outputs=bufs)
            got = y.Detections(bx, sc, cl, ix, cn).to_list()
            for g, w in zip(got, wl):
                assert torch.equal(g["index"], w["index"]) and torch.equal(g["boxes"], w["boxes"])
                assert torch.equal(g["scores"], w["scores"]) and torch.equal(g["classes"], w["classes"])
        for a, b in zip(eng.last_levels(), levels):
            assert torch.equal(a, b)
        assert (_stat("graph_captures") - cap0) == (1 if graph else 0)
        eng.close()


def test_graph_forward_matches_eager_and_recaptures_on_new_pointers():
    ck = synth_ckpt("edge_n", 3, 64)
    x1 = model_ref.synth_input(2, 64, seed=1).cuda()
    x2 = model_ref.synth_input(2, 64, seed=2).cuda()
    eager = _engine(ck)
    w1, w2 = eager(x1), eager(x2)
    eng = _engine(ck, graph=True)
    out = [torch.empty_like(o) for o in w1]
    g0, c0 = _stat("graph_launches"), _stat("graph_captures")
    for rep in range(3):
        eng.forward(x1, out=out)
        for a, b in zip(out, w1):
            assert torch.equal(a, b)
    assert _stat("graph_captures") - c0 == 1 and _stat("graph_launches") - g0 == 2
    for rep in range(2):               # same outputs, another input pointer: its own graph
        eng.forward(x2, out=out)
        for a, b in zip(out, w2):
            assert torch.equal(a, b)
    assert _stat("graph_captures") - c0 == 2
    # a side stream works too (the capture runs on an engine-owned stream)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.forward(x1, out=out)
    s.synchronize()
    for a, b in zip(out, w1):
        assert torch.equal(a, b)


def test_pdl_on_off_same_bits_and_launch_records_are_cached():
    ck = synth_ckpt("edge_n", 80, 320)
    x = model_ref.synth_input(2, 320, seed=6).cuda()
    a = _engine(ck, pdl=True)
    b = _engine(ck, pdl=False)
    oa = [torch.empty(s, device="cuda") for s in [(2, 1, 40, 40, 85), (2, 1, 20, 20, 85), (2, 1, 10, 10, 85)]]
    a.forward(x, out=oa)
    p0 = _stat("prepares")
    for _ in range(5):
        a.forward(x, out=oa)           # same pointers: no launch record is rebuilt
    assert _stat("prepares") == p0
    ob = b(x)
    for u, v in zip(oa, ob):
        assert torch.equal(u, v)
    want = model_ref.forward_ref(ck["state_dict"], ck["meta"], x.cpu())
    for u, w in zip(oa, want):
        assert float((u.cpu() - w).abs().max()) <= 1e-3


def test_packed_payload_written_by_the_kernel():
    import yololite_b200 as y
    rng = np.random.RandomState(3)
    lv = [torch.from_numpy((rng.randn(3, 1, s, s, 9) * 2).astype(np.float32)).cuda() for s in (16, 8, 4)]
    for l in lv:
        l[..., 4] += 1.0
    cap = 336                                            # = N: cannot overflow
    pp = y.PostProcessor()
    packed = torch.full((3, cap + 1, 6), -7.0, device="cuda")
    d = pp(lv, 128, 0.25, 0.5, 300, cap=cap, packed=packed)
    ref = d.to_list()
    got = y.post.unpack(packed)
    assert sum(len(r["index"]) for r in ref) > 5
    for g, r in zip(got, ref):
        assert torch.equal(g["boxes"], r["boxes"]) and torch.equal(g["scores"], r["scores"]) and torch.equal(g["classes"], r["classes"])
    # overflow is flagged in the header row
    small = torch.zeros((3, 3, 6), device="cuda")
    pp2 = y.PostProcessor()
    pp2(lv, 128, 0.25, 0.5, 300, cap=2, packed=small)
    assert float(small[:, 0, 1].max()) == 1.0 and float(small[:, 0, 0].max()) == 2.0
    with pytest.raises(RuntimeError):
        y.post.unpack(small)
    # packed-only call through the C ABI (all other outputs NULL)
    eng_ck = synth_ckpt("edge_n", 4, 64, obj_shift=3.0)
    eng = _engine(eng_ck)
    x = model_ref.synth_input(2, 64, seed=1).cuda()
    pk = torch.zeros((2, 33, 6), device="cuda")
    eng.detect(x, 64, 0.1, 0.5, 300, cap=32, packed=pk)
    bx, sc, cl, ix, cn = eng.detect(x, 64, 0.1, 0.5, 300, cap=32)
    for b, g in enumerate(y.post.unpack(pk)):
        c = int(cn[b])
        assert torch.equal(g["boxes"], bx[b, :c]) and torch.equal(g["classes"], cl[b, :c])


def test_detections_are_copies_and_out_buffers_are_validated():
    import yololite_b200 as y
    ck = synth_ckpt("edge_n", 4, 64, obj_shift=3.0)
    eng = _engine(ck)
    x1 = model_ref.synth_input(1, 64, seed=1).cuda()
    x2 = model_ref.synth_input(1, 64, seed=2).cuda()
    r1 = y.detect(eng(x1), 64, 0.1, 0.5, 300)
    keep = [{k: v.clone() for k, v in d.items()} for d in r1]
    y.detect(eng(x2), 64, 0.1, 0.5, 300)           # the reference returns fresh tensors: an earlier result must survive
    for a, b in zip(r1, keep):
        assert all(torch.equal(a[k], b[k]) for k in a)
    with pytest.raises(ValueError):
        eng.forward(x1, out=[torch.empty(1, device="cuda")] * 3)
    with pytest.raises(ValueError):
        eng.forward(x1, out=[torch.empty((1, 1, 8, 8, 9)), torch.empty((1, 1, 4, 4, 9)), torch.empty((1, 1, 2, 2, 9))])
    with pytest.raises(RuntimeError):
        eng.to("cpu")
    assert eng.to("cuda") is eng


def test_current_device_is_preserved_and_two_engines_on_two_devices():
    import yololite_b200 as y
    ck = synth_ckpt("edge_n", 3, 64)
    x = model_ref.synth_input(1, 64, seed=3)
    want = model_ref.forward_ref(ck["state_dict"], ck["meta"], x)
    n = torch.cuda.device_count()
    e0 = _engine(ck, "cuda:0")
    assert torch.cuda.current_device() == 0
    o0 = e0(x.cuda(0))
    if n < 2:
        # single GPU: still check that a second engine in the same process / thread works (per-device function attributes)
        e1 = _engine(ck, "cuda:0")
        o1 = e1(x.cuda(0))
    else:
        e1 = _engine(ck, "cuda:1")
        assert torch.cuda.current_device() == 0           # creating / running an engine elsewhere does not move the caller
        o1 = e1(x.cuda(1))
        assert torch.cuda.current_device() == 0
        d1 = y.detect(o1, 64, 0.001, 0.65, 0)              # postprocess on cuda:1 buffers while cuda:0 is current
        d0 = y.detect(o0, 64, 0.001, 0.65, 0)
        assert torch.equal(d0[0]["index"].cpu(), d1[0]["index"].cpu())
    for a, b, w in zip(o0, o1, want):
        assert torch.equal(a.cpu(), b.cpu())
        assert float((a.cpu() - w).abs().max()) <= 1e-3


@pytest.mark.parametrize("H,W", [(350, 350), (90, 102)])
def test_unfused_stem_fallback_for_widths_the_fused_kernel_cannot_take(H, W):
    """edge_n at sizes with W % 4 != 0 (the reference accepts any --img_size): the fused stem kernel's TMA needs 16-byte rows, so
    the engine runs conv_stem -> blocks.0.0 on the tf32 kernel and blocks.0.1 as its own launch."""
    ck = synth_ckpt("edge_n", 3, 64)
    eng = _engine(ck)
    x = model_ref.synth_input(1, 352, seed=2)[:, :, :H, :W].contiguous()
    want = model_ref.forward_ref(ck["state_dict"], ck["meta"], x)
    outs = eng(x.cuda())
    for o, w in zip(outs, want):
        assert o.shape == w.shape
        assert float((o.cpu() - w).abs().max()) <= 1e-3


def test_predict_batch_uses_one_graph_launch_per_call(tmp_path):
    import yololite_b200 as y
    ck = synth_ckpt("edge_n", 5, 64, obj_shift=3.0)
    p = tmp_path / "m.pt"
    torch.save(ck, str(p))
    m = y.YoloLite(str(p), device="cuda:0")
    img = torch.randint(0, 256, (4, 64, 64, 3), dtype=torch.uint8, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    first = None
    g0 = _stat("graph_launches")
    for _ in range(4):
        d, geo = m.predict_batch(img, conf=0.1, iou=0.5, cap=128)
        cur = d.to_list()
        if first is None:
            first = cur
        for a, b in zip(cur, first):
            assert torch.equal(a["index"], b["index"]) and torch.equal(a["boxes"], b["boxes"])
    assert _stat("graph_launches") - g0 == 3
    # against the oracle on the same images
    from oracle import pre_ref
    xo = torch.from_numpy(np.concatenate([pre_ref.preprocess_ref(im, 64)[0] for im in img.cpu().numpy()]))
    lv = model_ref.forward_ref(ck["state_dict"], ck["meta"], xo)
    ref = post_ref.detect_ref([l.numpy() for l in lv], 64, 0.1, 0.5, 300)
    assert sum(len(r["index"]) for r in ref) > 0
    for a, r in zip(first, ref):
        assert np.array_equal(np.sort(a["index"].cpu().numpy()), np.sort(r["index"]))


def test_sustained_back_to_back_forwards_do_not_deadlock():
    """600 forwards of the benchmark configuration back to back (PDL on): regression test for a 1-in-1000 mbarrier phase-aliasing
    deadlock in the fused depthwise -> pointwise kernel (odd halo ring depth with two alternating producer groups); a protocol
    error traps, which surfaces here as a CUDA error."""
    ck = synth_ckpt("edge_n", 80, 640)
    eng = _engine(ck)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((64, 3, 640, 640), device="cuda", generator=g)
    outs = eng(x)
    first = [o.clone() for o in outs]
    for _ in range(600):
        eng.forward(x, out=outs)
    torch.cuda.synchronize()
    import yololite_b200 as y
    assert y.lib().yl_stat(b"trap_word10") == 0
    for a, b in zip(outs, first):
        assert torch.equal(a, b) and bool(torch.isfinite(a).all())
