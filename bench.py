#!/usr/bin/env python
"""Benchmark of the hot path: YoloLite detection forward + fused postprocess, images/s.

    python bench.py --gpus 1 --steps 20 --warmup 5                      # this engine on B200
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1      # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): edge_n, 640x640, batch 64 per GPU, nc=80, synthetic uniform-random RGB through
the reference normalisation, random-init weights.  A step = model.forward(x) + postprocess (sigmoid, decode,
score > conf, class-wise NMS) of one batch.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

MB = 1e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="edge_n")
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--img", type=int, default=640)
    ap.add_argument("--nc", type=int, default=80)
    ap.add_argument("--conf", type=float, default=0.25)
    ap.add_argument("--iou", type=float, default=0.5)
    ap.add_argument("--max-det", type=int, default=300)
    ap.add_argument("--cap", type=int, default=1024, help="detections kept per image in the output buffers")
    ap.add_argument("--cand-frac", type=float, default=0.01, help="target fraction of anchors passing conf")
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tc", action="store_true", help="fp32 SIMT kernels only (A/B against the tcgen05 path)")
    ap.add_argument("--dump-ops", default="", help="write per-op timings to this JSON file")
    return ap.parse_args()


def workload_name(a):
    return f"{a.model} {a.img}px batch={a.batch}/GPU nc={a.nc} forward+postprocess(conf={a.conf},iou={a.iou})"


def synth_input_u8(B, S, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    return torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8, device=device)


def normalise(u8):
    mean = torch.tensor([0.485, 0.456, 0.406], device=u8.device)
    std = torch.tensor([0.229, 0.224, 0.225], device=u8.device)
    return ((u8.float() / 255.0 - mean) / std).permute(0, 3, 1, 2).contiguous()


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "note": "nvml unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- CPU reference (oracle port)
def cpu_reference(ckpt, a, seconds, min_batches=1, max_batches=64):
    """The reference's CPU path restated (oracle/model_ref.forward_ref + oracle/post_ref.detect_ref), all host threads."""
    from oracle import model_ref, post_ref
    torch.set_num_threads(os.cpu_count())
    x = model_ref.synth_input(a.cpu_batch, a.img, seed=0)

    def one():
        lv = model_ref.forward_ref(ckpt["state_dict"], ckpt["meta"], x)
        post_ref.detect_ref([l.numpy() for l in lv], a.img, a.conf, a.iou, a.max_det)

    one()                                                   # warm-up (evaluate.py:253-303 uses 2; the sample is bounded)
    t0 = time.perf_counter()
    n = 0
    while n < max_batches and (n < min_batches or time.perf_counter() - t0 < seconds):
        one()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n * a.cpu_batch / dt, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} batches of {a.cpu_batch} images @{a.img}px, forward+postprocess, torch {torch.__version__} CPU fp32 "
                      f"({dt:.1f} s)"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from yololite_b200 import synth
    meta = synth.make_meta(a.model, a.nc, a.img)
    ckpt = synth.random_checkpoint(meta, seed=0)
    from oracle import model_ref, post_ref
    torch.set_num_threads(os.cpu_count())
    x = model_ref.synth_input(a.cpu_batch, a.img, seed=0)

    def step():
        lv = model_ref.forward_ref(ckpt["state_dict"], ckpt["meta"], x)
        post_ref.detect_ref([l.numpy() for l in lv], a.img, a.conf, a.iou, a.max_det)

    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    v = a.steps * a.cpu_batch / dt
    sample = f"each step = {a.cpu_batch} images @{a.img}px (bounded sample of the batch-{a.batch} workload), forward+postprocess"
    print(json.dumps({
        "impl": "reference", "metric": "images/s", "value": v, "unit": "images/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "reference CPU path = oracle port of model_v2.py forward + "
                   "utils_ms.py decode + tools/infer.py NMS loop (the reference itself needs the un-vendored timm)"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------- this engine
def calibrate_obj_bias(y, ckpt_fn, x, a):
    """Shift the objectness bias so that ~cand_frac of the anchors pass `conf` (fresh-init bias gives none)."""
    eng = y.YoloLiteB200(**ckpt_fn(None), device=x.device)
    lv = eng(x[: min(8, x.shape[0])])
    flat = torch.cat([l.reshape(-1, l.shape[-1]) for l in lv])
    obj, cls = flat[:, 4], flat[:, 5:].sigmoid().amax(-1)
    lo, hi = -20.0, 20.0
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        frac = float((((obj + mid).sigmoid() * cls) > a.conf).float().mean())
        lo, hi = (mid, hi) if frac < a.cand_frac else (lo, mid)
    eng.close()
    return -np.log(99.0) + 0.5 * (lo + hi)


def run_b200(a):
    import yololite_b200 as y
    from yololite_b200 import dist as ydist, synth
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL prints its version banner there when NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B, S = a.batch, a.img
    meta = synth.make_meta(a.model, a.nc, S)

    def ckpt_fn(obj_bias):
        ck = synth.random_checkpoint(meta, seed=0, obj_bias=obj_bias)
        return {"state_dict": ck["state_dict"], "meta": ck["meta"]}

    x = normalise(synth_input_u8(B, S, 1234 + rank, dev))
    obj_bias = calibrate_obj_bias(y, ckpt_fn, x, a)
    ck = ckpt_fn(obj_bias)
    eng = y.YoloLiteB200(**ck, device=dev, tensor_cores=not a.no_tc)
    post = y.PostProcessor()
    shapes = eng.level_shapes(B, S, S)
    outs = [torch.empty((B, A, sh, sw, D), device=dev) for (A, sh, sw, D) in shapes]
    N = sum(A * sh * sw for (A, sh, sw, D) in shapes)
    gathered = None
    if world > 1:
        packed = torch.empty((B, a.cap, 6), device=dev)
        gathered = torch.empty((world * B, a.cap, 6), device=dev)
        gcounts = torch.empty((world * B,), device=dev, dtype=torch.int32)
    extra_launches = 0

    def step(xin, pp=None):
        eng.forward(xin, out=outs)
        d = (pp or post)(outs, S, a.conf, a.iou, a.max_det, cap=a.cap)
        if world > 1:      # the path's one exchange: gather the fixed-capacity detections
            ydist.pack_detections(d.boxes, d.scores, d.classes, out=packed)
            ydist.gather_detections(packed, d.counts, out=gathered, out_counts=gcounts)
        return d

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(a.warmup, 3)):
        d = step(x)
    sync_all()
    cnt = d.counts.cpu().numpy()
    assert not (cnt & (1 << 30)).any(), "detection capacity overflow: raise --cap"
    dets_per_img = float(cnt.mean())

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        sync_all()
        e0.record()
        for _ in range(a.steps):
            step(x)
        e1.record()
        sync_all()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    value = world * B * a.steps / (ms / 1e3)
    n_ops = len(eng.program.ops)
    launches_per_step = n_ops + 1 + (3 if world > 1 else 0)

    # ---- per-kernel roofline: CUDA events around every launch, same stream, averaged over a few forwards
    peak, peak_src = measured_peak()
    reps = 5
    acc = np.zeros(n_ops)
    for _ in range(reps):
        acc += np.array([t for _, t in eng.profile_ops(x)])
    acc /= reps
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(reps):
        post(outs, S, a.conf, a.iou, a.max_det, cap=a.cap)
    p1.record()
    torch.cuda.synchronize(dev)
    post_ms = p0.elapsed_time(p1) / reps
    fwd_ms = float(acc.sum())
    kinds = {0: "stem_kernel", 1: "conv", 2: "dw_kernel", 3: "dwpw", 4: "stem+conv3x3s2"}
    # spatial sizes per op for the byte model
    per_op = []
    hw = {}

    def out_hw(h, w, k, s):
        return (h + 2 * (k // 2) - k) // s + 1, (w + 2 * (k // 2) - k) // s + 1
    for i, op in enumerate(eng.program.ops):
        hin, win = (S, S) if op["src"] < 0 else hw[op["src"]]
        if op["kind"] == 4:
            ho, wo = out_hw(*out_hw(hin, win, 3, 2), op["k"], op["stride"])
        elif op["kind"] == 3:      # fused depthwise -> pointwise: the depthwise stage (k2, stride2) sets the output size
            ho, wo = out_hw(hin, win, op["k2"], max(1, op.get("stride2", 0)))
        else:
            ho, wo = out_hw(hin, win, op["k"], op["stride"])
        if op["dst"] >= 0:
            hw[op["dst"]] = (ho, wo)
        wbytes = 4 * (op["k"] * op["k"] * op["cin"] * op["cout"] + op["cout"] + (op["k2"] * op["k2"] * op["cin"] if op["kind"] == 3 else 0))
        if op["kind"] == 2:
            wbytes = 4 * (op["k"] * op["k"] * op["cin"] + op["cout"])
        nbytes = 4 * B * (hin * win * op["cin"] + ho * wo * op["cout"]) + wbytes
        if op["res"] >= 0:
            nbytes += 4 * B * ho * wo * op["cout"]
        if op["up"] >= 0:
            nbytes += 4 * B * (ho // 2) * (wo // 2) * op["cout"]
        name = kinds[op["kind"]]
        if op["kind"] in (1, 3, 4):
            kk = op["cin"] if op["kind"] == 3 or op["k"] == 1 else op["k"] * op["k"] * (32 if op["kind"] == 4 else op["cin"])
            on_tc = (not a.no_tc) and op["wt_off"] >= 0 and (op["kind"] == 4 or (kk >= 32 and op["cout"] >= 32))
            name = ("stem2_kernel<" if op["kind"] == 4 and op["w3_off"] >= 0 else "tc_conv_kernel<" if on_tc else "conv_gemm_kernel<") + name + ">"
        per_op.append((name, i, nbytes, acc[i], op))
    cand = [(t, name, i, nb) for (name, i, nb, t, op) in per_op]
    cand.append((post_ms, "post_kernel", -1, 4 * B * N * (5 + a.nc)))
    t_top, name_top, i_top, nb_top = max(cand)
    achieved = nb_top / (t_top / 1e3) / 1e9
    desc = name_top if i_top < 0 else (f"{name_top} op#{i_top} {per_op[i_top][4]['cin']}->{per_op[i_top][4]['cout']} "
                                       f"k{per_op[i_top][4]['k']}s{per_op[i_top][4]['stride']}")
    # DRAM traffic of that kernel per launch (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of
    # the same shape, committed under profiles/; see profiles/traffic.json for the source file of each entry)
    traffic = None
    try:
        with open(os.path.join(REPO, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        key = name_top.split("<")[0] + (f":{per_op[i_top][4]['cin']}->{per_op[i_top][4]['cout']}" if i_top >= 0 else "")
        if key in tj and tj[key].get("batch") == B and tj[key].get("img") == S:
            traffic = tj[key]["dram_bytes"]
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": desc, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel_ms": t_top, "algorithmic_bytes": nb_top,
                "share_of_step": t_top / (fwd_ms + post_ms)}
    compulsory = (3 * S * S + N * (5 + a.nc)) * 4
    sum_bytes = sum(nb for (_, _, nb, _, _) in per_op) + 4 * B * N * (5 + a.nc)
    roofline_step = {"bound": "hbm", "achieved": value / world * compulsory / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": value / world * compulsory / 1e9 / peak, "bytes_per_image_compulsory": compulsory,
                     "sum_per_kernel_bytes_per_image": sum_bytes / B,
                     "frac_of_per_kernel_traffic_roofline": (sum_bytes / (fwd_ms + post_ms) * 1e3 / 1e9) / peak,
                     "forward_ms": fwd_ms, "post_ms": post_ms}
    top5 = sorted(cand, reverse=True)[:6]
    if a.dump_ops and rank == 0:
        with open(a.dump_ops, "w") as f:
            json.dump([{"i": i, "kind": nm, "cin": op["cin"], "cout": op["cout"], "k": op["k"], "s": op["stride"], "tc": op["wt_off"] >= 0,
                        "ms": float(t), "MB": nb / 1e6, "GBps": nb / (t / 1e3) / 1e9} for (nm, i, nb, t, op) in per_op] +
                      [{"i": -1, "kind": "post_kernel", "ms": post_ms}], f, indent=0)

    # ---- end to end through the public API with HOST buffers: pinned host input -> H2D -> (preprocess) -> forward+post
    #      -> D2H of the results, every step; the H2D copy of step i+1 overlaps the compute of step i (copy stream).
    e2e = e2e_fp32 = None
    if not a.no_e2e:
        hb = torch.empty((B, a.cap, 4)).pin_memory()
        hs = torch.empty((B, a.cap)).pin_memory()
        hc = torch.empty((B, a.cap), dtype=torch.int64).pin_memory()
        hn = torch.empty((B,), dtype=torch.int32).pin_memory()
        s_copy, s_comp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        xpre = [torch.empty_like(x), torch.empty_like(x)]
        posts = [y.PostProcessor(), y.PostProcessor()]      # detection buffers double-buffered against the output stage

        u8_direct = eng.supports_u8(S, S) and os.environ.get("YL_BENCH_U8_DIRECT", "1") != "0"
        sep_out = os.environ.get("YL_BENCH_SEP_OUT", "0") != "0"
        dbg_no_h2d = os.environ.get("YL_BENCH_DBG_NO_H2D", "0") == "1"      # diagnostics only: such a run is not an e2e number
        dbg_no_d2h = os.environ.get("YL_BENCH_DBG_NO_D2H", "0") == "1"      # D2H on its own stream measured slower (A/B on one box)

        def e2e_run(k, host, devbuf, from_u8):
            """Three-stage pipeline over CUDA streams, every step: (copy stream) H2D of the step's pinned host input; (compute
            stream) letterbox/normalise kernel for images, forward, postprocess; (output stream) D2H of the detections into
            pinned host buffers.  Step i+1's H2D and step i-1's D2H overlap step i's compute."""
            st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            st.record(s_copy)
            keep = [None, None]
            for i in range(k):
                j = i & 1
                with torch.cuda.stream(s_copy):
                    if i >= 2:
                        s_copy.wait_event(freed[j])
                    if not dbg_no_h2d:
                        devbuf[j].copy_(host, non_blocking=True)
                    copied[j].record(s_copy)
                with torch.cuda.stream(s_comp):
                    s_comp.wait_event(copied[j])
                    if i >= 2:
                        s_comp.wait_event(done[j])          # the D2H of step i-2 has read its detection buffers' slot
                    if from_u8 and u8_direct:
                        # 640x640 images need no letterbox resize: the stem kernel reads the uint8 BGR batch itself (BGR->RGB,
                        # /255, (x-mean)/std folded into its weights) -- the path YoloLite.predict_batch takes
                        eng.forward_u8(devbuf[j], out=outs)
                        dd = posts[j](outs, S, a.conf, a.iou, a.max_det, cap=a.cap)
                        if world > 1:
                            ydist.pack_detections(dd.boxes, dd.scores, dd.classes, out=packed)
                            ydist.gather_detections(packed, dd.counts, out=gathered, out_counts=gcounts)
                    else:
                        if from_u8:      # uint8 HWC BGR -> letterbox + RGB + normalise + CHW on the GPU
                            xin, _ = y.preprocess_batch(devbuf[j], S, out=xpre[j])
                        else:
                            xin = devbuf[j]
                        dd = step(xin, posts[j])
                    freed[j].record(s_comp)
                with torch.cuda.stream(s_out if sep_out else s_comp):
                    if sep_out:
                        s_out.wait_event(freed[j])
                    if not dbg_no_d2h:
                        hb.copy_(dd.boxes, non_blocking=True); hs.copy_(dd.scores, non_blocking=True)
                        hc.copy_(dd.classes, non_blocking=True); hn.copy_(dd.counts, non_blocking=True)
                    done[j].record(s_out if sep_out else s_comp)
                    keep[j] = dd                             # keep the device results alive until their copy has been issued twice over
            s_comp.wait_stream(s_out)
            en.record(s_comp)
            torch.cuda.synchronize(dev)
            ms_e = st.elapsed_time(en)
            if world > 1:
                t = torch.tensor([ms_e], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_e = float(t)
            return ms_e

        d2h = int(hb.numel() * 4 + hs.numel() * 4 + hc.numel() * 8 + hn.numel() * 4)
        u8h = synth_input_u8(B, S, 1234 + rank, dev).cpu().pin_memory()
        u8d = [torch.empty((B, S, S, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        if dbg_no_h2d:
            for d_ in u8d:
                d_.copy_(u8h)
        e2e_run(3, u8h, u8d, True)
        ms_e = e2e_run(a.steps, u8h, u8d, True)
        e2e = {"value": world * B * a.steps / (ms_e / 1e3), "unit": "images/s", "h2d_bytes_per_step": int(u8h.numel()),
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e / a.steps,
               "api": ("YoloLiteB200.forward_u8(uint8 HWC BGR; normalisation folded into the stem kernel)" if u8_direct else
                       "preprocess_batch(uint8 HWC BGR) + YoloLiteB200.forward") + " + PostProcessor (= YoloLite.predict_batch) on pinned "
                      "host images; stream pipeline (H2D of step i+1 on a copy stream | GPU letterbox/normalise + forward + postprocess + D2H of the detections), every "
                      "stage runs every step inside the timed region"}
        xh = torch.empty((B, 3, S, S), dtype=torch.float32).pin_memory()
        xh.copy_(x.cpu())
        xd = [torch.empty_like(x), torch.empty_like(x)]
        e2e_run(3, xh, xd, False)
        ms_f = e2e_run(a.steps, xh, xd, False)
        e2e_fp32 = {"value": world * B * a.steps / (ms_f / 1e3), "unit": "images/s", "h2d_bytes_per_step": int(xh.numel() * 4),
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_f / a.steps,
                    "api": "YoloLiteB200.forward(x) + PostProcessor on pinned fp32 normalised host input (the reference's "
                           "model(x) signature); bound by the 314.6 MB/step PCIe copy"}
        del xd, xh

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_reference(ck, a, a.cpu_seconds)

    if rank == 0:
        out = {
            "metric": "images/s", "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "global_batch": world * B, "anchors_per_image": N,
                       "detections_per_image": dets_per_img, "l2": "inputs_exceed_l2 (314.6 MB fp32 input batch per GPU)",
                       "parallelism": f"dp{world} (batch sharded, weights replicated, one NCCL all_gather of [B,{a.cap},6] dets)"
                       if world > 1 else "single GPU", "weights": "random-init (He), BN identity, obj bias calibrated to "
                       f"{a.cand_frac:.1%} candidates"},
            "clocks": clk.summary(), "e2e": e2e, "e2e_fp32_input": e2e_fp32, "gpu_launches": launches_per_step * a.steps,
            "roofline": roofline, "roofline_step": roofline_step,
            "top_kernels_ms": [{"kernel": nm if i < 0 else f"{nm}#{i}", "ms": float(t), "GBps": nb / (t / 1e3) / 1e9}
                               for (t, nm, i, nb) in top5],
            "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            sys.stderr.write("bench.py: no CUDA device; this engine has no CPU path (use --impl reference for the CPU arm)\n")
            sys.exit(1)
        run_b200(args)
