#!/usr/bin/env python
"""Benchmark of the hot path: YoloLite detection forward + fused postprocess, images/s.

    python bench.py --gpus 1 --steps 100 --warmup 10                    # this engine on B200
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1      # the reference's CPU path (oracle port)

Headline workload (BASELINE.json configs[1]): edge_n, 640x640, batch 64 per GPU, nc=80, synthetic uniform-random RGB through
the reference normalisation, random-init weights.  A step = model.forward(x) + postprocess (sigmoid, decode, score > conf,
class-wise NMS) of one batch = ONE call of the engine's detect entry (one CUDA graph launch).  Prints ONE JSON line (rank 0);
the other BASELINE configurations that fit this box ride along under "configs": config 3's shard (edge_m, 640 px, 32 images per
GPU), config 0 (edge_n, 320 px, batch 1), config 5's FPN + heads (yololite_m + P2 from backbone features), and, for N > 1, the
strong-scaling run (global batch 64 split over the ranks).
"""
import argparse
import ctypes
import json
import os
import sys
import tempfile
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
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="edge_n")
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--img", type=int, default=640)
    ap.add_argument("--nc", type=int, default=80)
    ap.add_argument("--conf", type=float, default=0.25)
    ap.add_argument("--iou", type=float, default=0.5)
    ap.add_argument("--max-det", type=int, default=300)
    ap.add_argument("--cap", type=int, default=300, help="detections kept per image in the output / gather buffers (SURVEY 8e: max_det)")
    ap.add_argument("--cand-frac", type=float, default=0.01, help="target fraction of anchors passing conf")
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra configurations (config 0 / 3 / 5, strong scaling, torch eager)")
    ap.add_argument("--no-tc", action="store_true", help="fp32 SIMT kernels only (A/B against the tcgen05 path)")
    ap.add_argument("--no-graph", action="store_true", help="launch the ops eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-pdl", action="store_true", help="no programmatic dependent launch")
    ap.add_argument("--dump-ops", default="", help="write per-op timings to this JSON file")
    return ap.parse_args()


def workload_name(model, img, batch, nc, conf, iou):
    return f"{model} {img}px batch={batch}/GPU nc={nc} forward+postprocess(conf={conf},iou={iou})"


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
            time.sleep(0.01)

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
            j = json.load(f)
        return float(j["hbm_gbs"]), float(j.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- host memory placement
def gpu_numa_node(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        with open(f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node") as f:
            return int(f.read().strip())
    except Exception:
        return -1


class NumaLocal:
    """Prefer the GPU's NUMA node for the pinned staging buffers allocated inside the `with` block (set_mempolicy, best effort)."""

    def __init__(self, node):
        self.node, self.ok = node, False

    def __enter__(self):
        if self.node < 0:
            return self
        try:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = (ctypes.c_ulong * 16)()
            mask[self.node // 64] = 1 << (self.node % 64)
            self.ok = libc.syscall(238, 1, mask, 1024) == 0          # SYS_set_mempolicy, MPOL_PREFERRED (x86-64)
            self._libc = libc
        except Exception:
            self.ok = False
        return self

    def __exit__(self, *a):
        if self.ok:
            self._libc.syscall(238, 0, None, 0)                      # MPOL_DEFAULT


# ---------------------------------------------------------------------------------------------- CPU reference (oracle port)
def cpu_reference(ckpt, model, img, conf, iou, max_det, cpu_batch, seconds, min_batches=1, max_batches=64):
    """The reference's CPU path restated (oracle/model_ref.forward_ref + oracle/post_ref.detect_ref), all host threads."""
    from oracle import model_ref, post_ref
    torch.set_num_threads(os.cpu_count())
    x = model_ref.synth_input(cpu_batch, img, seed=0)

    def one():
        lv = model_ref.forward_ref(ckpt["state_dict"], ckpt["meta"], x)
        post_ref.detect_ref([l.numpy() for l in lv], img, conf, iou, max_det)

    one()                                                   # warm-up (evaluate.py:253-303 uses 2; the sample is bounded)
    t0 = time.perf_counter()
    n = 0
    while n < max_batches and (n < min_batches or time.perf_counter() - t0 < seconds):
        one()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n * cpu_batch / dt, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} batches of {cpu_batch} images @{img}px, forward+postprocess, torch {torch.__version__} CPU fp32 "
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
        "config": {"workload": workload_name(a.model, a.img, a.batch, a.nc, a.conf, a.iou),
                   "note": "reference CPU path = oracle port of model_v2.py forward + "
                   "utils_ms.py decode + tools/infer.py NMS loop (the reference itself needs the un-vendored timm)"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------- torch eager on the same GPU
def torch_eager_gpu(ck, x, img, conf, iou, max_det, steps=3):
    """The reference's own arithmetic on the SAME B200 through PyTorch eager (cuDNN / ATen kernels, TF32 off): the oracle's
    functional forward (model_v2.py restated op by op) + the reference's per-image postprocess loop (utils_ms.py decode,
    tools/infer.py:466-493 with torchvision.ops.nms).  This is the library path the hand-written kernels have to beat
    (BASELINE.md section 3.6); it is a baseline leg, never part of the product."""
    from oracle import model_ref
    try:
        from torchvision.ops import nms
    except Exception:
        return {"unavailable": "torchvision missing"}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = x.device
    sd = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in ck["state_dict"].items()}
    strides = model_ref.strides_ref(ck["meta"])

    def post(levels):
        out = []
        boxes, objs, clss = [], [], []
        for lv, st in zip(levels, strides):                        # utils_ms.py:25-123
            B, A, S, _, D = lv.shape
            gy, gx = torch.meshgrid(torch.arange(S, device=dev), torch.arange(S, device=dev), indexing="ij")
            t = lv.reshape(B, A * S * S, D)
            g = torch.stack([gx, gy], -1).reshape(1, S * S, 2).repeat(1, A, 1).float()
            stride = img / float(S)
            pxy = ((t[..., 0:2].sigmoid() * 2 - 0.5) + g) * stride
            pwh = torch.nn.functional.softplus(t[..., 2:4]) * stride
            b = torch.cat([pxy - pwh * 0.5, pxy + pwh * 0.5], -1).clamp(0, img - 1)
            boxes.append(b); objs.append(t[..., 4]); clss.append(t[..., 5:])
        box, obj, cls = torch.cat(boxes, 1), torch.cat(objs, 1), torch.cat(clss, 1)
        for b in range(box.shape[0]):                              # tools/infer.py:466-493
            confs, ci = cls[b].sigmoid().max(-1)
            sc = obj[b].sigmoid() * confs
            m = sc > conf
            bb, ss, cc = box[b][m], sc[m], ci[m]
            keep_b, keep_s, keep_c = [], [], []
            for c in cc.unique():
                mc = cc == c
                k = nms(bb[mc], ss[mc], iou)[:max_det]
                keep_b.append(bb[mc][k]); keep_s.append(ss[mc][k]); keep_c.append(torch.full((k.numel(),), int(c), device=dev))
            out.append((torch.cat(keep_b) if keep_b else bb[:0], torch.cat(keep_s) if keep_s else ss[:0]))
        return out

    def step():
        with torch.no_grad():
            post(model_ref.forward_ref(sd, ck["meta"], x))

    step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    f0.record()
    for _ in range(steps):
        with torch.no_grad():
            model_ref.forward_ref(sd, ck["meta"], x)
    f1.record()
    torch.cuda.synchronize(dev)
    B = x.shape[0]
    return {"value": B * steps / (e0.elapsed_time(e1) / 1e3), "unit": "images/s", "forward_only_images_per_s": B * steps / (f0.elapsed_time(f1) / 1e3),
            "ms_per_step": e0.elapsed_time(e1) / steps, "what": f"oracle functional forward (torch {torch.__version__} eager, cuDNN, TF32 off) + the reference's "
            "per-image decode / torchvision.ops.nms loop on the same GPU, device-resident input"}


# ---------------------------------------------------------------------------------------------- this engine
def calibrate_obj_bias(y, ckpt_fn, x, conf, cand_frac):
    """Shift the objectness bias so that ~cand_frac of the anchors pass `conf` (fresh-init bias gives none)."""
    eng = y.YoloLiteB200(**ckpt_fn(None), device=x.device)
    lv = eng(x[: min(8, x.shape[0])])
    flat = torch.cat([l.reshape(-1, l.shape[-1]) for l in lv])
    obj, cls = flat[:, 4], flat[:, 5:].sigmoid().amax(-1)
    lo, hi = -20.0, 20.0
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        frac = float((((obj + mid).sigmoid() * cls) > conf).float().mean())
        lo, hi = (mid, hi) if frac < cand_frac else (lo, mid)
    eng.close()
    return -np.log(99.0) + 0.5 * (lo + hi)


class Ctx:
    pass


def setup_dist():
    c = Ctx()
    c.rank = int(os.environ.get("RANK", "0"))
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    c.dev = torch.device(f"cuda:{c.local}")
    torch.cuda.set_device(c.dev)
    c.dist = None
    if c.world > 1:
        # keep stdout to the ONE JSON line: NCCL prints its version banner there when NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # whatever NCCL prints goes to stderr
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=c.dev)
        c.dist = dist
    return c


def sync_all(c):
    torch.cuda.synchronize(c.dev)
    if c.world > 1:
        c.dist.barrier()
        torch.cuda.synchronize(c.dev)


def max_over_ranks(c, ms):
    if c.world > 1:
        t = torch.tensor([ms], device=c.dev)
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
        return float(t)
    return ms


def measure(c, a, model, B, S, steps, warmup, full, clocks=False):
    """One configuration: device-resident `value`, per-kernel roofline, end-to-end through the public API.  `full` adds the
    per-op profile, the fp32-input e2e variant and the CPU baseline."""
    import yololite_b200 as y
    from yololite_b200 import dist as ydist, synth
    dev, world, rank = c.dev, c.world, c.rank
    meta = synth.make_meta(model, a.nc, S)

    def ckpt_fn(obj_bias):
        ck = synth.random_checkpoint(meta, seed=0, obj_bias=obj_bias)
        return {"state_dict": ck["state_dict"], "meta": ck["meta"]}

    x = normalise(synth_input_u8(B, S, 1234 + rank, dev))
    obj_bias = calibrate_obj_bias(y, ckpt_fn, x, a.conf, a.cand_frac)
    ck = ckpt_fn(obj_bias)
    eng = y.YoloLiteB200(**ck, device=dev, tensor_cores=not a.no_tc, graph=not a.no_graph, pdl=not a.no_pdl)
    shapes = eng.level_shapes(B, S, S)
    N = sum(A * sh * sw for (A, sh, sw, D) in shapes)
    cap = a.cap
    # the kernel writes the gather payload itself: [B, cap+1, 6] (row 0 = count / overflow / K), double-buffered so that the
    # all-gather of step i (NCCL stream) overlaps the forward of step i+1
    packed = [torch.zeros((B, cap + 1, 6), device=dev) for _ in range(2)]
    gathered = [torch.empty((world * B, cap + 1, 6), device=dev) for _ in range(2)] if world > 1 else None
    works = [None, None]

    def step(xin, i):
        j = i & 1
        if works[j] is not None:
            works[j].wait()                                   # the gather that read packed[j] two steps ago
            works[j] = None
        eng.detect(xin, S, a.conf, a.iou, a.max_det, cap=cap, packed=packed[j])
        if world > 1:      # the path's ONE exchange: all-gather of the kernel-written payload, asynchronous
            _, works[j] = ydist.gather_packed(packed[j], out=gathered[j], async_op=True)

    def drain():
        for j in range(2):
            if works[j] is not None:
                works[j].wait()
                works[j] = None

    for i in range(max(warmup, 3)):
        step(x, i)
    drain()
    sync_all(c)
    head = packed[(max(warmup, 3) - 1) & 1][:, 0, :3].cpu().numpy()
    assert not (head[:, 1] != 0).any(), "detection capacity overflow: raise --cap"
    dets_per_img = float(head[:, 0].mean())

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(c.local) if clocks else None
    if clk:
        clk.__enter__()
    sync_all(c)
    e0.record()
    for i in range(steps):
        step(x, i)
    drain()
    e1.record()
    sync_all(c)
    if clk:
        clk.__exit__()
    ms = max_over_ranks(c, e0.elapsed_time(e1))
    value = world * B * steps / (ms / 1e3)
    n_ops = len(eng.program.ops)
    res = {"value": value, "ms_per_step": ms / steps, "steps": steps, "anchors_per_image": N, "detections_per_image": dets_per_img,
           "kernels_per_step": n_ops + 1, "workload": workload_name(model, S, B, a.nc, a.conf, a.iou),
           "clocks": clk.summary() if clk else None}

    # ---- per-kernel roofline: CUDA events around every launch (no PDL / graph there), same stream, averaged over a few forwards
    peak, _, peak_src = measured_peak()
    reps = 5
    acc = np.zeros(n_ops)
    for _ in range(reps):
        acc += np.array([t for _, t in eng.profile_ops(x)])
    acc /= reps
    post = y.PostProcessor()
    outs = eng(x)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    post(outs, S, a.conf, a.iou, a.max_det, cap=cap)
    p0.record()
    for _ in range(reps):
        post(outs, S, a.conf, a.iou, a.max_det, cap=cap)
    p1.record()
    torch.cuda.synchronize(dev)
    post_ms = p0.elapsed_time(p1) / reps
    del outs
    fwd_ms = float(acc.sum())
    kinds = {0: "stem_kernel", 1: "conv", 2: "dw_kernel", 3: "dwpw", 4: "stem+conv3x3s2"}
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
        if key in tj and tj[key].get("batch") == B and tj[key].get("img") == S and tj[key].get("model", "edge_n") == model:
            traffic = tj[key]["dram_bytes"]
    except Exception:
        traffic = None
    res["roofline"] = {"bound": "hbm", "kernel": desc, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                       "traffic": traffic, "peak_source": peak_src, "kernel_ms": t_top, "algorithmic_bytes": nb_top,
                       "share_of_step": t_top / (fwd_ms + post_ms)}
    compulsory = (3 * S * S + N * (5 + a.nc)) * 4
    sum_bytes = sum(nb for (_, _, nb, _, _) in per_op) + 4 * B * N * (5 + a.nc)
    res["roofline_step"] = {"bound": "hbm", "achieved": value / world * compulsory / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": value / world * compulsory / 1e9 / peak, "bytes_per_image_compulsory": compulsory,
                            "sum_per_kernel_bytes_per_image": sum_bytes / B,
                            "frac_of_per_kernel_traffic_roofline": (sum_bytes / (ms / steps) * 1e3 / 1e9) / peak,
                            "forward_ms_sum_of_kernels": fwd_ms, "post_ms": post_ms}
    res["top_kernels_ms"] = [{"kernel": nm if i < 0 else f"{nm}#{i}", "ms": float(t), "GBps": nb / (t / 1e3) / 1e9}
                             for (t, nm, i, nb) in sorted(cand, reverse=True)[:6]]
    if a.dump_ops and rank == 0 and full:
        with open(a.dump_ops, "w") as f:
            json.dump([{"i": i, "kind": nm, "cin": op["cin"], "cout": op["cout"], "k": op["k"], "s": op["stride"], "tc": op["wt_off"] >= 0,
                        "ms": float(t), "MB": nb / 1e6, "GBps": nb / (t / 1e3) / 1e9} for (nm, i, nb, t, op) in per_op] +
                      [{"i": -1, "kind": "post_kernel", "ms": post_ms}], f, indent=0)

    # ---- end to end through the public API: YoloLite(weights).predict_batch(images) on pinned HOST uint8 images, every step:
    #      H2D of the step's images (copy stream, overlapping the previous step's compute) -> predict_batch (letterbox / normalise
    #      folded into the stem kernel, forward, postprocess: one CUDA graph launch) -> D2H of the detections.
    if not a.no_e2e:
        tmpd = tempfile.mkdtemp(prefix="yl_bench_")
        wpath = os.path.join(tmpd, "ckpt.pt")
        torch.save({"state_dict": ck["state_dict"], "meta": ck["meta"]}, wpath)
        m = y.YoloLite(wpath, device=dev, graph=not a.no_graph)
        node = gpu_numa_node(c.local)
        with NumaLocal(node) as numa:
            u8h = synth_input_u8(B, S, 1234 + rank, dev).cpu().pin_memory()
            hb = torch.empty((B, cap, 4)).pin_memory()
            hs = torch.empty((B, cap)).pin_memory()
            hc = torch.empty((B, cap), dtype=torch.int64).pin_memory()
            hn = torch.empty((B,), dtype=torch.int32).pin_memory()
            xh = torch.empty((B, 3, S, S), dtype=torch.float32).pin_memory() if full else None
        s_copy, s_comp = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]
        u8d = [torch.empty((B, S, S, 3), dtype=torch.uint8, device=dev) for _ in range(2)]

        dbg = os.environ.get("YL_BENCH_DBG", "").split(",")      # diagnostics only (noh2d / nod2h): such a run is not an e2e number

        def e2e_run(k, host, devbuf, fp32):
            st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            if world > 1:
                c.dist.barrier()
            st.record(s_copy)
            for i in range(k):
                j = i & 1
                with torch.cuda.stream(s_copy):
                    if i >= 2:
                        s_copy.wait_event(freed[j])
                    if "noh2d" not in dbg:
                        devbuf[j].copy_(host, non_blocking=True)
                    copied[j].record(s_copy)
                with torch.cuda.stream(s_comp):
                    s_comp.wait_event(copied[j])
                    if works[j] is not None:
                        works[j].wait()
                        works[j] = None
                    if fp32:       # the reference's model(x) signature: normalised fp32 NCHW from the host
                        bx, sc, cl, ix, cn = m.model.detect(devbuf[j], S, a.conf, a.iou, a.max_det, cap=cap, outputs=e2e_out)
                    else:
                        d, _ = m.predict_batch(devbuf[j], conf=a.conf, iou=a.iou, max_det=a.max_det, cap=cap,
                                               packed=packed[j] if world > 1 else None)
                        bx, sc, cl, cn = d.boxes, d.scores, d.classes, d.counts
                    freed[j].record(s_comp)
                    if world > 1 and not fp32:
                        _, works[j] = ydist.gather_packed(packed[j], out=gathered[j], async_op=True)
                    if "nod2h" not in dbg:
                        hb.copy_(bx, non_blocking=True); hs.copy_(sc, non_blocking=True)
                        hc.copy_(cl, non_blocking=True); hn.copy_(cn, non_blocking=True)
            with torch.cuda.stream(s_comp):
                drain()
                en.record(s_comp)
            torch.cuda.synchronize(dev)
            return max_over_ranks(c, st.elapsed_time(en))

        d2h = int(hb.numel() * 4 + hs.numel() * 4 + hc.numel() * 8 + hn.numel() * 4)
        e2e_steps = min(steps, 100)
        e2e_run(4, u8h, u8d, False)
        ms_e = e2e_run(e2e_steps, u8h, u8d, False)
        res["e2e"] = {"value": world * B * e2e_steps / (ms_e / 1e3), "unit": "images/s", "h2d_bytes_per_step": int(u8h.numel()),
                      "d2h_bytes_per_step": d2h, "ms_per_step": ms_e / e2e_steps,
                      "h2d_GBps_aggregate": world * u8h.numel() / (ms_e / e2e_steps) / 1e6,
                      "pinned_numa": {"gpu_node": node, "mempolicy_preferred_set": bool(numa.ok)},
                      "api": "YoloLite(weights).predict_batch(uint8 HWC BGR images) on pinned host images: H2D of step i+1 on a copy stream | "
                             "one engine call (stem kernel reads the uint8 image, forward, fused postprocess; one CUDA graph launch) | D2H "
                             "of boxes / scores / classes / counts; every stage runs every step inside the timed region"}
        if full:
            xh.copy_(x.cpu())
            xd = [torch.empty_like(x), torch.empty_like(x)]
            e2e_out = (torch.empty((B, cap, 4), device=dev), torch.empty((B, cap), device=dev), torch.empty((B, cap), device=dev, dtype=torch.int64),
                       torch.empty((B, cap), device=dev, dtype=torch.int64), torch.zeros((B,), device=dev, dtype=torch.int32))
            f_steps = min(steps, 30)
            e2e_run(3, xh, xd, True)
            ms_f = e2e_run(f_steps, xh, xd, True)
            res["e2e_fp32_input"] = {"value": world * B * f_steps / (ms_f / 1e3), "unit": "images/s", "h2d_bytes_per_step": int(xh.numel() * 4),
                                     "d2h_bytes_per_step": d2h, "ms_per_step": ms_f / f_steps,
                                     "h2d_GBps_aggregate": world * xh.numel() * 4 / (ms_f / f_steps) / 1e6,
                                     "api": "YoloLiteB200.detect(x) on pinned fp32 normalised host input (the reference's model(x) signature); "
                                            "bound by the host->device copy of the fp32 batch"}
            del xd
        del m

    if full and rank == 0 and world == 1 and not a.no_cpu_baseline:
        res["cpu_baseline"] = cpu_reference(ck, model, S, a.conf, a.iou, a.max_det, a.cpu_batch, a.cpu_seconds)
    if full and rank == 0 and world == 1 and not a.no_extra:
        res["torch_eager_gpu"] = torch_eager_gpu(ck, x[: min(B, 16)], S, a.conf, a.iou, a.max_det)
    eng.close()
    return res


def measure_features(c, a, steps):
    """BASELINE config 5's FPN + heads (yololite_m + P2, fpn 328, 34 000 anchors) from synthetic backbone features: the tensor-pipe
    bound part of the path (74.7 GMAC / image, 65.8 of them dense 3x3)."""
    import yololite_b200 as y
    from yololite_b200 import synth
    dev = c.dev
    B, S = 4, 640
    meta = synth.make_meta("yololite_m", a.nc, S, use_p2=True)
    chs = synth.FEATURE_CHANNELS[meta["backbone"]]
    ck = synth.random_checkpoint(meta, seed=0, feat_chs=chs)
    eng = y.YoloLiteB200(ck["state_dict"], meta, device=dev, from_features=True, graph=not a.no_graph)
    g = torch.Generator(device=dev).manual_seed(5)
    feats = [torch.randn((B, ch, S // r, S // r), device=dev, generator=g).abs_().contiguous(memory_format=torch.channels_last)
             for ch, r in zip(chs, (4, 8, 16, 32))]
    outs = eng.forward_features(feats)
    post = y.PostProcessor()
    N = sum(o.shape[1] * o.shape[2] * o.shape[3] for o in outs)

    def step():
        eng.forward_features(feats, out=outs)
        post(outs, S, a.conf, a.iou, a.max_det, cap=a.cap)

    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    macs = 0
    px = {4: (S // 4) ** 2, 8: (S // 8) ** 2, 16: (S // 16) ** 2, 32: (S // 32) ** 2}
    red = {}
    for op in eng.program.ops:
        r = {-2: 4, -3: 8, -4: 16, -5: 32}[op["src"]] if op["src"] <= -2 else red[op["src"]]
        if op["dst"] >= 0:
            red[op["dst"]] = r
        macs += px[r] * op["cout"] * op["cin"] * op["k"] ** 2 + (px[r] * op["cin"] * op["k2"] ** 2 if op["kind"] == 3 else 0)
    _, tf_peak, src = measured_peak()
    tflops = 2 * macs * B / (ms / 1e3) / 1e12
    eng.close()
    return {"workload": f"yololite_m +P2 FPN+heads from backbone features, {S}px batch={B} nc={a.nc} ({N} anchors) + postprocess",
            "value": B / (ms / 1e3), "unit": "images/s", "ms_per_step": ms, "gmac_per_image": macs / 1e9,
            "roofline": {"bound": "tensor", "achieved": tflops, "peak": tf_peak / 3.0, "unit": "TFLOP/s", "frac": tflops / (tf_peak / 3.0),
                         "traffic": None, "peak_source": src + ": dense bf16 burst / 3 (six bf16 products per fp32 multiply at twice the "
                         "fp32-equivalent k depth = 3 tensor-pipe passes per useful FLOP)", "useful_tflops": tflops}}


def run_b200(a):
    c = setup_dist()
    B, S = a.batch, a.img
    main = measure(c, a, a.model, B, S, a.steps, a.warmup, full=True, clocks=True)
    extras = {}
    if not a.no_extra:
        if a.model == "edge_n" and S == 640:
            # BASELINE config 3: edge_m, 640 px, global batch 256 over 8 GPUs = 32 images per GPU
            extras["config3_edge_m_640_b32"] = measure(c, a, "edge_m", 32, 640, max(10, a.steps // 4), 5, full=False)
            # BASELINE config 0: edge_n, 320 px, batch 1 (the reference CLI's per-image loop, tools/infer.py:435-456)
            extras["config0_edge_n_320_b1"] = measure(c, a, "edge_n", 1, 320, max(50, a.steps), 10, full=False)
            if c.world == 1:
                try:
                    extras["config5_yololite_m_p2_fpn_heads"] = measure_features(c, a, max(3, a.steps // 20))
                except Exception as ex:      # never lose the headline line to an extra
                    extras["config5_yololite_m_p2_fpn_heads"] = {"error": repr(ex)[:300]}
        if c.world > 1 and B % c.world == 0:
            # strong scaling: the SAME global batch of 64 split over the ranks (SURVEY.md section 8d "report both")
            s = measure(c, a, a.model, B // c.world, S, a.steps, a.warmup, full=False)
            extras["strong_scaling"] = {"global_batch": B, "per_gpu_batch": B // c.world, "value": s["value"], "ms_per_step": s["ms_per_step"],
                                        "e2e": s.get("e2e", {}).get("value")}
    if c.rank == 0:
        world = c.world
        out = {
            "metric": "images/s", "value": main["value"], "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": main["workload"], "global_batch": world * B, "anchors_per_image": main["anchors_per_image"],
                       "detections_per_image": main["detections_per_image"], "l2": "inputs_exceed_l2 (314.6 MB fp32 input batch per GPU)",
                       "parallelism": f"dp{world} (batch sharded, weights replicated, ONE async NCCL all_gather of the kernel-written [B,{a.cap + 1},6] payload per step)"
                       if world > 1 else "single GPU", "weights": "random-init (He), BN identity, obj bias calibrated to "
                       f"{a.cand_frac:.1%} candidates", "launch": ("one CUDA graph launch per step" if not a.no_graph else "eager launches") +
                       (", programmatic dependent launch" if not a.no_pdl else "")},
            "clocks": main["clocks"], "e2e": main.get("e2e"), "e2e_fp32_input": main.get("e2e_fp32_input"),
            "gpu_launches": main["kernels_per_step"] * a.steps,
            "roofline": main["roofline"], "roofline_step": main["roofline_step"], "top_kernels_ms": main["top_kernels_ms"],
            "cpu_baseline": main.get("cpu_baseline"), "torch_eager_gpu": main.get("torch_eager_gpu"),
            "configs": extras,
        }
        print(json.dumps(out))
    if c.world > 1:
        c.dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            sys.stderr.write("bench.py: no CUDA device; this engine has no CPU path (use --impl reference for the CPU arm)\n")
            sys.exit(1)
        run_b200(args)
