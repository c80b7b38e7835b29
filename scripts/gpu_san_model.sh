#!/bin/bash
# compute-sanitizer memcheck over WHOLE forwards (every kernel of the lowered programs incl. the fused postprocess): smoke() (edge_n 320),
# and one edge_m / edge_s forward with P2 + P6 at an odd size
set -u
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_san_memcheck_smoke.log 2>&1
echo "[memcheck smoke] $(grep -E 'ERROR SUMMARY|smoke ok' gpurun_out/r02_san_memcheck_smoke.log | tr '\n' ' ')"
cat > /tmp/san_fwd.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
import yololite_b200 as y
from yololite_b200 import synth
for model, S, B, p2, p6 in (("edge_m", 320, 2, True, False), ("edge_s", 224, 3, False, True)):
    meta = synth.make_meta(model, 7, S, use_p2=p2, use_p6=p6)
    ck = synth.random_checkpoint(meta, seed=1)
    eng = y.YoloLiteB200(ck["state_dict"], meta, device="cuda:0")
    x = torch.randn((B, 3, S, S), device="cuda:0")
    outs = eng(x)
    d = y.detect(outs, S, 0.25, 0.5, 300)
    torch.cuda.synchronize()
    print(model, S, [tuple(o.shape) for o in outs], "ok")
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san_fwd.py > gpurun_out/r02_san_memcheck_models.log 2>&1
echo "[memcheck models] $(grep -E 'ERROR SUMMARY| ok$' gpurun_out/r02_san_memcheck_models.log | tr '\n' ' ')"
