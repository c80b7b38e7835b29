"""Diagnostic (build with NVCC_EXTRA=-DYL_TIMELINE): in-kernel %globaltimer stamps of CTA 0 for ONE launch of a tc kernel class."""
import ctypes, json, os, subprocess, sys
sys.path.insert(0, os.getcwd())
import torch
sys.argv = [sys.argv[0]] + sys.argv[1:]
import importlib.util
spec = importlib.util.spec_from_file_location("bench_op", os.path.join(os.getcwd(), "scripts", "bench_op.py"))
bo = importlib.util.module_from_spec(spec); spec.loader.exec_module(bo)
from yololite_b200 import _lib as L
lib = L.lib()
NAMES = {0: "entry", 1: "prologue done", 2: "dependency resolved", 3: "first A slab landed", 4: "MMA: first stage ready", 5: "first tile MMAs issued",
         6: "first accumulator ready", 8: "last tile MMAs issued", 9: "last tile handed to store engine", 10: "stores complete", 11: "teardown barrier"}
def run(label, argv):
    sys.argv = ["bench_op.py"] + argv + ["--iters", "1"]
    os.environ["YL_BENCH_OP_WARMUP"] = "3"
    bo.main()                                  # warm: instruction cache, L2-resident weights
    torch.cuda.synchronize()
    lib.yl_stat(b"debug_reset")
    os.environ["YL_BENCH_OP_WARMUP"] = "0"
    bo.main()                                  # exactly ONE launch is stamped
    torch.cuda.synchronize()
    w = {k: lib.yl_stat(b"trap_word%d" % (16 + k)) for k in NAMES}
    t0 = w[0]
    print(label)
    for k in sorted(NAMES):
        if w[k]: print(f"   {NAMES[k]:34s} +{(w[k]-t0)/1000:8.2f} us")
for label, argv in (("pw 64->256 @20x20 b64", "--kind conv --cin 64 --cout 256 --hw 20 --act 1 --tc 1".split()),
                    ("pw 48->96 @40x40 b64", "--kind conv --cin 48 --cout 96 --hw 40 --act 1 --tc 1".split()),
                    ("dwpw k5 256->64 @20x20 b64", "--kind dwpw --cin 256 --cout 64 --hw 20 --k2 5 --act 0 --act2 1 --res 1 --tc 1".split()),
                    ("dwpw k3 96->96 @80x80 b64", "--kind dwpw --cin 96 --cout 96 --hw 80 --k2 3 --act 1 --tc 1".split()),
                    ("pw 64->256 @20x20 b1", "--kind conv --cin 64 --cout 256 --hw 20 --act 1 --tc 1 --batch 1".split())):
    run(label, argv)
