import sys, os, torch, ctypes
sys.path.insert(0, os.getcwd())
import yololite_b200 as y
B,S,nc=64,640,80
g=torch.Generator(device='cuda').manual_seed(0)
lv=[torch.randn((B,1,s,s,5+nc),device='cuda',generator=g) for s in (80,40,20)]
for l in lv: l[...,4]-=4.0
post=y.PostProcessor()
for _ in range(3): d=post(lv,S,0.25,0.5,300,cap=1024)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d=post(lv,S,0.25,0.5,300,cap=1024)
e1.record(); torch.cuda.synchronize()
print("post ms", e0.elapsed_time(e1)/5, "dets/img", float(d.counts.float().mean()))
