"""Host->device copy bandwidth of one 78.6 MB pinned uint8 batch: one stream vs the batch split over 2 / 4 streams."""
import torch, time
dev = torch.device("cuda:0")
B, S = 64, 640
h = torch.randint(0, 255, (B, S, S, 3), dtype=torch.uint8).pin_memory()
d = torch.empty_like(h, device=dev)
for nstream in (1, 2, 4, 8):
    ss = [torch.cuda.Stream(dev) for _ in range(nstream)]
    ch = B // nstream
    def go():
        for i, s in enumerate(ss):
            with torch.cuda.stream(s):
                d[i * ch:(i + 1) * ch].copy_(h[i * ch:(i + 1) * ch], non_blocking=True)
    for _ in range(3): go()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 30
    for _ in range(n): go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"streams {nstream}: {dt*1e3:.3f} ms per batch, {h.numel()/dt/1e9:.1f} GB/s")
