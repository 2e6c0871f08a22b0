"""Host->device copy bandwidth from pinned memory on this box: one large copy, the e2e step's tensor list, one arena."""
import torch, time
dev = "cuda"
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for mb in (1, 4, 17.7, 66.7, 256):
    h = torch.empty(int(mb * 1e6) // 4, dtype=torch.float32).pin_memory(); d = torch.empty_like(h, device=dev)
    ms = t(lambda: d.copy_(h, non_blocking=True))
    print("single copy %6.1f MB: %.3f ms  %.1f GB/s" % (mb, ms, mb / ms))
sizes = [17.7] * 3 + [4.42, 1.1, 0.28] + [5.9, 1.47, 0.37, 0.09] + [0.001] * 8
hs = [torch.empty(int(s * 1e6) // 4 + 1, dtype=torch.float32).pin_memory() for s in sizes]
ds = [torch.empty_like(h, device=dev) for h in hs]
ms = t(lambda: [d.copy_(h, non_blocking=True) for d, h in zip(ds, hs)])
print("tensor list (%d copies, %.1f MB): %.3f ms  %.1f GB/s" % (len(sizes), sum(sizes), ms, sum(sizes) / ms))
ms = t(lambda: [h.to(dev, non_blocking=True) for h in hs])
print("tensor list via .to(): %.3f ms  %.1f GB/s" % (ms, sum(sizes) / ms))
