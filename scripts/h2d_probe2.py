"""Pinned H2D of one 66.7 MB arena: one copy vs. split over 2-4 streams."""
import torch
dev = "cuda"
n = 66726912
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream() for _ in range(4)]
def run(parts):
    evs = []
    step = (n // parts + 255) // 256 * 256
    for p in range(parts):
        lo, hi = p * step, min(n, (p + 1) * step)
        with torch.cuda.stream(streams[p]):
            d[lo:hi].copy_(h[lo:hi], non_blocking=True)
            e = torch.cuda.Event(); e.record(); evs.append(e)
    for e in evs: torch.cuda.current_stream().wait_event(e)
for parts in (1, 2, 3, 4, 1):
    for _ in range(3): run(parts)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(30): run(parts)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 30
    print("parts %d: %.3f ms  %.1f GB/s" % (parts, ms, n / ms / 1e6))
