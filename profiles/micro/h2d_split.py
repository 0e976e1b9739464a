"""Host->device bandwidth of one 12 MB pinned batch: one copy vs the same bytes split over 2 / 4 streams (copy engines)."""
import torch

dev = torch.device("cuda:0")
n = 12 * 1024 * 1024 // 4
src = torch.empty(n, dtype=torch.float32).pin_memory()
dst = torch.empty(n, dtype=torch.float32, device=dev)
streams = [torch.cuda.Stream() for _ in range(4)]


def run(parts, iters=200):
    chunk = n // parts
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    s.record()
    for _ in range(iters):
        for p in range(parts):
            st = streams[p]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                dst[p * chunk:(p + 1) * chunk].copy_(src[p * chunk:(p + 1) * chunk], non_blocking=True)
        for p in range(parts):
            cur.wait_stream(streams[p])
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    return ms, n * 4 / ms / 1e6


for parts in (1, 2, 4, 1, 2):
    ms, gbs = run(parts)
    print(f"parts={parts}: {ms * 1e3:.1f} us per 12 MiB = {gbs:.1f} GB/s", flush=True)
