"""Host<->device copy bandwidth of the box (pinned memory, 1 GiB, CUDA events): the floor under the end-to-end number of bench.py,
which moves U + C bytes in and C + U bytes out per round trip."""
import torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
def timed(fn, reps=4):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
t = timed(lambda: d.copy_(h, non_blocking=True)); print(f"H2D 1 GiB: {t:7.2f} ms  {n / t / 1e6:6.1f} GB/s")
t = timed(lambda: h.copy_(d, non_blocking=True)); print(f"D2H 1 GiB: {t:7.2f} ms  {n / t / 1e6:6.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
t = timed(both); print(f"H2D + D2H concurrently, 1 GiB each: {t:7.2f} ms  {2 * n / t / 1e6:6.1f} GB/s aggregate")
