"""Developer tool: PCIe copy bandwidth of the box (pinned host memory), one direction at a time and both at once -- the ceiling
of bench.py's e2e line (2 x 8.6 GB per step at 512^3)."""
import torch
n = 1 << 30          # 8 GiB of doubles
h_in = torch.empty(n, dtype=torch.float64, pin_memory=True)
h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        for s in (s1, s2):
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
gb = n * 8 / 1e9
t = timed(h2d); print("H2D alone: %.1f GB/s" % (gb / t * 1e3))
t = timed(d2h); print("D2H alone: %.1f GB/s" % (gb / t * 1e3))
t = timed(lambda: (h2d(), d2h())); print("H2D + D2H at once: %.1f GB/s each way (%.1f ms for 2 x %.1f GB)" % (gb / t * 1e3, t, gb))
