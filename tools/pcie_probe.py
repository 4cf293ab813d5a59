"""Raw pinned-memory H2D / D2H bandwidth of this box (alone and full duplex): the ceiling of bench.py's e2e number."""
import torch, time
n = 268 * 721 * 1440
h_in = torch.empty(n).pin_memory(); h_out = torch.empty(n).pin_memory()
d_in = torch.empty(n, device="cuda"); d_out = torch.randn(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return n * 4 * reps / (time.perf_counter() - t) / 1e9
run(1, 1, 1)
print(f"H2D alone {run(1,0):.1f} GB/s  D2H alone {run(0,1):.1f} GB/s  duplex (each dir) {run(1,1):.1f} GB/s")
