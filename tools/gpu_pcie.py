"""PCIe ceiling on the GPU box: pinned H2D, D2H and both at once (what bounds the end-to-end number)."""
import torch, time
n = 1 << 30
h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5, chunk=None):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if chunk is None:
            if h2d:
                with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
        else:
            for o in range(0, n, chunk):
                if h2d:
                    with torch.cuda.stream(s1): d_a[o:o+chunk].copy_(h_a[o:o+chunk], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2): h_b[o:o+chunk].copy_(d_b[o:o+chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    return n / dt / 1e9
run(True, True, 1)
print("H2D alone  %.1f GB/s" % run(True, False))
print("D2H alone  %.1f GB/s" % run(False, True))
print("both       %.1f GB/s each way" % run(True, True))
print("both, 64 MiB chunks %.1f GB/s each way" % run(True, True, chunk=64 << 20))
print("both, 8 MiB chunks  %.1f GB/s each way" % run(True, True, chunk=8 << 20))
