"""Experiment: kernels reading / writing pinned host memory directly (zero-copy over PCIe) vs device memory."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import fdeflate_b200 as F
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
TB = 262400
ctx = F.Context(0); ctx2 = F.Context(0)
dev = torch.device("cuda:0"); i64 = torch.int64
s = torch.cuda.current_stream().cuda_stream
tiles = torch.empty(n * TB, dtype=torch.uint8, device=dev)
ctx.synth_tiles_device(tiles.data_ptr(), 0, n, 256, 256, 2024, s)
bound = ctx.ultrafast_bound(TB)
t_off = torch.arange(n, dtype=i64, device=dev) * TB; t_len = torch.full((n,), TB, dtype=i64, device=dev)
c_off = torch.arange(n, dtype=i64, device=dev) * bound; c_cap = torch.full((n,), bound, dtype=i64, device=dev)
comp = torch.zeros(n * bound, dtype=torch.uint8, device=dev)
c_len = torch.zeros(n, dtype=i64, device=dev); c_st = torch.zeros(n, dtype=torch.int32, device=dev)
out = torch.empty(n * TB, dtype=torch.uint8, device=dev)
o_len = torch.zeros(n, dtype=i64, device=dev); o_st = torch.zeros(n, dtype=torch.int32, device=dev)
h_tiles = torch.empty(n * TB, dtype=torch.uint8, pin_memory=True)
h_comp = torch.zeros(n * bound, dtype=torch.uint8, pin_memory=True)
h_comp_in = torch.zeros(n * bound, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n * TB, dtype=torch.uint8, pin_memory=True)
def deflate(src, dst, c=ctx, st=s):
    c.deflate_ultrafast_device(src.data_ptr(), t_off.data_ptr(), t_len.data_ptr(), dst.data_ptr(), c_off.data_ptr(),
                               c_cap.data_ptr(), c_len.data_ptr(), c_st.data_ptr(), n, st)
def inflate(src, dst, c=ctx, st=s):
    c.inflate_device(src.data_ptr(), c_off.data_ptr(), c_len.data_ptr(), dst.data_ptr(), t_off.data_ptr(), t_len.data_ptr(),
                     o_len.data_ptr(), 0, o_st.data_ptr(), n, 0, st)
deflate(tiles, comp); torch.cuda.synchronize()
h_tiles.copy_(tiles); h_comp_in.copy_(comp); torch.cuda.synchronize()
cb = int(c_len.sum())
def timeit(f, reps=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
for name, f, nb in (("deflate dev->dev", lambda: deflate(tiles, comp), 0), ("deflate HOST->dev", lambda: deflate(h_tiles, comp), n*TB),
              ("deflate dev->HOST", lambda: deflate(tiles, h_comp), cb), ("deflate HOST->HOST", lambda: deflate(h_tiles, h_comp), n*TB),
              ("inflate dev->dev", lambda: inflate(comp, out), 0), ("inflate HOST->dev", lambda: inflate(h_comp_in, out), cb),
              ("inflate dev->HOST", lambda: inflate(comp, h_out), n*TB), ("inflate HOST->HOST", lambda: inflate(h_comp_in, h_out), n*TB)):
    dt = timeit(f)
    print(f"{name}: {dt*1e3:.2f} ms  {n*TB/dt/1e9:.1f} GB/s uncompressed" + (f"  (PCIe major direction {nb/dt/1e9:.1f} GB/s)" if nb else ""))
assert int(o_st.abs().sum()) == 0 and bool((h_out == h_tiles).all())
# both directions at once on two streams / contexts
s2 = torch.cuda.Stream()
def both():
    deflate(h_tiles, h_comp, ctx, s)
    inflate(h_comp_in, h_out, ctx2, s2.cuda_stream)
dt = timeit(both)
print(f"deflate+inflate HOST->HOST concurrently: {dt*1e3:.2f} ms  {2*n*TB/dt/1e9:.1f} GB/s uncompressed (each way {(n*TB+cb)/dt/1e9:.1f} GB/s)")
assert bool((h_out == h_tiles).all())
