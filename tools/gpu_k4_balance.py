"""How much of the fast-path inflate time is the tail of the longest streams?  Same kernel on (a) the bench's 4096
different tiles and (b) 4096 copies of one median-sized tile (every warp has the same amount of work)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import fdeflate_b200 as F
n, TB = 4096, 262400
ctx = F.Context(0); dev = torch.device("cuda:0"); i64 = torch.int64
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
ctx.deflate_ultrafast_device(tiles.data_ptr(), t_off.data_ptr(), t_len.data_ptr(), comp.data_ptr(), c_off.data_ptr(), c_cap.data_ptr(), c_len.data_ptr(), c_st.data_ptr(), n, s)
torch.cuda.synchronize()
def inflate(off, ln):
    ctx.inflate_device(comp.data_ptr(), off.data_ptr(), ln.data_ptr(), out.data_ptr(), t_off.data_ptr(), t_len.data_ptr(), o_len.data_ptr(), 0, o_st.data_ptr(), n, 0, s)
def timed(f, reps=10):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timed(lambda: inflate(c_off, c_len))
cl = c_len.cpu().numpy()
print(f"4096 different tiles: {ms:.3f} ms = {n*TB/ms/1e6:.1f} GB/s  (compressed size min/mean/max = {cl.min()}/{cl.mean():.0f}/{cl.max()})")
med = int(np.argsort(cl)[n // 2])
off1 = torch.full((n,), int(med * bound), dtype=i64, device=dev); len1 = torch.full((n,), int(cl[med]), dtype=i64, device=dev)
ms1 = timed(lambda: inflate(off1, len1))
print(f"4096 copies of the median tile ({cl[med]} B): {ms1:.3f} ms = {n*TB/ms1/1e6:.1f} GB/s")
srt = torch.from_numpy(np.argsort(-cl).astype(np.int64)).to(dev)
ms2 = timed(lambda: inflate(c_off[srt].contiguous(), c_len[srt].contiguous()))
print(f"4096 tiles, longest first: {ms2:.3f} ms")
