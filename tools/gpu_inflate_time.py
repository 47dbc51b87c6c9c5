"""Device-resident time of the fast-path inflate kernel on the bench tiles, checksum ignored, no verification
(for A/B experiments that change the result)."""
import sys
sys.path.insert(0, ".")
import torch
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
out = torch.empty(n * TB, dtype=torch.uint8, device=dev); o_len = torch.zeros(n, dtype=i64, device=dev); o_st = torch.zeros(n, dtype=torch.int32, device=dev)
ctx.deflate_ultrafast_device(tiles.data_ptr(), t_off.data_ptr(), t_len.data_ptr(), comp.data_ptr(), c_off.data_ptr(), c_cap.data_ptr(), c_len.data_ptr(), c_st.data_ptr(), n, s)
f = lambda: ctx.inflate_device(comp.data_ptr(), c_off.data_ptr(), c_len.data_ptr(), out.data_ptr(), t_off.data_ptr(), t_len.data_ptr(), o_len.data_ptr(), 0, o_st.data_ptr(), n, 1, s)
f(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); [f() for _ in range(20)]; e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"{sys.argv[1] if len(sys.argv) > 1 else ''} inflate {ms:.4f} ms = {n*TB/ms/1e6:.1f} GB/s  (general-kernel streams {ctx.last_general_count(s)}, statuses != 0: {int((o_st != 0).sum())})")
