"""Throughput of the stored ("level 0") deflate kernel and of inflating its output (device-resident)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import fdeflate_b200 as F
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
TB = 262400
ctx = F.Context(0); dev = torch.device("cuda:0"); i64 = torch.int64
s = torch.cuda.current_stream().cuda_stream
tiles = torch.empty(n * TB, dtype=torch.uint8, device=dev)
ctx.synth_tiles_device(tiles.data_ptr(), 0, n, 256, 256, 2024, s)
bound = ctx.stored_bound(TB)
t_off = torch.arange(n, dtype=i64, device=dev) * TB; t_len = torch.full((n,), TB, dtype=i64, device=dev)
c_off = torch.arange(n, dtype=i64, device=dev) * bound; c_cap = torch.full((n,), bound, dtype=i64, device=dev)
comp = torch.zeros(n * bound, dtype=torch.uint8, device=dev)
c_len = torch.zeros(n, dtype=i64, device=dev); c_st = torch.zeros(n, dtype=torch.int32, device=dev)
out = torch.empty(n * TB, dtype=torch.uint8, device=dev)
o_len = torch.zeros(n, dtype=i64, device=dev); o_st = torch.zeros(n, dtype=torch.int32, device=dev)
def deflate():
    ctx.deflate_stored_device(tiles.data_ptr(), t_off.data_ptr(), t_len.data_ptr(), comp.data_ptr(), c_off.data_ptr(),
                              c_cap.data_ptr(), c_len.data_ptr(), c_st.data_ptr(), n, s)
def inflate():
    ctx.inflate_device(comp.data_ptr(), c_off.data_ptr(), c_len.data_ptr(), out.data_ptr(), t_off.data_ptr(), t_len.data_ptr(),
                       o_len.data_ptr(), 0, o_st.data_ptr(), n, 0, s)
def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms_d = timed(deflate); ms_i = timed(inflate)
assert int(c_st.abs().sum()) == 0 and int(o_st.abs().sum()) == 0 and torch.equal(out, tiles)
import zlib, json
h = comp[:int(c_len[0])].cpu().numpy().tobytes()
assert zlib.decompress(h) == tiles[:TB].cpu().numpy().tobytes()
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6553.3
alg = n * TB + int(c_len.sum())
print(f"stored deflate: {ms_d:.3f} ms = {n*TB/ms_d/1e6:.1f} GB/s uncompressed, algorithmic {alg/ms_d/1e6:.1f} GB/s = {alg/ms_d/1e6/peak:.3f} of HBM peak {peak}")
print(f"inflate of stored streams (general kernel): {ms_i:.3f} ms = {n*TB/ms_i/1e6:.1f} GB/s uncompressed, algorithmic {alg/ms_i/1e6:.1f} GB/s = {alg/ms_i/1e6/peak:.3f} of HBM peak")
