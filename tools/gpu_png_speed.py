"""Throughput of the PNG row-filter kernels on 4096 tiles of 256x256 RGBA (device-resident) and of the fused
inflate -> unfilter step."""
import sys, json, os
sys.path.insert(0, ".")
import numpy as np, torch
import fdeflate_b200 as F
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H, S, B = 256, 1024, 4
FILT, RAW = H * (1 + S), H * S
ctx = F.Context(0); dev = torch.device("cuda:0"); i64 = torch.int64
s = torch.cuda.current_stream().cuda_stream
tiles = torch.empty(n * FILT, dtype=torch.uint8, device=dev)
ctx.synth_tiles_device(tiles.data_ptr(), 0, n, 256, 256, 2024, s)
f_off = torch.arange(n, dtype=i64, device=dev) * FILT
r_off = torch.arange(n, dtype=i64, device=dev) * RAW
h = torch.full((n,), H, dtype=torch.int32, device=dev); st = torch.full((n,), S, dtype=torch.int32, device=dev)
b = torch.full((n,), B, dtype=torch.int32, device=dev)
raw = torch.empty(n * RAW, dtype=torch.uint8, device=dev); refilt = torch.empty(n * FILT, dtype=torch.uint8, device=dev)
status = torch.zeros(n, dtype=torch.int32, device=dev)
def unfilter(): ctx.png_unfilter_device(tiles.data_ptr(), f_off.data_ptr(), raw.data_ptr(), r_off.data_ptr(), h.data_ptr(), st.data_ptr(), b.data_ptr(), status.data_ptr(), n, s)
def filt(mode): ctx.png_filter_device(raw.data_ptr(), r_off.data_ptr(), refilt.data_ptr(), f_off.data_ptr(), h.data_ptr(), st.data_ptr(), b.data_ptr(), mode, status.data_ptr(), n, s)
def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6553.3
alg = n * (FILT + RAW)
ms = timed(unfilter); assert int(status.abs().sum()) == 0
print(f"unfilter: {ms:.3f} ms = {n*RAW/ms/1e6:.1f} GB/s of pixels, algorithmic {alg/ms/1e6:.1f} GB/s = {alg/ms/1e6/peak:.3f} of HBM peak")
# the generator's filter choice is Sub on row 0 and Paeth below: re-filtering with Paeth must give the tiles back except row 0
for mode, name in ((4, "Paeth"), (1, "Sub"), (5, "adaptive")):
    ms = timed(lambda: filt(mode)); assert int(status.abs().sum()) == 0
    print(f"filter {name}: {ms:.3f} ms = {n*RAW/ms/1e6:.1f} GB/s of pixels, algorithmic {alg/ms/1e6:.1f} GB/s = {alg/ms/1e6/peak:.3f} of HBM peak")
filt(4); torch.cuda.synchronize()
a = tiles.view(n, H, 1 + S); c = refilt.view(n, H, 1 + S)
assert torch.equal(a[:, 1:], c[:, 1:]), "Paeth rows differ from the generator's"
