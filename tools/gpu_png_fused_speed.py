"""PNG encode on the device, 4096 raw 256x256 RGBA tiles: filter kernel + ultra-fast deflate (two kernels, the filtered
image goes through device memory) against fdb_png_encode_batch_device (the filter inside the encoder).  Outputs compared."""
import sys, json, os
sys.path.insert(0, ".")
import torch
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
f_len = torch.full((n,), FILT, dtype=i64, device=dev)
h = torch.full((n,), H, dtype=torch.int32, device=dev); st = torch.full((n,), S, dtype=torch.int32, device=dev)
b = torch.full((n,), B, dtype=torch.int32, device=dev)
raw = torch.empty(n * RAW + 64, dtype=torch.uint8, device=dev); filt = torch.empty(n * FILT, dtype=torch.uint8, device=dev)
status = torch.zeros(n, dtype=torch.int32, device=dev); fstatus = torch.zeros(n, dtype=torch.int32, device=dev)
ctx.png_unfilter_device(tiles.data_ptr(), f_off.data_ptr(), raw.data_ptr(), r_off.data_ptr(), h.data_ptr(), st.data_ptr(), b.data_ptr(), status.data_ptr(), n, s)
bound = ctx.ultrafast_bound(FILT)
c_off = torch.arange(n, dtype=i64, device=dev) * bound; c_cap = torch.full((n,), bound, dtype=i64, device=dev)
comp_a = torch.zeros(n * bound, dtype=torch.uint8, device=dev); comp_b = torch.zeros(n * bound, dtype=torch.uint8, device=dev)
len_a = torch.zeros(n, dtype=i64, device=dev); len_b = torch.zeros(n, dtype=i64, device=dev)
def two(mode):
    ctx.png_filter_device(raw.data_ptr(), r_off.data_ptr(), filt.data_ptr(), f_off.data_ptr(), h.data_ptr(), st.data_ptr(), b.data_ptr(), mode, status.data_ptr(), n, s)
    ctx.deflate_ultrafast_device(filt.data_ptr(), f_off.data_ptr(), f_len.data_ptr(), comp_a.data_ptr(), c_off.data_ptr(), c_cap.data_ptr(), len_a.data_ptr(), status.data_ptr(), n, s)
def fused(mode):
    ctx.png_encode_device(raw.data_ptr(), r_off.data_ptr(), h.data_ptr(), st.data_ptr(), b.data_ptr(), mode, comp_b.data_ptr(), c_off.data_ptr(), c_cap.data_ptr(), len_b.data_ptr(), fstatus.data_ptr(), status.data_ptr(), n, s)
def timed(f, reps=5):
    f(); f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for mode, name in ((4, "Paeth"), (1, "Sub"), (2, "Up"), (3, "Average"), (0, "None")):
    comp_a.zero_(); comp_b.zero_()
    t2 = timed(lambda: two(mode)); t1 = timed(lambda: fused(mode))
    ok = bool(torch.equal(len_a, len_b)) and bool(torch.equal(comp_a, comp_b)) and int(fstatus.abs().sum()) == 0 and int(status.abs().sum()) == 0
    print(f"mode {mode} ({name}): filter + deflate {t2:.3f} ms = {n*RAW/t2/1e6:.0f} GB/s of pixels | fused {t1:.3f} ms = {n*RAW/t1/1e6:.0f} GB/s | same bytes: {ok} | ratio {float(len_a.sum())/(n*RAW):.3f}", flush=True)
