"""configs[4] shape (8192 streams of 64 KiB..16 MiB per GPU, device-resident): inflate / deflate time as a function of
the size above which a stream is cut into spans / segments.   python tools/gpu_sweep_threshold.py [n_streams]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import fdeflate_b200 as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
W = 1024; ROW = 1 + 4 * W
rng = np.random.default_rng(5)
sizes = np.exp(rng.uniform(np.log(64 << 10), np.log(16 << 20), n))
heights = np.maximum(1, (sizes / ROW).astype(np.int64))
lens = heights * ROW
ctx = F.Context(0); dev = torch.device("cuda:0"); i64 = torch.int64
s = torch.cuda.current_stream().cuda_stream
offs = np.zeros(n, dtype=np.int64); offs[1:] = np.cumsum((lens[:-1] + 15) & ~15)
total = int(offs[-1] + lens[-1])
raw = torch.empty(total + 16, dtype=torch.uint8, device=dev)
for i in range(n):
    ctx.synth_tiles_device(raw.data_ptr() + int(offs[i]), 1000 + i, 1, W, int(heights[i]), 5, s)
bounds = np.array([ctx.ultrafast_bound(int(l)) for l in lens], dtype=np.int64)
coffs = np.zeros(n, dtype=np.int64); coffs[1:] = np.cumsum(bounds[:-1])
comp = torch.empty(int(coffs[-1] + bounds[-1]), dtype=torch.uint8, device=dev)
out = torch.empty(total + 16, dtype=torch.uint8, device=dev)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d_off, d_len, d_coff, d_ccap = T(offs), T(lens), T(coffs), T(bounds)
c_len = torch.zeros(n, dtype=i64, device=dev); c_st = torch.zeros(n, dtype=torch.int32, device=dev)
o_len = torch.zeros(n, dtype=i64, device=dev); o_st = torch.zeros(n, dtype=torch.int32, device=dev)
ctx.set_split_large(True)
def deflate():
    ctx.deflate_ultrafast_device(raw.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), comp.data_ptr(), d_coff.data_ptr(),
                                 d_ccap.data_ptr(), c_len.data_ptr(), c_st.data_ptr(), n, s)
def inflate():
    ctx.inflate_device(comp.data_ptr(), d_coff.data_ptr(), c_len.data_ptr(), out.data_ptr(), d_off.data_ptr(), d_len.data_ptr(),
                       o_len.data_ptr(), 0, o_st.data_ptr(), n, F.FLAG_SPLIT_LARGE, s)
def timed(f, reps=2):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
deflate(); torch.cuda.synchronize()
unc = int(lens.sum())
print(f"{n} streams, {unc/1e9:.2f} GB, ratio {int(c_len.sum())/unc:.3f}")
for thr in (256 << 10, 512 << 10, 1 << 20, 2 << 20, 4 << 20, 8 << 20, 1 << 40):
    ctx.set_split_threshold(thr, 0)
    ms = timed(inflate)
    assert int(o_st.abs().sum()) == 0 and torch.equal(o_len, d_len)
    print(f"inflate: cut compressed streams >= {thr/2**20:9.2f} MiB: {ms:8.2f} ms = {unc/ms/1e6:7.1f} GB/s  (spans {ctx.last_split_spans(s)})")
ctx.set_split_threshold(0, 0)
for thr in (256 << 10, 1 << 20, 4 << 20, 1 << 40):
    ctx.set_split_threshold(0, thr)
    ms = timed(deflate)
    assert int(c_st.abs().sum()) == 0
    print(f"deflate: cut inputs >= {thr/2**20:9.2f} MiB (batch policy on top): {ms:8.2f} ms = {unc/ms/1e6:7.1f} GB/s")
