"""BASELINE configs[4] shape at a reduced stream count: streams of 64 KiB .. 16 MiB (log-uniform, seeded),
ultra-fast deflate then inflate, device-resident, in natural (random) order and longest-first.
usage: python tools/gpu_sweep.py [n_streams] [max_MiB]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import fdeflate_b200 as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
max_mib = float(sys.argv[2]) if len(sys.argv) > 2 else 16
W = 1024; ROW = 1 + 4 * W
rng = np.random.default_rng(5)
sizes = np.exp(rng.uniform(np.log(64 << 10), np.log(max_mib * (1 << 20)), n))
heights = np.maximum(1, (sizes / ROW).astype(np.int64))
ctx = F.Context(0); dev = torch.device("cuda:0"); i64 = torch.int64
s = torch.cuda.current_stream().cuda_stream

def run(order_name, heights):
    lens = heights * ROW
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
    def deflate():
        ctx.deflate_ultrafast_device(raw.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), comp.data_ptr(), d_coff.data_ptr(),
                                     d_ccap.data_ptr(), c_len.data_ptr(), c_st.data_ptr(), n, s)
    def inflate(flags=F.FLAG_SPLIT_LARGE):
        ctx.inflate_device(comp.data_ptr(), d_coff.data_ptr(), c_len.data_ptr(), out.data_ptr(), d_off.data_ptr(), d_len.data_ptr(),
                           o_len.data_ptr(), 0, o_st.data_ptr(), n, flags, s)
    def timed(f, reps=3):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ms_d1 = timed(deflate)
    ref_len = c_len.clone(); ref_comp = comp.clone()
    ctx.set_split_large(True); comp.zero_()
    ms_d = timed(deflate)
    ctx.set_split_large(False)
    assert int(c_st.abs().sum()) == 0 and torch.equal(c_len, ref_len)
    for i in range(0, n, max(1, n // 64)):  # (bytes past a stream's length are unspecified)
        a = int(coffs[i]); assert torch.equal(comp[a:a + int(c_len[i])], ref_comp[a:a + int(c_len[i])]), "segment path differs"
    del ref_comp
    ms_i1 = timed(lambda: inflate(0))
    assert int(o_st.abs().sum()) == 0
    out.zero_()
    ms_i = timed(inflate)
    spans = ctx.last_split_spans(s)
    assert int(c_st.abs().sum()) == 0 and int(o_st.abs().sum()) == 0 and ctx.last_general_count(s) == 0
    assert torch.equal(o_len, d_len)
    for i in range(0, n, max(1, n // 16)):
        a = int(offs[i]); assert torch.equal(out[a:a + int(lens[i])], raw[a:a + int(lens[i])])
    unc = int(lens.sum())
    print(f"{order_name:14s} {n} streams, {unc/1e9:.2f} GB, ratio {int(c_len.sum())/unc:.3f}: deflate one warp/stream {ms_d1:8.2f} ms = {unc/ms_d1/1e6:6.1f} GB/s, segment by segment {ms_d:8.2f} ms = {unc/ms_d/1e6:7.1f} GB/s   "
          f"inflate one warp/stream {ms_i1:8.2f} ms = {unc/ms_i1/1e6:6.1f} GB/s, span by span ({spans} spans) {ms_i:8.2f} ms = {unc/ms_i/1e6:7.1f} GB/s   (largest stream {lens.max()/1e6:.1f} MB)")
    del raw, comp, out
    torch.cuda.empty_cache()

run("natural order", heights)
run("longest first", np.sort(heights)[::-1].copy())
