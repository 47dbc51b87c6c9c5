"""Throughput of the GENERAL inflate kernel (K3) on zlib-compressed synthetic tiles (device-resident)."""
import sys, time, zlib
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, ".")
import numpy as np, torch
import fdeflate_b200 as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lib = F.NativeLib(sys.argv[2]) if len(sys.argv) > 2 else None   # a variant build (A/B runs)
ctx = F.Context(0, lib)
tiles = F.synth_tiles_host(0, n, 256, 256, 1, lib)
raw = [t.tobytes() for t in tiles]
for level, name in ((6, "zlib-6"), (1, "zlib-1"), (0, "stored")):
    with ThreadPoolExecutor(16) as ex:
        comp = list(ex.map(lambda r: zlib.compress(r, level), raw))
    t0 = time.perf_counter()
    st, outs, _ = ctx.inflate_batch(comp, [len(r) for r in raw])
    dt = time.perf_counter() - t0
    assert (st == 0).all() and outs[0] == raw[0] and outs[-1] == raw[-1]
    # device-resident timing
    in_base, in_off, in_len = ctx._pack(comp)
    out_off = np.arange(n, dtype=np.uint64) * len(raw[0])
    out_cap = np.full(n, len(raw[0]), dtype=np.uint64)
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(in_base).to(dev)
    d_io, d_il = torch.from_numpy(in_off.astype(np.int64)).to(dev), torch.from_numpy(in_len.astype(np.int64)).to(dev)
    d_oo, d_oc = torch.from_numpy(out_off.astype(np.int64)).to(dev), torch.from_numpy(out_cap.astype(np.int64)).to(dev)
    d_out = torch.empty(n * len(raw[0]), dtype=torch.uint8, device=dev)
    d_ol = torch.zeros(n, dtype=torch.int64, device=dev); d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    def run():
        ctx.inflate_device(d_in.data_ptr(), d_io.data_ptr(), d_il.data_ptr(), d_out.data_ptr(), d_oo.data_ptr(), d_oc.data_ptr(),
                           d_ol.data_ptr(), 0, d_st.data_ptr(), n, 0, s)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    assert int(d_st.abs().sum()) == 0
    print(f"{sys.argv[2] if len(sys.argv) > 2 else 'default'} {name}: ratio {sum(map(len, comp)) / (n * len(raw[0])):.3f}  {n * len(raw[0]) / ms / 1e6:.1f} GB/s uncompressed ({ms:.2f} ms, {n} streams)")
