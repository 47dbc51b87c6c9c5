"""Unfilter throughput for 1-, 2-, 3- and 4-byte pixels (random filtered rows with all five filter types)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import fdeflate_b200 as F
ctx = F.Context(0); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
n, H, Wpx = 2048, 256, 512
for bpp in (1, 2, 3, 4, 6):
    S = Wpx * bpp; FILT = H * (1 + S); RAW = H * S
    filt = torch.randint(0, 256, (n, H, 1 + S), dtype=torch.uint8, device=dev)
    filt[:, :, 0] = torch.randint(0, 5, (n, H), dtype=torch.uint8, device=dev)
    filt = filt.reshape(-1).contiguous()
    f_off = torch.arange(n, dtype=torch.int64, device=dev) * FILT; r_off = torch.arange(n, dtype=torch.int64, device=dev) * RAW
    h = torch.full((n,), H, dtype=torch.int32, device=dev); st = torch.full((n,), S, dtype=torch.int32, device=dev); b = torch.full((n,), bpp, dtype=torch.int32, device=dev)
    raw = torch.empty(n * RAW, dtype=torch.uint8, device=dev); status = torch.zeros(n, dtype=torch.int32, device=dev)
    f = lambda: ctx.png_unfilter_device(filt.data_ptr(), f_off.data_ptr(), raw.data_ptr(), r_off.data_ptr(), h.data_ptr(), st.data_ptr(), b.data_ptr(), status.data_ptr(), n, s)
    f(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [f() for _ in range(3)]; e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 3
    assert int(status.abs().sum()) == 0
    print(f"bpp {bpp}: {ms:.3f} ms for {n*RAW/1e9:.2f} GB of pixels = {n*RAW/ms/1e6:.1f} GB/s")
