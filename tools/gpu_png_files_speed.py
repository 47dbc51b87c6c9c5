"""fdb_png_encode_files_batch / fdb_png_decode_files_batch on 4096 images of 256x256 RGBA (pinned host buffers)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import fdeflate_b200 as F
from fdeflate_b200.api import _ptr
n, H, S = 4096, 256, 1024
ctx = F.Context(0)
pin = lambda nb: torch.zeros(nb, dtype=torch.uint8, pin_memory=True).numpy()
tiles = F.synth_tiles_host(0, 64, 256, 256, 2024)
st, px = ctx.png_unfilter_batch([t.tobytes() for t in tiles], [(H, S, 4)] * 64)
raw = pin(n * H * S); raw2 = pin(n * H * S)
for i in range(n): raw[i*H*S:(i+1)*H*S] = np.frombuffer(px[i % 64], np.uint8)
raw_off = np.arange(n, dtype=np.uint64) * (H * S)
w = np.full(n, 256, np.uint32); h = np.full(n, 256, np.uint32); d = np.full(n, 8, np.uint32); c = np.full(n, 6, np.uint32)
cap = (int(ctx.lib.L.fdb_png_file_bound(256, 256, 8, 6)) + 15) // 16 * 16
f_off = np.arange(n, dtype=np.uint64) * cap; f_cap = np.full(n, cap, np.uint64); f_len = np.zeros(n, np.uint64); status = np.zeros(n, np.int32)
files = pin(n * cap)
def enc():
    rc = ctx.lib.L.fdb_png_encode_files_batch(ctx._h, _ptr(raw), _ptr(raw_off), _ptr(w), _ptr(h), _ptr(d), _ptr(c), 4, _ptr(files), _ptr(f_off), _ptr(f_cap), _ptr(f_len), _ptr(status), n)
    assert rc == 0 and (status == 0).all()
def dec():
    rc = ctx.lib.L.fdb_png_decode_files_batch(ctx._h, _ptr(files), _ptr(f_off), _ptr(f_len), _ptr(raw2), _ptr(raw_off), _ptr(status), n)
    assert rc == 0 and (status == 0).all()
for name, f in (("encode PNG files (Paeth, ultra-fast deflate)", enc), ("decode them again", dec)):
    f(); t0 = time.perf_counter()
    for _ in range(3): f()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {dt*1e3:.1f} ms per {n} images = {n*H*S/dt/1e9:.1f} GB/s of pixels")
assert (raw2 == raw).all()
print(f"mean file {f_len.mean():.0f} B")
