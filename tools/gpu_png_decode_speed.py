"""End to end through fdb_png_decode_batch / fdb_png_encode_batch (host buffers, copies included; not pipelined):
4096 tiles of 256x256 RGBA as ultra-fast zlib streams -> pixels and back."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import fdeflate_b200 as F
from fdeflate_b200.api import _ptr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H, S, B = 256, 1024, 4
ctx = F.Context(0)
tiles = F.synth_tiles_host(0, n, 256, 256, 2024)
pin = lambda nbytes: torch.zeros(nbytes, dtype=torch.uint8, pin_memory=True).numpy()
raw_off = np.arange(n, dtype=np.uint64) * (H * S); h = np.full(n, H, np.uint32); s = np.full(n, S, np.uint32); b = np.full(n, B, np.uint32)
bound = ctx.ultrafast_bound(H * (1 + S))
z_off = np.arange(n, dtype=np.uint64) * bound; z_cap = np.full(n, bound, np.uint64)
raw = pin(n * H * S); z = pin(n * bound); raw2 = pin(n * H * S)
st, px = ctx.png_unfilter_batch([t.tobytes() for t in tiles[:64]], [(H, S, B)] * 64)
for i in range(n): raw[i * H * S:(i + 1) * H * S] = np.frombuffer(px[i % 64], dtype=np.uint8)
z_len = np.zeros(n, np.uint64); status = np.zeros(n, np.int32)
def encode():
    rc = ctx.lib.L.fdb_png_encode_batch(ctx._h, _ptr(raw), _ptr(raw_off), _ptr(h), _ptr(s), _ptr(b), 4, _ptr(z), _ptr(z_off), _ptr(z_cap), _ptr(z_len), _ptr(status), n)
    assert rc == 0 and (status == 0).all()
packed = {}
def decode():
    if not packed:  # the streams back to back, as they sit in PNG files
        off = np.zeros(n, np.uint64); off[1:] = np.cumsum((z_len[:-1] + np.uint64(15)) & ~np.uint64(15))
        buf = pin(int(off[-1] + z_len[-1]) + 16)
        for i in range(n): buf[int(off[i]):int(off[i]) + int(z_len[i])] = z[int(z_off[i]):int(z_off[i]) + int(z_len[i])]
        packed["off"], packed["buf"] = off, buf
    rc = ctx.lib.L.fdb_png_decode_batch(ctx._h, _ptr(packed["buf"]), _ptr(packed["off"]), _ptr(z_len), _ptr(raw2), _ptr(raw_off), _ptr(h), _ptr(s), _ptr(b), _ptr(status), n)
    assert rc == 0 and (status == 0).all()
for name, f in (("encode (filter Paeth + ultra-fast deflate)", encode), ("decode (inflate + unfilter)", decode)):
    f(); t0 = time.perf_counter()
    for _ in range(3): f()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {dt*1e3:.1f} ms per {n} tiles = {n*H*S/dt/1e9:.1f} GB/s of pixels (host buffers)")
assert (raw2 == raw).all()
# ---- whole PNG files through the library's own container walk (fdb_png_probe_batch + fdb_png_decode_files_batch)
import struct, binascii
def chunk(t, b): return struct.pack(">I", len(b)) + t + b + struct.pack(">I", binascii.crc32(t + b) & 0xffffffff)
ihdr = chunk(b"IHDR", struct.pack(">IIBBBBB", 256, 256, 8, 6, 0, 0, 0)); iend = chunk(b"IEND", b"")
sig = b"\x89PNG\r\n\x1a\n"
for label, piece in (("one IDAT chunk per file", 1 << 30), ("8 KiB IDAT chunks (gather)", 8192)):
    files = []
    for i in range(n):
        zi = z[int(z_off[i]):int(z_off[i]) + int(z_len[i])].tobytes()
        files.append(sig + ihdr + b"".join(chunk(b"IDAT", zi[k:k + piece]) for k in range(0, len(zi), piece)) + iend)
    lens = np.array([len(f) for f in files], np.uint64); offs = np.zeros(n, np.uint64); offs[1:] = np.cumsum(lens[:-1])
    fbuf = pin(int(lens.sum()) + 16)
    for i, f in enumerate(files): fbuf[int(offs[i]):int(offs[i]) + len(f)] = np.frombuffer(f, np.uint8)
    w_, h_, d_, c_, s_ = (np.zeros(n, np.uint32) for _ in range(5))
    t0 = time.perf_counter()
    assert ctx.lib.L.fdb_png_probe_batch(_ptr(fbuf), _ptr(offs), _ptr(lens), _ptr(w_), _ptr(h_), _ptr(d_), _ptr(c_), _ptr(s_), _ptr(status), n) == 0
    t_probe = time.perf_counter() - t0
    assert (status == 0).all() and (h_ == 256).all() and (s_ == 1024).all()
    raw2[:] = 0
    def decode_files():
        rc = ctx.lib.L.fdb_png_decode_files_batch(ctx._h, _ptr(fbuf), _ptr(offs), _ptr(lens), _ptr(raw2), _ptr(raw_off), _ptr(status), n)
        assert rc == 0 and (status == 0).all()
    decode_files(); t0 = time.perf_counter()
    for _ in range(3): decode_files()
    dt = (time.perf_counter() - t0) / 3
    assert (raw2 == raw).all()
    print(f"decode PNG files, {label}: probe {t_probe*1e3:.1f} ms, decode {dt*1e3:.1f} ms per {n} files = {n*H*S/dt/1e9:.1f} GB/s of pixels")
