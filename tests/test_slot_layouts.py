"""Slot layouts the header allows but no packer produces (round-1 review findings): slots in any order, evenly
spaced slots whose last one is short, PNG encode jobs that are rejected before they reach the device.  Every call
goes through the C ABI with guard bytes around the caller's buffers; nothing outside a slot's hull may change.
Runs on the emulator build on CPU and on the CUDA build with -m gpu."""
import ctypes as C
import random
import zlib

import numpy as np
import pytest

import cases

GUARD = 4096


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _guarded(n):
    buf = np.full(n + 2 * GUARD, 0xA5, dtype=np.uint8)
    return buf, buf[GUARD:GUARD + n]


def _guards_intact(buf, n):
    return bool((buf[:GUARD] == 0xA5).all() and (buf[GUARD + n:] == 0xA5).all())


def _inflate(ctx, streams, in_off, out_off, caps):
    n = len(streams)
    in_off = np.array(in_off, dtype=np.uint64)
    in_len = np.array([len(s) for s in streams], dtype=np.uint64)
    in_size = int(max(o + l for o, l in zip(in_off, in_len)))
    in_base = np.zeros(in_size, dtype=np.uint8)
    for s, o in zip(streams, in_off):
        in_base[int(o): int(o) + len(s)] = np.frombuffer(s, dtype=np.uint8)
    out_off = np.array(out_off, dtype=np.uint64)
    caps = np.array(caps, dtype=np.uint64)
    out_size = int(max(o + c for o, c in zip(out_off, caps)))
    buf, out = _guarded(out_size)
    out_len = np.zeros(n, dtype=np.uint64)
    cons = np.zeros(n, dtype=np.uint64)
    st = np.full(n, -7, dtype=np.int32)
    rc = ctx.lib.L.fdb_inflate_batch(ctx._h, _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out), _ptr(out_off), _ptr(caps),
                                     _ptr(out_len), _ptr(cons), _ptr(st), n, 0)
    assert rc == 0 and _guards_intact(buf, out_size)
    return st, [out[int(o): int(o) + int(l)].tobytes() for o, l in zip(out_off, out_len)]


def _suite(ctx, oracle):
    rng = random.Random(3)
    datas = [cases.sparse_bytes(rng, n) for n in (500, 3000, 900)]
    zs = [zlib.compress(d, 6) for d in datas[:2]] + [oracle.compress_ultra_fast(datas[2])]
    caps = [len(d) for d in datas]
    # input slots permuted (0, 2000, 1000 in the review), output slots permuted, both descending
    for in_off, out_off in (([0, 8000, 4000], [0, 4096, 8192]), ([0, 4000, 8000], [0, 8192, 4096]),
                            ([8000, 4000, 0], [8192, 4096, 0]), ([16, 9000, 5003], [7, 9001, 4100])):
        st, outs = _inflate(ctx, zs, in_off, out_off, caps)
        assert list(st) == [0, 0, 0] and outs == datas, (in_off, out_off)
    # ten evenly spaced slots, the last one short: the 2-D copy must not run past the caller's buffer
    small = [bytes(rng.getrandbits(8) for _ in range(600))] * 9 + [b"tail" * 10]
    zs = [zlib.compress(d, 1) for d in small]
    st, outs = _inflate(ctx, zs, [1024 * i for i in range(10)], [1024 * i for i in range(10)], [1000] * 9 + [40])
    assert list(st) == [0] * 10 and outs == small
    # ... and on the way up: inputs evenly spaced, last input short, first inputs long
    big = [cases.sparse_bytes(rng, 60000)] * 3 + [b"x"]
    zs = [zlib.compress(d, 0) for d in big]
    stride = (len(zs[0]) + 63) // 64 * 64
    st, outs = _inflate(ctx, zs, [stride * i for i in range(4)], [65536 * i for i in range(4)], [60000] * 3 + [1])
    assert list(st) == [0] * 4 and outs == big
    # deflate with a short last slot: OutputBufferTooSmall for it, guards intact
    n = 6
    inputs = [cases.sparse_bytes(rng, 5000) for _ in range(n)]
    in_base, in_off, in_len = ctx._pack(inputs)
    bound = int(ctx.lib.L.fdb_deflate_ultrafast_bound(5000))
    caps = np.array([bound] * (n - 1) + [48], dtype=np.uint64)
    out_off = np.arange(n, dtype=np.uint64) * np.uint64(bound)
    out_size = int(out_off[-1]) + 48
    buf, out = _guarded(out_size)
    out_len = np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=np.int32)
    assert ctx.lib.L.fdb_deflate_ultrafast_batch(ctx._h, _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out), _ptr(out_off),
                                                  _ptr(caps), _ptr(out_len), _ptr(st), n) == 0
    assert _guards_intact(buf, out_size) and list(st) == [0] * (n - 1) + [18]
    for i in range(n - 1):
        assert out[int(out_off[i]): int(out_off[i]) + int(out_len[i])].tobytes() == oracle.compress_ultra_fast(inputs[i])


def _encode_files_suite(ctx):
    """images that never reach the device (bad geometry, slot too small) around good ones: their slots stay
    untouched, the good files are intact (review: they used to share one device slot and write past a zero-size one)"""
    from fdeflate_b200 import png

    rng = np.random.default_rng(1)
    img = rng.integers(0, 255, size=(40, 30, 4), dtype=np.uint8)
    raws = [img.tobytes()] * 5
    raw_base, raw_off, _ = ctx._pack(raws)
    w = np.array([30] * 5, dtype=np.uint32)
    h = np.array([40] * 5, dtype=np.uint32)
    depth = np.array([1, 8, 8, 8, 1], dtype=np.uint32)   # bit depth 1: not encoded here
    color = np.array([6] * 5, dtype=np.uint32)
    bound = int(ctx.lib.L.fdb_png_file_bound(30, 40, 8, 6))
    assert ctx.lib.L.fdb_png_file_bound(30, 40, 1, 6) == 0
    caps = np.array([0, bound, 20, bound, 0], dtype=np.uint64)   # slot 2: too small for the framing
    f_off = np.array([0, 0, bound, bound + 32, 2 * bound + 32], dtype=np.uint64)
    total = 2 * bound + 32
    buf, files = _guarded(total)
    files[:] = 0x5A
    f_len = np.zeros(5, dtype=np.uint64)
    st = np.zeros(5, dtype=np.int32)
    rc = ctx.lib.L.fdb_png_encode_files_batch(ctx._h, _ptr(raw_base), _ptr(raw_off), _ptr(w), _ptr(h), _ptr(depth), _ptr(color), 4,
                                              _ptr(files), _ptr(f_off), _ptr(caps), _ptr(f_len), _ptr(st), 5)
    assert rc == 0 and _guards_intact(buf, total)
    assert list(st) == [20, 0, 18, 0, 20] and list(f_len[[0, 2, 4]]) == [0, 0, 0]
    assert (files[bound: bound + 32] == 0x5A).all()  # the too-small slot and the padding behind it
    for i in (1, 3):
        f = files[int(f_off[i]): int(f_off[i]) + int(f_len[i])].tobytes()
        assert np.array_equal(png.decode_batch([f], ctx)[0], img)


@pytest.mark.emul
def test_slot_layouts_on_emulator(emul_ctx, oracle):
    _suite(emul_ctx, oracle)
    _encode_files_suite(emul_ctx)


@pytest.mark.gpu
def test_slot_layouts_on_gpu(gpu_ctx, oracle):
    _suite(gpu_ctx, oracle)
    _encode_files_suite(gpu_ctx)
