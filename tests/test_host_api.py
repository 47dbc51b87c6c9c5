"""The host-side mirror of the reference's public API (fdeflate_b200/api.py) -- same names, argument
meaning and error behaviour as src/lib.rs:29-36 -- exercised the way the reference's own unit tests
exercise the crate.  Runs on the emulator build on CPU and on the CUDA build with -m gpu."""
import io
import random
import zlib

import numpy as np
import pytest

import cases
import fdeflate_b200 as F


def _suite(ctx, oracle):
    # decompress.rs:1235-1259 it_works / constant / random (round trips through an independent encoder)
    rng = random.Random(1)
    for data in [b"Hello world!", bytes(50), bytes([5]) * 2048, bytes([128]) * 2048, bytes([254]) * 2048,
                 bytes(rng.randrange(5) for _ in range(50000))]:
        assert F.decompress_to_vec(zlib.compress(data, 3), ctx) == data
        # ultrafast.rs:201-224 round trips, plus byte parity with the oracle
        z = F.compress_to_vec_ultra_fast(data, ctx)
        assert z == oracle.compress_ultra_fast(data) and zlib.decompress(z) == data
        assert F.decompress_to_vec(z, ctx) == data
        s = F.compress_to_vec_stored(data, ctx)
        assert s == oracle.compress_stored(data) and F.decompress_to_vec(s, ctx) == data

    # decompress.rs:1261-1280 ignore_adler32
    z = bytearray(zlib.compress(b"Hello world!", 1))
    z[-1] = (z[-1] + 1) & 0xFF
    with pytest.raises(F.DecompressionError) as e:
        F.decompress_to_vec(bytes(z), ctx)
    assert e.value.kind == "WrongChecksum" and e.value == F.DecompressionError(15)
    d = F.Decompressor(ctx)
    d.ignore_adler32()
    out = np.zeros(1024, np.uint8)
    consumed, produced = d.read(bytes(z), out, 0)
    assert out[:produced].tobytes() == b"Hello world!" and d.is_done()

    # decompress.rs:1282-1307 checksum_after_eof: the last byte arrives in a second read
    z = zlib.compress(b"Hello world!", 1)
    d = F.Decompressor(ctx)
    out = np.zeros(1024, np.uint8)
    c1, p1 = d.read(z[:-1], out, 0)
    assert (c1, p1) == (len(z) - 1, 12) and not d.is_done()
    c2, p2 = d.read(z[-1:], out[:12], 12)
    assert (c2, p2) == (1, 0) and d.is_done() and out[:12].tobytes() == b"Hello world!"
    assert d.read(b"xx", out, 12) == (0, 0)  # :185-187 after Done

    # decompress.rs:1309-1325 zero_length: spliced empty stored blocks, zero-size output
    e0 = bytearray(zlib.compress(b"", 1))
    for _ in range(10):
        e0[2:2] = bytes([0, 0, 0, 0xFF, 0xFF])
    d = F.Decompressor(ctx)
    assert d.read(bytes(e0), np.zeros(0, np.uint8), 0) == (len(e0), 0) and d.is_done()

    # chunking invariance of the streaming facade (fuzz targets inflate_bytewise*, inflate_split)
    data = cases.sparse_bytes(rng, 3000)
    z = zlib.compress(data, 6)
    d = F.Decompressor(ctx)
    out = np.zeros(len(data) + 10, np.uint8)
    pos = 0
    for i in range(0, len(z), 37):
        c, p = d.read(z[i:i + 37], out, pos)
        pos += p
    assert d.is_done() and out[:pos].tobytes() == data
    # small rolling output window: "output full" post-condition
    d = F.Decompressor(ctx)
    got = bytearray()
    win = np.zeros(100, np.uint8)
    c, p = d.read(z, win, 0)
    got += win[:p].tobytes()
    while not d.is_done():
        c, p = d.read(b"", win, 0)
        got += win[:p].tobytes()
        assert p > 0 or d.is_done()
    assert bytes(got) == data

    # bounded API: decompress.rs:1111-1144
    assert F.decompress_to_vec_bounded(z, len(data), ctx) == data
    with pytest.raises(F.BoundedDecompressionError) as e:
        F.decompress_to_vec_bounded(z, len(data) - 1, ctx)
    assert e.value.kind == "OutputTooLarge" and e.value.partial_output == data[:-1]
    with pytest.raises(F.BoundedDecompressionError) as e:
        F.decompress_to_vec_bounded(z[:-5], 1 << 20, ctx)
    assert e.value.inner.kind == "InsufficientInput"
    for name, g in cases.golden_streams()[-2:]:
        with pytest.raises(F.DecompressionError) as e:
            F.decompress_to_vec(g, ctx)
        assert e.value.kind == "BadLiteralLengthHuffmanTree"

    # UltraFastCompressor: new / write_data / finish (ultrafast.rs:70-181), single and multiple calls
    w = F.UltraFastCompressor(io.BytesIO(), ctx)
    w.write_data(data)
    assert w.finish().getvalue() == oracle.compress_ultra_fast(data)
    w = F.UltraFastCompressor(io.BytesIO(), ctx)
    parts = [data[:1001], data[1001:1001], data[1001:2500], data[2500:]]
    for p in parts:
        w.write_data(p)
    multi = w.finish().getvalue()
    assert zlib.decompress(multi) == data and F.decompress_to_vec(multi, ctx) == data
    assert multi == oracle.compress_ultra_fast_calls(parts)  # the reference's bytes for this call pattern
    assert F.UltraFastCompressor(io.BytesIO(), ctx).finish().getvalue() == oracle.compress_ultra_fast(b"")

    # Compressor::new(w, 0, zlib): stored; other levels are out of scope
    c = F.Compressor(io.BytesIO(), 0, True, ctx)
    c.write_data(data[:700])
    c.write_data(data[700:])
    assert c.finish().getvalue() == oracle.compress_stored(data)
    raw = F.Compressor(io.BytesIO(), 0, False, ctx)
    raw.write_data(data)
    assert zlib.decompress(raw.finish().getvalue(), -15) == data
    with pytest.raises(NotImplementedError):
        F.Compressor(io.BytesIO(), 1, True, ctx)


def test_multi_call_ultrafast_matches_reference_semantics(oracle):
    """the reference restarts its 8-byte chunking and run state at every write_data call.
    The splice in api.py must therefore equal the oracle run per call and concatenated bit-wise."""
    from fdeflate_b200.api import _splice_ultrafast

    rng = random.Random(2)
    calls = [cases.sparse_bytes(rng, n) for n in (13, 0, 800, 5, 4096)]
    streams = [oracle.compress_ultra_fast(c) for c in calls]
    spliced = _splice_ultrafast(streams, calls)
    assert zlib.decompress(spliced) == b"".join(calls)
    assert _splice_ultrafast(streams[:1], calls[:1]) == streams[0]
    # byte for byte against the oracle's call-by-call compressor (new / write_data / finish, ultrafast.rs:70-181)
    assert spliced == oracle.compress_ultra_fast_calls(calls)
    for _ in range(300):
        d = cases.sparse_bytes(rng, rng.randrange(1, 3000))
        cuts = sorted(rng.randrange(len(d) + 1) for _ in range(rng.randrange(1, 7)))
        calls = [d[a:b] for a, b in zip([0] + cuts, cuts + [len(d)])]
        streams = [oracle.compress_ultra_fast(c) for c in calls]
        assert _splice_ultrafast(streams, calls) == oracle.compress_ultra_fast_calls(calls)


@pytest.mark.emul
def test_host_api_on_emulator(emul_ctx, oracle):
    _suite(emul_ctx, oracle)


@pytest.mark.gpu
def test_host_api_on_gpu(gpu_ctx, oracle):
    _suite(gpu_ctx, oracle)


@pytest.mark.gpu
def test_ultrafast_deflate_writes_pinned_host_slots_itself(monkeypatch, oracle):
    """fdb_deflate_ultrafast_batch with a PINNED output buffer and FDB_DIRECT_OUT=1: the kernel stores into the caller's
    slots directly (mapped host memory), so every stream is byte-identical to the oracle and nothing outside
    [off, off + out_len) is touched -- not even the rest of the slot, which the copy path would overwrite up to the
    widest row's width."""
    import torch

    monkeypatch.setenv("FDB_DIRECT_OUT", "1")  # read when the context is created
    gpu_ctx = F.Context(0)

    rng = random.Random(7)
    sizes = [0, 1, 7, 8, 9, 63, 64, 65, 1000, 2047, 2048, 2049, 70000, 262400, 5, 300000]
    raw = [cases.sparse_bytes(rng, n) for n in sizes] + [bytes(40000), bytes([3]) * 5000]
    n = len(raw)
    in_base, in_off, in_len = F.Context._pack(raw, 16)
    stride = (max(gpu_ctx.ultrafast_bound(len(r)) for r in raw) + 64 + 15) // 16 * 16
    # evenly spaced slots (the layout the 2-D payload copy handles), every other one starting at an odd address
    out_off = np.array([64 + i * stride + (3 if i % 2 else 0) for i in range(n)], dtype=np.uint64)
    out_cap = np.array([gpu_ctx.ultrafast_bound(len(r)) for r in raw], dtype=np.uint64)
    h_out = torch.full((64 + n * stride + 64,), 0xAB, dtype=torch.uint8, pin_memory=True)
    out_len, st = gpu_ctx.deflate_ultrafast_packed(in_base, in_off, in_len, h_out.numpy(), out_off, out_cap)
    assert (st == 0).all()
    got = h_out.numpy()
    mask = np.ones(got.shape[0], dtype=bool)
    for i in range(n):
        a, l = int(out_off[i]), int(out_len[i])
        assert got[a:a + l].tobytes() == oracle.compress_ultra_fast(raw[i]), "stream %d" % i
        mask[a:a + l] = False
    assert (got[mask] == 0xAB).all(), "bytes outside the streams were written"
    # a slot that is too small reports it and leaves the other streams alone
    h_out.fill_(0xAB)
    small = out_cap.copy()
    small[13] = 100
    out_len2, st2 = gpu_ctx.deflate_ultrafast_packed(in_base, in_off, in_len, h_out.numpy(), out_off, small)
    assert st2[13] == 18 and (np.delete(st2, 13) == 0).all() and (np.delete(out_len2, 13) == np.delete(out_len, 13)).all()
    got = h_out.numpy()
    a = int(out_off[13])
    assert (got[a + 100:a + stride - 3] == 0xAB).all()
    for i in (0, 9, 12, 14, 17):
        a, l = int(out_off[i]), int(out_len[i])
        assert got[a:a + l].tobytes() == oracle.compress_ultra_fast(raw[i])
