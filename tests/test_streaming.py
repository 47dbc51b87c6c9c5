"""The streaming decoder: Decompressor::read with its state kept on the device.  The reference tests it
with decompress_by_chunks (src/decompress/tests/test_utils.rs:47-87) and the fuzz targets inflate_bytewise{,2,3} /
inflate_split: the same input fed whole, byte by byte and in arbitrary chunks must give the same output or an error in
every case.  Here the chunked runs go through fdb_stream_* (emulator on CPU, CUDA with -m gpu) and are compared with the
oracle's own streaming decoder; many decoders advance in one call; byte-wise feeding is linear in the stream length."""
import random
import time
import zlib

import numpy as np
import pytest

import cases
import fdeflate_b200 as F


def by_chunks(ctx, data: bytes, chunks, room: int = 1_000_000, ignore_adler: bool = True):
    """test_utils.rs:47-87 with our Decompressor: -> (status, output); status = 0 or the DecompressionError number,
    -2 when the input ran out before the stream was complete, -5 for too many iterations"""
    d = F.Decompressor(ctx)
    if ignore_adler:
        d.ignore_adler32()
    out = np.zeros(room, dtype=np.uint8)
    in_pos = out_pos = 0
    chunks = iter(chunks)
    it = 0
    try:
        while not d.is_done():
            it += 1
            if it > 20000:
                return -5, out[:out_pos].tobytes()
            c = next(chunks, 0)
            if c == 0 and in_pos >= len(data):
                return -2, out[:out_pos].tobytes()
            try:
                consumed, written = d.read(data[in_pos:in_pos + c], out, out_pos)
            except F.DecompressionError as e:
                return e.code, out[:out_pos].tobytes()
            in_pos += consumed
            out_pos += written
        return 0, out[:out_pos].tobytes()
    finally:
        d.close()


def _oracle_by_chunks(oracle, data, chunks):
    """the oracle's streaming decoder over the same chunks -> (status, output): 0, an error, or -2 (stream incomplete)"""
    d = oracle.Decompressor()
    d.ignore_adler32()
    out = np.zeros(1_000_000, dtype=np.uint8)
    in_pos = out_pos = 0
    it = 0
    chunks = iter(chunks)
    while not d.is_done():
        it += 1
        c = next(chunks, 0)
        end = min(in_pos + c, len(data))
        st, consumed, written = d.read(data[in_pos:end], out, out_pos)
        if st != 0:
            return st, out[:out_pos].tobytes()
        in_pos += consumed
        out_pos += written
        if consumed == 0 and written == 0 and c == 0:
            return -2, out[:out_pos].tobytes()
        if it > 200000:
            return -5, b""
    return 0, out[:out_pos].tobytes()


def _streams(oracle):
    rng = random.Random(6)
    out = [d for _, d in cases.golden_streams()]
    for lvl, strat in ((0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (9, zlib.Z_FILTERED)):
        data = cases.payload(rng, rng.randrange(6), 6000)
        co = zlib.compressobj(lvl, zlib.DEFLATED, 15, 8, strat)
        out.append(co.compress(data[:3000]) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(data[3000:]) + co.flush())
    out.append(oracle.compress_ultra_fast(cases.sparse_bytes(rng, 5000)))
    out.append(oracle.compress_stored(cases.sparse_bytes(rng, 70000)))
    skew = bytes(min(255, int(rng.expovariate(0.08))) for _ in range(20000))  # 13-15 bit codes
    out.append(zlib.compress(skew, 9))
    return out


def _suite(ctx, oracle, bytewise_limit):
    rng = random.Random(9)
    for z in _streams(oracle):
        whole = by_chunks(ctx, z, [len(z)])
        want = _oracle_by_chunks(oracle, z, [len(z)])
        assert whole[0] == want[0] and (want[0] != 0 or whole[1] == want[1])
        patterns = [[1] * len(z)] if len(z) <= bytewise_limit else []
        patterns.append([rng.choice([1, 2, 3, 7, 64, 500]) for _ in range(len(z) + 1)])
        patterns.append([len(z) // 2, len(z)])                                    # inflate_split
        patterns.append([0, 1, 0, 0, 5] + [rng.randrange(0, 40) for _ in range(len(z))])
        for p in patterns:
            got = by_chunks(ctx, z, p)
            # inflate_bytewise3: same output, or an error both ways (the error may surface as "incomplete" when the
            # stream is cut short, which a chunked run cannot tell from "more input to come")
            assert (got[0] == 0) == (whole[0] == 0), (got[0], whole[0], len(z), p[:8])
            if whole[0] == 0:
                assert got[1] == whole[1]
    # small output rooms: the caller's buffer is full again and again (a match cut by the end of the room is resumed)
    z = zlib.compress(bytes(5000) + cases.sparse_bytes(rng, 3000) + b"ab" * 2000, 6)
    want = zlib.decompress(z)
    for room in (1, 2, 3, 17, 258, 1000):
        d = F.Decompressor(ctx)
        out = bytearray()
        buf = np.zeros(room, dtype=np.uint8)
        fed = False
        for _ in range(100000):
            if d.is_done():
                break
            _, w = d.read(b"" if fed else z, buf, 0)
            fed = True
            out += buf[:w].tobytes()
        assert d.is_done() and bytes(out) == want, room
        d.close()
    # checksum: verified at the end unless ignored (decompress.rs:1261-1307)
    bad = bytearray(zlib.compress(b"Hello world!", 1))
    bad[-1] ^= 1
    assert by_chunks(ctx, bytes(bad), [5] * 10, ignore_adler=False)[0] == 15
    assert by_chunks(ctx, bytes(bad), [5] * 10, ignore_adler=True) == (0, b"Hello world!")
    # many decoders in one call, each at its own pace
    zs = _streams(oracle)[-7:]
    expect = [_oracle_by_chunks(oracle, z, [len(z)]) for z in zs]
    ids = ctx.stream_open(len(zs))
    pos = [0] * len(zs)
    outs = [bytearray() for _ in zs]
    final = [None] * len(zs)
    for step in range(400):
        datas = []
        for k, z in enumerate(zs):
            c = rng.randrange(0, 900)
            datas.append(z[pos[k]:pos[k] + c])
            pos[k] += c
        st, chunks = ctx.stream_read(ids, datas, [rng.randrange(0, 3000) for _ in zs], F.FLAG_IGNORE_ADLER32)
        for k in range(len(zs)):
            outs[k] += chunks[k]
            if st[k] >= 0 and final[k] is None:
                final[k] = int(st[k])
        if all(f is not None for f in final):
            break
    ctx.stream_close(ids)
    for k in range(len(zs)):
        assert final[k] == expect[k][0] and (expect[k][0] != 0 or bytes(outs[k]) == expect[k][1]), k


@pytest.mark.emul
def test_streaming_on_emulator(emul_ctx, oracle):
    _suite(emul_ctx, oracle, bytewise_limit=700)


@pytest.mark.gpu
def test_streaming_on_gpu(gpu_ctx, oracle):
    _suite(gpu_ctx, oracle, bytewise_limit=4000)


@pytest.mark.gpu
def test_streaming_bytewise_is_linear(gpu_ctx, oracle):
    """byte-wise feeding of a long single-block stream must cost O(n), not O(n^2): twice the bytes, about
    twice the time (the old facade re-inflated the whole prefix on every call: 4x)."""
    rng = random.Random(1)

    def run(n):
        data = cases.sparse_bytes(rng, n)
        z = oracle.compress_ultra_fast(data)  # one block for the whole stream
        t0 = time.perf_counter()
        st, out = by_chunks(gpu_ctx, z, [1] * len(z), room=n + 16)
        dt = time.perf_counter() - t0
        assert st == 0 and out == data
        return dt / len(z)

    run(2000)
    per_byte_small = run(20000)
    per_byte_big = run(80000)
    assert per_byte_big < 1.6 * per_byte_small, (per_byte_small, per_byte_big)
    # ... and a 1 MiB stream in 4 KiB reads
    data = cases.sparse_bytes(rng, 1 << 20)
    z = zlib.compress(data, 6)
    st, out = by_chunks(gpu_ctx, z, [4096] * (len(z) // 4096 + 2), room=(1 << 20) + 16)
    assert st == 0 and out == data
