"""CPU-side logic tests of the CUDA kernel SOURCES, run on the test-only fiber SIMT emulator
(tests/emul) and checked bit for bit against the oracle.  These cover the host logic and the warp
algorithms (scans, carries, table builds, sub-sequence synchronisation) on a machine without a GPU;
the `-m gpu` tests repeat them on the real library and hardware."""
import os
import random
import zlib

import pytest

import cases
import parity
from fdeflate_b200 import FLAG_GENERAL_ONLY, FLAG_IGNORE_ADLER32

FULL = os.environ.get("FDB_TESTS_FULL") == "1"  # every variant on every long stream (minutes on the emulator)

pytestmark = pytest.mark.emul


def test_deflate_ultrafast_byte_identical(emul_ctx, oracle):
    inputs = cases.compress_inputs(5, 25, [10, 100, 1000, 5000, 20000])
    parity.check_deflate_ultrafast(emul_ctx, inputs, align=16)
    parity.check_deflate_ultrafast(emul_ctx, inputs[:30], align=1)  # unaligned in/out slots


def test_deflate_stored_byte_identical(emul_ctx, oracle):
    rng = random.Random(2)
    inputs = [b"", b"a", bytes(65534), bytes(65535), bytes(65536), cases.sparse_bytes(rng, 131070),
              cases.sparse_bytes(rng, 131071), cases.sparse_bytes(rng, 70000)]
    parity.check_deflate_stored(emul_ctx, inputs, align=16)
    parity.check_deflate_stored(emul_ctx, inputs, align=1)


def test_inflate_golden_vectors(emul_ctx, oracle):
    g = [(d, 4096) for _, d in cases.golden_streams()]
    parity.check_inflate(emul_ctx, g, FLAG_GENERAL_ONLY)
    parity.check_inflate(emul_ctx, g, FLAG_GENERAL_ONLY | FLAG_IGNORE_ADLER32)
    parity.check_inflate(emul_ctx, g, 0)
    st = parity.check_inflate(emul_ctx, g[-3:], FLAG_IGNORE_ADLER32)
    assert list(st) == [0, 9, 9]  # example1 Ok (281 bytes), examples 2/3 BadLiteralLengthHuffmanTree


def test_inflate_general_mixed_streams(emul_ctx, oracle):
    c = cases.mixed_zlib_cases(21, 25, [0, 1, 5, 100, 1000, 5000, 20000])
    parity.check_inflate(emul_ctx, c, FLAG_GENERAL_ONLY)
    parity.check_inflate(emul_ctx, c[:120], FLAG_GENERAL_ONLY | FLAG_IGNORE_ADLER32, align=1)


def test_inflate_fast_path_ultrafast_streams(emul_ctx, emul_lib, oracle):
    from fdeflate_b200 import synth_tiles_host

    rng = random.Random(4)
    datas = [b"", b"a", bytes(1), bytes(100000)] + [cases.sparse_bytes(rng, n) for n in (100, 1000, 3000, 50000, 120000)]
    datas += [t.tobytes() for t in synth_tiles_host(0, 2, 256, 256, 99, emul_lib)]
    streams = [oracle.compress_ultra_fast(d) for d in datas]
    exact = [(s, len(d)) for s, d in zip(streams, datas)]
    parity.check_inflate(emul_ctx, exact, 0, expect_general=0)          # every stream stays on the fast path
    parity.check_inflate(emul_ctx, exact, 0, align=1, expect_general=0)
    parity.check_inflate(emul_ctx, [(s, c + 7) for s, c in exact], 0, expect_general=0)
    parity.check_inflate(emul_ctx, [(s, max(0, c - 1)) for s, c in exact], 0)  # OutputTooLarge via the general kernel
    dmg = []
    for s, c in exact:
        dmg += cases.damaged(rng, s, c)
    parity.check_inflate(emul_ctx, dmg, 0)
    parity.check_inflate(emul_ctx, dmg, FLAG_IGNORE_ADLER32)


def test_inflate_fast_path_foreign_token_sequences(emul_ctx, oracle):
    crafted = cases.crafted_uf_cases(3)
    c = [(s, len(e) if e is not None else 100000) for s, e in crafted]
    for (s, cap), (_, e) in zip(c, crafted):
        if e is not None:
            assert oracle.inflate_into(s, cap)[:2] == (0, e)
    parity.check_inflate(emul_ctx, c, 0)


def test_inflate_fast_path_rows_and_gaps(emul_ctx, oracle):
    """The single-pass decoder's lane rows: literal-only rows, zero runs inside the rows, runs that become
    gaps, a gap behind every other literal, exact-fit and short slots, unaligned slots."""
    fast, overflow = cases.uf_row_cases(11)
    for s, e in fast + overflow:
        assert e is not None and oracle.inflate_into(s, len(e))[:2] == (0, e)
    c = [(s, len(e)) for s, e in fast]
    parity.check_inflate(emul_ctx, c, 0, expect_general=0)
    parity.check_inflate(emul_ctx, c, 0, align=1, expect_general=0)
    parity.check_inflate(emul_ctx, [(s, n + 5) for s, n in c], 0, expect_general=0)
    parity.check_inflate(emul_ctx, [(s, n - 1) for s, n in c], 0)
    parity.check_inflate(emul_ctx, [(s, len(e)) for s, e in overflow], 0, expect_general=0)
    rng = random.Random(5)
    dmg = []
    for s, n in c:
        dmg += cases.damaged(rng, s, n)
    parity.check_inflate(emul_ctx, dmg, 0)


def test_batch_composition_invariance(emul_ctx, oracle):
    """a stream's result must not depend on its neighbours or its position in the batch."""
    rng = random.Random(8)
    c = cases.mixed_zlib_cases(5, 6, [100, 3000])
    c += [(oracle.compress_ultra_fast(cases.sparse_bytes(rng, 4000)), 4000) for _ in range(5)]
    base = emul_ctx.inflate_batch([x[0] for x in c], [x[1] for x in c])
    perm = list(range(len(c)))
    rng.shuffle(perm)
    sh = emul_ctx.inflate_batch([c[i][0] for i in perm], [c[i][1] for i in perm])
    for k, i in enumerate(perm):
        assert sh[0][k] == base[0][i] and sh[1][k] == base[1][i]


def test_host_pipeline_chunking_does_not_change_results(emul_ctx, oracle):
    """The host-buffer entry points cut a batch into chunks that overlap copies and kernels; results
    must not depend on the chunk size (here: tiny chunks, so every batch becomes many chunks)."""
    rng = random.Random(12)
    inputs = cases.compress_inputs(9, 12, [100, 3000, 20000])
    mixed = cases.mixed_zlib_cases(31, 8, [0, 100, 5000])
    uf = [(oracle.compress_ultra_fast(d), len(d)) for d in inputs[:40]]
    try:
        emul_ctx.set_pipeline_chunk(4096)
        parity.check_deflate_ultrafast(emul_ctx, inputs, align=16)
        parity.check_deflate_stored(emul_ctx, inputs[:20], align=16)
        parity.check_inflate(emul_ctx, uf, 0, expect_general=0)
        parity.check_inflate(emul_ctx, mixed + uf, 0)
        parity.check_inflate(emul_ctx, mixed, FLAG_GENERAL_ONLY)
        dmg = []
        for s, c in uf[:10]:
            dmg += cases.damaged(rng, s, c)
        parity.check_inflate(emul_ctx, dmg, 0)
    finally:
        emul_ctx.set_pipeline_chunk(0)


def test_inflate_parallel_block_decode(emul_ctx, oracle):
    """Streams long enough for the general kernel's sub-sequence-parallel block decode (dynamic and
    fixed blocks, flush points, literal-only and match-heavy data, codes longer than the tables, runs
    of distance-1 matches), with exact-fit / one-short slots and truncated inputs: status and bytes
    must equal the oracle's (the parallel path only commits regular segments and hands everything
    else to the sequential reader at a token boundary)."""
    from fdeflate_b200 import synth_tiles_host

    rng = random.Random(77)
    tile = synth_tiles_host(3, 1, 256, 256, 5, emul_ctx.lib)[0].tobytes()
    text = bytes(rng.choice(b"abcdefghij klmnop\n") for _ in range(60000))
    skew = bytes(min(255, int(rng.expovariate(0.08))) for _ in range(50000))  # 13-15 bit codes
    rnd = bytes(rng.getrandbits(8) for _ in range(30000))
    mix = tile[:40000] + bytes(20000) + text[:30000] + bytes([7]) * 9000
    c = []
    for data in (tile[:70000], text, skew, rnd, mix):
        for lvl in (1, 6, 9):
            c.append((zlib.compress(data, lvl), data))
        co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
        c.append((co.compress(data) + co.flush(), data))
        co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_HUFFMAN_ONLY)
        c.append((co.compress(data) + co.flush(), data))
        co, parts = zlib.compressobj(6), []
        for i in range(0, len(data), 17000):
            parts += [co.compress(data[i:i + 17000]), co.flush(zlib.Z_SYNC_FLUSH if i % 2 else zlib.Z_FULL_FLUSH)]
        c.append((b"".join(parts) + co.flush(), data))
    exact = [(s, len(d)) for s, d in c]
    st = parity.check_inflate(emul_ctx, exact, FLAG_GENERAL_ONLY)
    assert (st == 0).all()
    parity.check_inflate(emul_ctx, exact, 0, align=1)
    # the emulator is slow: the slot / damage variants take every other stream here (FDB_TESTS_FULL=1: all of them;
    # the GPU twin in test_gpu_parity.py always runs them all)
    sub = exact if FULL else exact[::2]
    parity.check_inflate(emul_ctx, [(s, n - 1) for s, n in sub], FLAG_GENERAL_ONLY)       # OutputTooLarge
    parity.check_inflate(emul_ctx, [(s, n // 2 + 5) for s, n in sub], FLAG_GENERAL_ONLY)  # ... mid-stream
    cut = []
    for s, n in sub:
        cut += [(s[: len(s) - 5], n), (s[: len(s) // 2], n), (s[: max(0, len(s) - 1500)], n)]
        b = bytearray(s)
        b[len(b) // 2] ^= 0x10  # a flipped bit mid-stream: whatever the oracle says
        cut.append((bytes(b), n))
    parity.check_inflate(emul_ctx, cut, FLAG_GENERAL_ONLY)
    parity.check_inflate(emul_ctx, cut if FULL else cut[::4], FLAG_GENERAL_ONLY | FLAG_IGNORE_ADLER32)


def _long_uf_cases(oracle, lib, seed):
    """ultra-fast-format streams long enough (>= 256 KiB compressed) for the span-by-span inflate path"""
    from fdeflate_b200 import synth_tiles_host

    rng = random.Random(seed)
    noise = bytes(rng.getrandbits(8) for _ in range(290000))                      # ~12 bits per byte
    tile = synth_tiles_host(7, 1, 1024, 160, 5, lib)[0].tobytes()                 # PNG-filtered rows
    # literals with zero runs of every length placed across span boundaries (a run token belongs to the span
    # its code starts in), and a run longer than a whole span's worth of output
    runs, lits = bytearray(), 0
    while lits < 230000:
        k = rng.choice((1, 3, 50, 4000))
        runs += bytes(rng.getrandbits(8) | 1 for _ in range(k))
        runs += bytes(rng.choice((1, 7, 8, 9, 64, 258, 259, 517, 3000, 70000)))
        lits += k
    datas = [noise, tile, bytes(runs), noise[:200000] + bytes(300000) + noise[200000:]]
    return [(oracle.compress_ultra_fast(d), d) for d in datas]


def test_inflate_long_streams_span_by_span(emul_ctx, emul_lib, oracle):
    """Long ultra-fast-format streams are cut into spans decoded by different warps (count pass from a
    guessed bit, chain check, write pass).  Bytes, lengths, consumed and status must equal the oracle's,
    whatever the slot alignment; damaged / truncated / short-slot cases must come out as the oracle says
    (the span path declines them and the general kernel reports)."""
    c = _long_uf_cases(oracle, emul_lib, 41)
    assert all(len(s) >= 262144 for s, _ in c)
    exact = [(s, len(d)) for s, d in c]
    parity.check_inflate(emul_ctx, exact, 0, expect_general=0)
    assert emul_ctx.last_split_spans() >= 4 * len(c) - 4   # the streams really took the span path
    sub = exact if FULL else exact[:2]   # (the emulator is slow; the GPU twin runs every variant on twice the streams)
    parity.check_inflate(emul_ctx, sub, 0, align=1, expect_general=0)
    assert emul_ctx.last_split_spans() > 0
    # mixed with short streams and general zlib streams in one batch
    small = [(oracle.compress_ultra_fast(d), len(d)) for d in cases.compress_inputs(2, 6, [100, 5000])]
    mixed = cases.mixed_zlib_cases(4, 6, [100, 3000])
    parity.check_inflate(emul_ctx, small[:5] + exact[:2] + mixed[:10] + exact[2:] + small[5:10], 0)
    # slots: generous, one short, half
    parity.check_inflate(emul_ctx, [(s, n + 100) for s, n in sub], 0, expect_general=0)
    parity.check_inflate(emul_ctx, [(s, n - 1) for s, n in sub], 0)
    parity.check_inflate(emul_ctx, [(s, n // 2) for s, n in exact[:2]], 0)
    # damage: truncation, a flipped bit early / late, a wrong checksum, trailing bytes
    rng = random.Random(5)
    dmg = []
    for s, n in exact[:3] if FULL else exact[1:2]:
        dmg += [(s[: len(s) - 3], n), (s[: len(s) // 2], n), (s + b"xyz", n)]
        for pos in (60, len(s) // 3, len(s) - 10):
            b = bytearray(s)
            b[pos] ^= 1 << rng.randrange(8)
            dmg.append((bytes(b), n))
        b = bytearray(s)
        b[-1] ^= 0xff
        dmg.append((bytes(b), n))
    parity.check_inflate(emul_ctx, dmg, 0)
    parity.check_inflate(emul_ctx, dmg[::2], FLAG_IGNORE_ADLER32)


def _long_deflate_inputs(lib, seed):
    """inputs of >= 256 KiB for the segment-by-segment deflate path: zero runs that start before, end after
    and span whole 64 KiB segments, run lengths around the 258-byte token limit at the boundaries, lengths
    that are and are not multiples of 8 / 512 / 64 KiB"""
    from fdeflate_b200 import synth_tiles_host

    rng = random.Random(seed)
    seg = 65536
    noise = bytes(rng.getrandbits(8) for _ in range(5 * seg + 77))
    tile = synth_tiles_host(9, 1, 1024, 100, 5, lib)[0].tobytes()
    a = bytearray(rng.getrandbits(8) | 1 for _ in range(6 * seg))
    for b0, ln in ((seg - 5, 10), (2 * seg - 300, 258 + 300), (3 * seg - 1, 1), (3 * seg + 0, 9), (4 * seg - 258, 258 * 2 + 1),
                   (5 * seg - 8, 8), (5 * seg + 512 - 3, 700)):
        a[b0:b0 + ln] = bytes(ln)
    zeros_mid = noise[:seg + 11] + bytes(3 * seg + 5) + noise[:seg]          # a run that covers whole segments
    all_zero = bytes(4 * seg + 3)
    tail_zero = noise[:2 * seg] + bytes(2 * seg + 100)                        # the run is still pending at the end
    exact = noise[:4 * seg]                                                   # no remainder segment
    return [noise, tile, bytes(a), zeros_mid, all_zero, tail_zero, exact, cases.sparse_bytes(rng, 5 * seg + 1000)]


def test_deflate_long_inputs_segment_by_segment(emul_ctx, emul_lib, oracle):
    """Inputs of >= 256 KiB are encoded by many warps (count pass, prefix sum of bit counts, write pass with
    atomicOr on the words two segments share); the bytes must equal the oracle's single sequential pass."""
    inputs = _long_deflate_inputs(emul_lib, 11)
    try:
        emul_ctx.set_split_threshold(0, 4 * 65536)  # (host-buffer calls switch at 1 MiB by default: keep the emulator run short)
        parity.check_deflate_ultrafast(emul_ctx, inputs, align=16)
        parity.check_deflate_ultrafast(emul_ctx, inputs[:4], align=1)
        small = cases.compress_inputs(3, 6, [10, 3000])
        l0 = emul_ctx.launch_count
        parity.check_deflate_ultrafast(emul_ctx, small[:5] + inputs[2:5] + small[5:10], align=16)
        assert emul_ctx.launch_count - l0 == 8  # total, plan, count, scan, write, the two work-order kernels + the one-warp-per-stream kernel
        _check_deflate_slot_sizes(emul_ctx, oracle, inputs[2])
    finally:
        emul_ctx.set_split_threshold(0, 0)


def _check_deflate_slot_sizes(ctx, oracle, data):
    """a slot of exactly the encoded size works, one byte less gives OutputBufferTooSmall and length 0"""
    import numpy as np

    ref = oracle.compress_ultra_fast(data)
    src = np.frombuffer(data, dtype=np.uint8).copy()
    for cap, want in ((len(ref), 0), (len(ref) - 1, 18), (len(ref) // 2, 18)):
        out = np.zeros(len(ref) + 64, dtype=np.uint8)
        ln, st = ctx.deflate_ultrafast_packed(src, [0], [len(data)], out, [0], [cap])
        assert st[0] == want and ln[0] == (len(ref) if want == 0 else 0)
        if want == 0:
            assert out[:len(ref)].tobytes() == ref


def _truncation_sweep_cases(oracle, lib, small=False):
    """valid streams cut at every byte of their last tokens, with slots around the exact size: the corner where the
    reference's status depends on which literals share one of ITS table entries (decompress.rs:852) -- the general
    kernel's table is smaller than the reference's and restates that pairing when a stream fails near its end"""
    from fdeflate_b200 import synth_tiles_host

    rng = random.Random(31)
    k = 5 if small else 1
    tile = synth_tiles_host(9, 1, 256, 256, 5, lib)[0].tobytes()[:30000 // k]
    text = bytes(rng.choice(b"abcdefghij klmnop\n") for _ in range(20000 // k))
    short = bytes(rng.choice(b"ab") for _ in range(6000 // k))      # 1..2-bit literal codes: long runs of paired literals
    skew = bytes(min(255, int(rng.expovariate(0.08))) for _ in range(20000 // k))
    out = []
    for data in (tile, text, short, skew):
        streams = [zlib.compress(data, 6), zlib.compress(data, 1)]
        for strat in (zlib.Z_HUFFMAN_ONLY, zlib.Z_FIXED):
            co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, strat)
            streams.append(co.compress(data) + co.flush())
        for z in streams:
            n = len(data)
            for cut in range(1, 9 if small else 14):
                for cap in ((n, n - 1) if small else (n, n - 1, n + 1, n - 2)):
                    out.append((z[:len(z) - cut], cap))
    return out


def test_inflate_truncated_ends_match_the_reference_pairing(emul_ctx, oracle):
    c = _truncation_sweep_cases(oracle, emul_ctx.lib, small=True)
    parity.check_inflate(emul_ctx, c, FLAG_GENERAL_ONLY)
