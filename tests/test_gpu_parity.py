"""-m gpu: the product library on a real B200, through the C ABI, against the oracle."""
import random
import zlib

import numpy as np
import pytest

import cases
import parity
from fdeflate_b200 import FLAG_GENERAL_ONLY, FLAG_IGNORE_ADLER32

pytestmark = pytest.mark.gpu


def test_library_is_the_cuda_build(gpu_ctx):
    assert "CUDA" in gpu_ctx.lib.version()


def test_deflate_ultrafast_byte_identical(gpu_ctx, oracle):
    inputs = cases.compress_inputs(5, 60, [10, 100, 1000, 5000, 20000, 70000, 300000])
    parity.check_deflate_ultrafast(gpu_ctx, inputs, align=16)
    parity.check_deflate_ultrafast(gpu_ctx, inputs, align=1)


def test_deflate_stored_byte_identical(gpu_ctx, oracle):
    rng = random.Random(2)
    inputs = [b"", b"a", bytes(65534), bytes(65535), bytes(65536), cases.sparse_bytes(rng, 131070),
              cases.sparse_bytes(rng, 131071), cases.sparse_bytes(rng, 700000)]
    parity.check_deflate_stored(gpu_ctx, inputs, align=16)
    parity.check_deflate_stored(gpu_ctx, inputs, align=1)


def test_inflate_golden_vectors(gpu_ctx, oracle):
    g = [(d, 4096) for _, d in cases.golden_streams()]
    parity.check_inflate(gpu_ctx, g, FLAG_GENERAL_ONLY)
    parity.check_inflate(gpu_ctx, g, FLAG_GENERAL_ONLY | FLAG_IGNORE_ADLER32)
    parity.check_inflate(gpu_ctx, g, 0)
    st = parity.check_inflate(gpu_ctx, g[-3:], FLAG_IGNORE_ADLER32)
    assert list(st) == [0, 9, 9]


def test_inflate_general_mixed_streams(gpu_ctx, oracle):
    # BASELINE config 4: stored / fixed / dynamic, flushes, small windows, long overlapping matches,
    # 13-15 bit codes, truncations, bit flips, capacity limits
    c = cases.mixed_zlib_cases(21, 120, [0, 1, 5, 100, 1000, 5000, 20000, 200000])
    parity.check_inflate(gpu_ctx, c, FLAG_GENERAL_ONLY)
    parity.check_inflate(gpu_ctx, c, FLAG_GENERAL_ONLY | FLAG_IGNORE_ADLER32, align=1)
    parity.check_inflate(gpu_ctx, c, 0)


def test_inflate_long_constant_runs(gpu_ctx, oracle):
    # chains of length-258 distance-1 matches (>= 1 MiB) and a 32 KiB-period pattern
    big = bytes(3 << 20)
    per = (bytes(range(256)) * 128 * 40)[: 1 << 20]
    c = [(zlib.compress(big, 6), len(big)), (zlib.compress(per, 9), len(per)),
         (oracle.compress_ultra_fast(big), len(big))]
    parity.check_inflate(gpu_ctx, c, 0)
    parity.check_inflate(gpu_ctx, c, FLAG_GENERAL_ONLY)


def test_inflate_fast_path_ultrafast_streams(gpu_ctx, oracle):
    from fdeflate_b200 import synth_tiles_host

    rng = random.Random(4)
    datas = [b"", b"a", bytes(1), bytes(100000)] + [cases.sparse_bytes(rng, n) for n in (100, 1000, 3000, 50000, 120000, 900000)]
    datas += [t.tobytes() for t in synth_tiles_host(0, 24, 256, 256, 99)]
    streams = [oracle.compress_ultra_fast(d) for d in datas]
    exact = [(s, len(d)) for s, d in zip(streams, datas)]
    parity.check_inflate(gpu_ctx, exact, 0, expect_general=0)
    parity.check_inflate(gpu_ctx, exact, 0, align=1, expect_general=0)
    parity.check_inflate(gpu_ctx, exact, FLAG_GENERAL_ONLY)
    parity.check_inflate(gpu_ctx, [(s, c + 7) for s, c in exact], 0, expect_general=0)
    parity.check_inflate(gpu_ctx, [(s, max(0, c - 1)) for s, c in exact], 0)
    dmg = []
    for s, c in exact:
        dmg += cases.damaged(rng, s, c)
    parity.check_inflate(gpu_ctx, dmg, 0)
    parity.check_inflate(gpu_ctx, dmg, FLAG_IGNORE_ADLER32)


def test_inflate_fast_path_foreign_token_sequences(gpu_ctx, oracle):
    crafted = cases.crafted_uf_cases(3, sizes=(0, 1, 2, 10, 100, 1000, 5000, 60000))
    c = [(s, len(e) if e is not None else 100000) for s, e in crafted]
    parity.check_inflate(gpu_ctx, c, 0)
    parity.check_inflate(gpu_ctx, c, FLAG_GENERAL_ONLY)


def test_batch_composition_invariance(gpu_ctx, oracle):
    rng = random.Random(8)
    c = cases.mixed_zlib_cases(5, 20, [100, 3000, 50000])
    c += [(oracle.compress_ultra_fast(cases.sparse_bytes(rng, 40000)), 40000) for _ in range(20)]
    base = gpu_ctx.inflate_batch([x[0] for x in c], [x[1] for x in c])
    perm = list(range(len(c)))
    rng.shuffle(perm)
    sh = gpu_ctx.inflate_batch([c[i][0] for i in perm], [c[i][1] for i in perm])
    for k, i in enumerate(perm):
        assert sh[0][k] == base[0][i] and sh[1][k] == base[1][i]


def test_config1_single_image_roundtrip(gpu_ctx, oracle):
    # BASELINE config 1: one 1024x1024 RGBA filtered image, compress then decompress
    from fdeflate_b200 import compress_to_vec_ultra_fast, decompress_to_vec, synth_tiles_host

    img = synth_tiles_host(0, 1, 1024, 1024, 7)[0].tobytes()
    assert len(img) == 4195328
    z = compress_to_vec_ultra_fast(img, gpu_ctx)
    assert z == oracle.compress_ultra_fast(img)
    assert decompress_to_vec(z, gpu_ctx) == img == zlib.decompress(z)


def _device_batch(gpu_ctx, n_tiles, seed):
    """config 2/3 at full size, device resident: synth tiles -> compress -> inflate; returns torch tensors"""
    import torch

    dev = torch.device("cuda:0")
    tb = 262400
    tiles = torch.empty(n_tiles * tb, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    gpu_ctx.synth_tiles_device(tiles.data_ptr(), 0, n_tiles, 256, 256, seed, s)
    bound = gpu_ctx.ultrafast_bound(tb)
    u64 = torch.uint64 if hasattr(torch, "uint64") else torch.int64
    in_off = (torch.arange(n_tiles, dtype=torch.int64, device=dev) * tb)
    in_len = torch.full((n_tiles,), tb, dtype=torch.int64, device=dev)
    c_off = (torch.arange(n_tiles, dtype=torch.int64, device=dev) * bound)
    c_cap = torch.full((n_tiles,), bound, dtype=torch.int64, device=dev)
    comp = torch.empty(n_tiles * bound, dtype=torch.uint8, device=dev)
    c_len = torch.zeros(n_tiles, dtype=torch.int64, device=dev)
    c_st = torch.full((n_tiles,), -7, dtype=torch.int32, device=dev)
    gpu_ctx.deflate_ultrafast_device(tiles.data_ptr(), in_off.data_ptr(), in_len.data_ptr(), comp.data_ptr(),
                                     c_off.data_ptr(), c_cap.data_ptr(), c_len.data_ptr(), c_st.data_ptr(), n_tiles, s)
    out = torch.empty(n_tiles * tb, dtype=torch.uint8, device=dev)
    o_len = torch.zeros(n_tiles, dtype=torch.int64, device=dev)
    o_st = torch.full((n_tiles,), -7, dtype=torch.int32, device=dev)
    gpu_ctx.inflate_device(comp.data_ptr(), c_off.data_ptr(), c_len.data_ptr(), out.data_ptr(), in_off.data_ptr(),
                           in_len.data_ptr(), o_len.data_ptr(), 0, o_st.data_ptr(), n_tiles, 0, s)
    torch.cuda.synchronize()
    return tiles, comp, c_off, c_len, c_st, out, o_len, o_st


def test_config2_config3_full_size_properties(gpu_ctx, oracle):
    """BASELINE configs 2 and 3 at full size (4096 tiles of 256x256 RGBA, 1.07 GB): the oracle is too
    slow for all of it, so check size-independent properties on everything (round trip equality,
    all statuses Ok, fast path taken) and byte parity against the oracle on a sample of tiles."""
    import torch

    n = 4096
    tiles, comp, c_off, c_len, c_st, out, o_len, o_st = _device_batch(gpu_ctx, n, 2024)
    assert int((c_st != 0).sum()) == 0 and int((o_st != 0).sum()) == 0
    assert gpu_ctx.last_general_count(torch.cuda.current_stream().cuda_stream) == 0
    assert int((o_len != 262400).sum()) == 0
    assert torch.equal(out, tiles)                      # encode -> decode round trip over 1.07 GB
    ratio = float(c_len.sum()) / (n * 262400)
    assert 0.2 < ratio < 0.6
    host_tiles = tiles.view(n, 262400)
    lens = c_len.cpu().numpy()
    offs = c_off.cpu().numpy()
    sample = [0, 4095] + [int(i) for i in np.random.default_rng().choice(n, size=6, replace=False)]  # a fresh sample every run
    for i in sample:
        t = host_tiles[i].cpu().numpy().tobytes()
        ref = oracle.compress_ultra_fast(t)
        got = comp[int(offs[i]): int(offs[i]) + int(lens[i])].cpu().numpy().tobytes()
        assert got == ref, f"tile {i}: compressed bytes differ from the oracle"
        assert oracle.decompress_to_vec(got) == (0, t)


def test_synth_host_and_device_agree(gpu_ctx):
    import torch

    from fdeflate_b200 import synth_tiles_host

    h = synth_tiles_host(5, 3, 256, 256, 77)
    d = torch.empty(3 * 262400, dtype=torch.uint8, device="cuda:0")
    gpu_ctx.synth_tiles_device(d.data_ptr(), 5, 3, 256, 256, 77, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().reshape(3, 262400), h)


def test_host_pipeline_chunking_does_not_change_results(gpu_ctx, oracle):
    """Host-buffer batches overlap copies and kernels chunk by chunk over several CUDA streams; results
    must not depend on the chunk size (tiny chunks here: every batch becomes up to 64 chunks)."""
    rng = random.Random(12)
    inputs = cases.compress_inputs(9, 40, [100, 3000, 20000, 70000])
    mixed = cases.mixed_zlib_cases(31, 30, [0, 100, 5000, 50000])
    uf = [(oracle.compress_ultra_fast(d), len(d)) for d in inputs]
    try:
        gpu_ctx.set_pipeline_chunk(16384)
        parity.check_deflate_ultrafast(gpu_ctx, inputs, align=16)
        parity.check_deflate_stored(gpu_ctx, inputs[:30], align=16)
        parity.check_inflate(gpu_ctx, uf, 0, expect_general=0)
        parity.check_inflate(gpu_ctx, mixed + uf, 0)
        parity.check_inflate(gpu_ctx, mixed, FLAG_GENERAL_ONLY)
        dmg = []
        for s, c in uf[:20]:
            dmg += cases.damaged(rng, s, c)
        parity.check_inflate(gpu_ctx, dmg, 0)
    finally:
        gpu_ctx.set_pipeline_chunk(0)


def test_inflate_parallel_block_decode(gpu_ctx, oracle):
    """Streams long enough for the general kernel's sub-sequence-parallel block decode (dynamic and
    fixed blocks, flush points, literal-only and match-heavy data, codes longer than the tables, runs
    of distance-1 matches), with exact-fit / one-short slots and truncated inputs: status and bytes
    must equal the oracle's (the parallel path only commits regular segments and hands everything
    else to the sequential reader at a token boundary)."""
    from fdeflate_b200 import synth_tiles_host

    rng = random.Random(77)
    tile = synth_tiles_host(3, 1, 256, 256, 5, gpu_ctx.lib)[0].tobytes()
    text = bytes(rng.choice(b"abcdefghij klmnop\n") for _ in range(240000))
    skew = bytes(min(255, int(rng.expovariate(0.08))) for _ in range(200000))  # 13-15 bit codes
    rnd = bytes(rng.getrandbits(8) for _ in range(120000))
    mix = tile[:160000] + bytes(80000) + text[:120000] + bytes([7]) * 36000
    c = []
    for data in (tile[:280000], text, skew, rnd, mix):
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
    st = parity.check_inflate(gpu_ctx, exact, FLAG_GENERAL_ONLY)
    assert (st == 0).all()
    parity.check_inflate(gpu_ctx, exact, 0, align=1)
    parity.check_inflate(gpu_ctx, [(s, n - 1) for s, n in exact], FLAG_GENERAL_ONLY)       # OutputTooLarge
    parity.check_inflate(gpu_ctx, [(s, n // 2 + 5) for s, n in exact], FLAG_GENERAL_ONLY)  # ... mid-stream
    cut = []
    for s, n in exact:
        cut += [(s[: len(s) - 5], n), (s[: len(s) // 2], n), (s[: max(0, len(s) - 1500)], n)]
        b = bytearray(s)
        b[len(b) // 2] ^= 0x10  # a flipped bit mid-stream: whatever the oracle says
        cut.append((bytes(b), n))
    parity.check_inflate(gpu_ctx, cut, FLAG_GENERAL_ONLY)
    parity.check_inflate(gpu_ctx, cut, FLAG_GENERAL_ONLY | FLAG_IGNORE_ADLER32)


def test_inflate_long_streams_span_by_span(gpu_ctx, oracle):
    """GPU twin of the emulator test: long ultra-fast-format streams decoded span by span by many warps
    (count pass from a guessed bit, chain check, write pass) must equal the oracle bit for bit; damaged,
    truncated and short-slot cases must come out exactly as the oracle says."""
    from test_emul_kernels import _long_uf_cases

    c = _long_uf_cases(oracle, gpu_ctx.lib, 41) + _long_uf_cases(oracle, gpu_ctx.lib, 42)
    exact = [(s, len(d)) for s, d in c]
    for align in (16, 1):
        parity.check_inflate(gpu_ctx, exact, 0, align=align, expect_general=0)
        assert gpu_ctx.last_split_spans() >= 4 * len(c) - 8
    small = [(oracle.compress_ultra_fast(d), len(d)) for d in cases.compress_inputs(2, 20, [100, 5000, 70000])]
    mixed = cases.mixed_zlib_cases(4, 10, [100, 3000, 50000])
    parity.check_inflate(gpu_ctx, small[:30] + exact[:3] + mixed[:40] + exact[3:] + small[30:60], 0)
    parity.check_inflate(gpu_ctx, [(s, n + 100) for s, n in exact], 0, expect_general=0)
    parity.check_inflate(gpu_ctx, [(s, n - 1) for s, n in exact], 0)
    parity.check_inflate(gpu_ctx, [(s, n // 2) for s, n in exact], 0)
    rng = random.Random(5)
    dmg = []
    for s, n in exact:
        dmg += [(s[: len(s) - 3], n), (s[: len(s) // 2], n), (s + b"xyz", n)]
        for pos in (60, len(s) // 3, len(s) - 10):
            b = bytearray(s)
            b[pos] ^= 1 << rng.randrange(8)
            dmg.append((bytes(b), n))
        b = bytearray(s)
        b[-1] ^= 0xff
        dmg.append((bytes(b), n))
    parity.check_inflate(gpu_ctx, dmg, 0)
    parity.check_inflate(gpu_ctx, dmg, FLAG_IGNORE_ADLER32)


def test_config5_ragged_large_streams_roundtrip(gpu_ctx, oracle):
    """BASELINE configs[4] at a reduced count: streams of 64 KiB .. 16 MiB (log-uniform) through the
    host-buffer calls: deflate -> inflate must give the input back, the biggest and a few others are
    compared with the oracle byte for byte, and the long ones must have taken the span path."""
    from fdeflate_b200 import synth_tiles_host

    rng = np.random.default_rng(5)
    sizes = np.exp(rng.uniform(np.log(64 << 10), np.log(16 << 20), 24))
    row = 1 + 4 * 1024
    datas = [synth_tiles_host(100 + i, 1, 1024, max(1, int(sz) // row), 5, gpu_ctx.lib)[0].tobytes() for i, sz in enumerate(sizes)]
    datas.append(bytes(3 << 20))                                     # one run of 3 MiB
    datas.append(np.random.default_rng(6).integers(0, 256, 2 << 20, dtype=np.uint8).tobytes())  # incompressible
    comp = gpu_ctx.deflate_ultrafast_batch(datas)
    order = np.argsort([len(d) for d in datas])
    for i in list(order[-2:]) + list(order[:3]):
        assert comp[i] == oracle.compress_ultra_fast(datas[i])
    st, outs, cons = gpu_ctx.inflate_batch(comp, [len(d) for d in datas])
    assert (st == 0).all() and gpu_ctx.last_general_count() == 0
    assert gpu_ctx.last_split_spans() > 0
    for d, o, s, k in zip(datas, outs, comp, cons):
        assert o == d and k == len(s)
    for d, s in zip(datas, comp):
        assert zlib.adler32(d) == int.from_bytes(s[-4:], "big")


def test_config5_sweep_shape_device_resident(gpu_ctx, oracle):
    """BASELINE configs[4] at a quarter of one GPU's share: 2048 streams of 64 KiB .. 16 MiB (log-uniform, ~6 GB),
    device-resident, ultra-fast deflate (segments) then inflate (spans) the way bench.py's sweep runs them.  Size-
    independent properties on everything (statuses, lengths, round trip, every stream on the fast path) and byte parity
    with the oracle in both directions on a random sample."""
    import torch

    import fdeflate_b200 as F

    n, W = 2048, 1024
    row = 1 + 4 * W
    rng = np.random.default_rng(5)
    sizes = np.exp(rng.uniform(np.log(64 << 10), np.log(16 << 20), n))
    heights = np.maximum(1, (sizes / row).astype(np.int64))
    lens = heights * row
    offs = np.zeros(n, dtype=np.int64)
    offs[1:] = np.cumsum((lens[:-1] + 15) & ~15)
    total = int(offs[-1] + lens[-1])
    dev = torch.device("cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    raw = torch.empty(total + 16, dtype=torch.uint8, device=dev)
    for i in range(n):
        gpu_ctx.synth_tiles_device(raw.data_ptr() + int(offs[i]), 1000 + i, 1, W, int(heights[i]), 5, s)
    bounds = np.array([gpu_ctx.ultrafast_bound(int(l)) for l in lens], dtype=np.int64)
    coffs = np.zeros(n, dtype=np.int64)
    coffs[1:] = np.cumsum(bounds[:-1])
    comp = torch.empty(int(coffs[-1] + bounds[-1]), dtype=torch.uint8, device=dev)
    out = torch.zeros(total + 16, dtype=torch.uint8, device=dev)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_off, d_len, d_coff, d_ccap = T(offs), T(lens), T(coffs), T(bounds)
    c_len = torch.zeros(n, dtype=torch.int64, device=dev)
    c_st = torch.zeros(n, dtype=torch.int32, device=dev)
    o_len = torch.zeros(n, dtype=torch.int64, device=dev)
    o_st = torch.zeros(n, dtype=torch.int32, device=dev)
    try:
        gpu_ctx.set_split_large(True)
        gpu_ctx.deflate_ultrafast_device(raw.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), comp.data_ptr(), d_coff.data_ptr(),
                                         d_ccap.data_ptr(), c_len.data_ptr(), c_st.data_ptr(), n, s)
        torch.cuda.synchronize()
        assert int(c_st.abs().sum()) == 0
        # the write pass of a span either reads the lane records its count pass left or counts again: a pool that
        # holds every span (1 GiB), the default one (256 MiB: about half of these 3.9 GB of streams) and none
        for pool in (1 << 30, 256 << 20, 0):
            gpu_ctx.set_split_scratch(pool)
            out.zero_()
            o_len.zero_()
            o_st.fill_(-1)
            gpu_ctx.inflate_device(comp.data_ptr(), d_coff.data_ptr(), c_len.data_ptr(), out.data_ptr(), d_off.data_ptr(), d_len.data_ptr(),
                                   o_len.data_ptr(), 0, o_st.data_ptr(), n, F.FLAG_SPLIT_LARGE, s)
            torch.cuda.synchronize()
            assert int(o_st.abs().sum()) == 0, pool
            assert gpu_ctx.last_general_count(s) == 0 and gpu_ctx.last_split_spans(s) > n
            assert torch.equal(o_len, d_len)
            for i in range(n):  # encode -> decode round trip over all 6 GB (slot by slot: the padding between slots is not output)
                a = int(offs[i])
                assert torch.equal(out[a:a + int(lens[i])], raw[a:a + int(lens[i])]), (pool, i)
    finally:
        gpu_ctx.set_split_large(False)
        gpu_ctx.set_split_scratch(256 << 20)
    h_clen = c_len.cpu().numpy()
    order = np.argsort(lens)
    sample = list(order[:2]) + [int(order[-1])] + [int(i) for i in np.random.default_rng().choice(n, size=5, replace=False)]
    for i in sample:
        src = raw[int(offs[i]):int(offs[i]) + int(lens[i])].cpu().numpy().tobytes()
        z = comp[int(coffs[i]):int(coffs[i]) + int(h_clen[i])].cpu().numpy().tobytes()
        assert z == oracle.compress_ultra_fast(src), f"stream {i} ({len(src)} bytes): deflate bytes differ from the oracle"
        assert oracle.inflate_into(z, len(src))[:2] == (0, src)


def test_deflate_long_inputs_segment_by_segment(gpu_ctx, oracle):
    """GPU twin of the emulator test: inputs of >= 256 KiB encoded by many warps must be byte-identical to
    the oracle's single sequential pass (real concurrency: the words two segments share are merged with
    atomicOr)."""
    from test_emul_kernels import _check_deflate_slot_sizes, _long_deflate_inputs

    inputs = _long_deflate_inputs(gpu_ctx.lib, 11) + _long_deflate_inputs(gpu_ctx.lib, 12)
    try:
        gpu_ctx.set_split_threshold(0, 4 * 65536)  # (host-buffer calls switch at 1 MiB by default)
        for _ in range(3):  # scheduling differs from run to run
            l0 = gpu_ctx.launch_count
            parity.check_deflate_ultrafast(gpu_ctx, inputs, align=16)
            assert gpu_ctx.launch_count - l0 == 8
        parity.check_deflate_ultrafast(gpu_ctx, inputs, align=1)
        small = cases.compress_inputs(3, 30, [10, 3000, 70000])
        parity.check_deflate_ultrafast(gpu_ctx, small[:40] + inputs[2:9] + small[40:80], align=16)
        _check_deflate_slot_sizes(gpu_ctx, oracle, inputs[2])
        _check_deflate_slot_sizes(gpu_ctx, oracle, inputs[4])
    finally:
        gpu_ctx.set_split_threshold(0, 0)


def test_inflate_truncated_ends_match_the_reference_pairing(gpu_ctx, oracle):
    """GPU twin of the emulator test: streams cut inside their last tokens, slots around the exact size"""
    from test_emul_kernels import _truncation_sweep_cases

    c = _truncation_sweep_cases(oracle, gpu_ctx.lib)
    parity.check_inflate(gpu_ctx, c, FLAG_GENERAL_ONLY)
    parity.check_inflate(gpu_ctx, c, 0, align=1)
