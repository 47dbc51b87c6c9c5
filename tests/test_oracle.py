"""Pins the CPU oracle against everything the reference's own tests hold for this path
(SURVEY.md 8c): golden .zz files, fuzz corpus, fixed-table known answers, huffman.rs KATs,
unit-test behaviours, plus a differential run against system zlib."""
import json
import random
import zlib
from pathlib import Path

import numpy as np
import pytest

import cases
import uf_craft

GOLD = Path(__file__).resolve().parent / "golden"
T = json.loads((GOLD / "reference_tables.json").read_text())


def test_constant_tables(oracle):
    # tables.rs:7-20, :28-55 and lib.rs:103-127 spot values
    assert list(oracle.const_table("huffman_lengths", 286)) == T["HUFFMAN_LENGTHS"]
    assert list(oracle.const_table("length_to_symbol", 256)) == T["LENGTH_TO_SYMBOL"]
    assert list(oracle.const_table("length_to_len_extra", 256)) == T["LENGTH_TO_LEN_EXTRA"]
    assert bytes(oracle.const_table("ultrafast_header", 54)) == bytes(T["ULTRAFAST_HEADER"])
    codes = oracle.const_table("huffman_codes", 286)
    assert list(codes[:8]) == [0, 2, 1, 5, 21, 29, 61, 51]
    assert (codes[255], codes[256], codes[257], codes[285]) == (6, 2303, 1279, 343)
    assert zlib.crc32(codes.astype("<u2").tobytes()) == 0x1CEB2DA0
    assert list(codes) == uf_craft.CODES


def test_tables_consistency(oracle):
    # decompress.rs:1198-1216 `tables`
    l2s = oracle.const_table("length_to_symbol", 256)
    l2e = oracle.const_table("length_to_len_extra", 256)
    for i, bits in enumerate(uf_craft.LEN_EXTRA):
        for j in range(1 << bits):
            if i == 27 and j == 31:
                continue
            assert l2e[uf_craft.LEN_BASE[i] + j - 3] == bits
            assert l2s[uf_craft.LEN_BASE[i] + j - 3] == i + 257


def test_fixed_tables_known_answer(oracle):
    # decompress.rs:1218-1233 `fixed_tables`: build_tables(288, FIXED_CODE_LENGTHS) == hard-coded tables
    lengths = [8] * 144 + [9] * 112 + [7] * 24 + [8] * 8 + [5] * 32
    st, lit, dist, _, _ = oracle.build_tables(288, lengths)
    assert st == 0
    assert list(lit[:512]) == T["FIXED_LITLEN_TABLE"]
    assert list(dist[:32]) == T["FIXED_DIST_TABLE"]
    for c in range(512, 4096, 512):
        assert (lit[c:c + 512] == lit[:512]).all()


LITERAL_ENTRY, SECONDARY = 0x8000, 0x2000


def _decode(oracle, lengths, bits: int):
    entries = oracle.const_table("litlen_table_entries", 288)
    ok, primary, secondary, _ = oracle.build_table(lengths, entries, 4096, False, True)
    assert ok
    e = int(primary[bits & 0xFFF])
    if e & LITERAL_ENTRY:
        n = (e >> 8) & 0xF
        return ("lit", [(e >> 16) & 0xFF, (e >> 24) & 0xFF][:n], e & 0xF)
    assert e & SECONDARY
    e2 = int(secondary[(e >> 16) + ((bits >> 12) & (e & 0xFF))])
    return ("sec", e2 >> 4, e2 & 0xF)


def _rev(s: str) -> int:
    return int(s.replace("_", "")[::-1], 2)


def test_huffman_rfc1951_examples(oracle):
    # huffman.rs:335-423
    assert _decode(oracle, [2, 1, 3, 3], _rev("0_0_0000000")) == ("lit", [1, 1], 2)
    assert _decode(oracle, [2, 1, 3, 3], _rev("110_110_00")) == ("lit", [2, 2], 6)
    assert _decode(oracle, [2, 1, 3, 3], _rev("111_111_00")) == ("lit", [3, 3], 6)
    assert _decode(oracle, [2, 1, 3, 3], _rev("0_10_00000")) == ("lit", [1, 0], 3)
    l2 = [3, 3, 3, 3, 3, 2, 4, 4]
    assert _decode(oracle, l2, _rev("010_011_00")) == ("lit", [0, 1], 6)
    assert _decode(oracle, l2, _rev("00_00_0000")) == ("lit", [5, 5], 4)
    assert _decode(oracle, l2, _rev("1111_1110")) == ("lit", [7, 6], 8)


def test_huffman_secondary_table(oracle):
    # huffman.rs:425-480
    lop = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 15]
    assert _decode(oracle, lop, _rev("0_0_000000")) == ("lit", [0, 0], 2)
    assert _decode(oracle, lop, _rev("1110_1110")) == ("lit", [3, 3], 8)
    assert _decode(oracle, lop, _rev("1111_1111_1111_1110")) == ("sec", 15, 15)
    assert _decode(oracle, lop, _rev("1111_1111_1111_1111")) == ("sec", 15, 15)


def test_incomplete_codes_rejected(oracle):
    # huffman.rs:39-75 / SURVEY F9
    entries = oracle.const_table("litlen_table_entries", 288)
    assert not oracle.build_table([1, 0, 0], entries, 4096, False, True)[0]
    assert oracle.build_table([1, 0, 0], [], 512, True, False)[0]
    assert not oracle.build_table([2, 0, 0], [], 512, True, False)[0]
    assert not oracle.build_table([1, 1, 1], entries, 4096, False, True)[0]


def test_golden_zz(oracle):
    # decompress.rs:1331-1384
    name, d = cases.golden_streams()[-3]
    whole = oracle.decompress_by_chunks(d, [len(d)])
    bytewise = oracle.decompress_by_chunks(d, iter(lambda: 1, 0))
    assert whole == bytewise and whole[0] == "Ok"
    assert len(whole[1]) == 281 and oracle.adler32(whole[1]) == 751299 == zlib.adler32(whole[1])
    for name, d in cases.golden_streams()[-2:]:
        assert oracle.decompress_by_chunks(d, [len(d)])[0] == "BadLiteralLengthHuffmanTree"
        assert oracle.decompress_by_chunks(d, iter(lambda: 1, 0))[0] == "BadLiteralLengthHuffmanTree"


def test_fuzz_corpus(oracle):
    # fuzz/corpus/inflate replayed by the reference CI (rust.yml:81-85); expected output = zlib's
    n = 0
    for name, d in cases.golden_streams()[:-3]:
        st, out = oracle.decompress_to_vec(d)
        assert st == 0 and out == zlib.decompress(d), name
        n += 1
    assert n == 66


def test_unit_test_behaviours(oracle):
    # decompress.rs:1261-1280 ignore_adler32, :1282-1307 checksum_after_eof, :1309-1325 zero_length
    z = zlib.compress(b"Hello world!", 1)
    bad = z[:-1] + bytes([(z[-1] + 1) & 0xFF])
    assert oracle.decompress_to_vec(bad)[0] == oracle.STATUS["WrongChecksum"]
    assert oracle.decompress_to_vec(bad, oracle.IGNORE_ADLER32) == (0, b"Hello world!")
    d = oracle.Decompressor()
    out = np.zeros(1024, np.uint8)
    st, c, p = d.read(z[:-1], out, 0)
    assert (st, c, p) == (0, len(z) - 1, 12) and not d.is_done()
    st, c, p = d.read(z[-1:], out[:12], 12)
    assert (st, c, p) == (0, 1, 0) and d.is_done()
    empty = bytearray(zlib.compress(b"", 1))
    for _ in range(10):
        empty[2:2] = bytes([0, 0, 0, 0xFF, 0xFF])
    d = oracle.Decompressor()
    st, c, p = d.read(bytes(empty), np.zeros(0, np.uint8), 0)
    assert (st, c, p) == (0, len(empty), 0) and d.is_done()


def test_quirks(oracle):
    # fixed-block symbols 286/287 act as end-of-block (zlib rejects them)
    w = uf_craft.BitWriter()
    w.put(0x78, 8); w.put(0x01, 8)
    w.put(1, 1); w.put(1, 2)               # BFINAL=1, fixed
    w.put(int("00110000"[::-1], 2) + 0, 8)  # literal 0 (code 00110000)
    w.put(int("11000110"[::-1], 2), 8)      # symbol 286 (code 11000110)
    body = w.bytes() + zlib.adler32(b"\x00").to_bytes(4, "big")
    assert oracle.decompress_to_vec(body) == (0, b"\x00")
    with pytest.raises(zlib.error):
        zlib.decompress(body)
    # F10: trailing bytes are ignored
    z = zlib.compress(b"abc") + b"garbage"
    assert oracle.decompress_to_vec(z) == (0, b"abc")


def test_ultrafast_restatement_vectors(oracle):
    # SURVEY 8c restatement-derived vectors (regression anchors; parity unpinned by reference tests)
    hdr = bytes(T["ULTRAFAST_HEADER"])[:53]
    vec = [(b"", "ef1f0100000001"), (b"Hello world!", "ef8d3fe0c33ffca31efc491ff5b11ffefe0ff9471d09045e"),
           (bytes(7), "0f00f84700070001"), (bytes(8), "8f67fa4700080001"), (bytes(9) + b"\x01", "8f6742ff08000b0002"),
           (bytes(2048), "8fabaebaeaaaabaebaeaff9d7f0408000001")]
    for data, tail in vec:
        c = oracle.compress_ultra_fast(data)
        assert c[:53] == hdr and c[53:].hex() == tail
        assert zlib.decompress(c) == data and oracle.decompress_to_vec(c) == (0, data)


def test_ultrafast_roundtrips(oracle):
    # ultrafast.rs:201-224: "Hello world!", constant 2048-byte buffers, random 2048-byte buffers
    rng = random.Random(1)
    datas = [b"Hello world!"] + [bytes([v]) * 2048 for v in (0, 5, 128, 254)]
    datas += [bytes(rng.getrandbits(8) for _ in range(2048)) for _ in range(10)]
    datas += cases.compress_inputs(2, 20, [100, 3000, 20000])
    for d in datas:
        c = oracle.compress_ultra_fast(d)
        assert zlib.decompress(c) == d
        assert oracle.decompress_to_vec(c) == (0, d)


def test_ultrafast_run_rule(oracle):
    """the data-parallel run rule the CUDA encoder implements predicts the size of the
    sequential encoder's output exactly."""
    rng = random.Random(3)
    L = T["HUFFMAN_LENGTHS"]
    for _ in range(200):
        d = cases.sparse_bytes(rng, rng.choice([1, 8, 50, 300, 3000]))
        n8 = len(d) & ~7
        bits = 429 + 12
        i = 0
        run = 0

        def run_bits(R):
            b = 2 + 10 * ((R - 1) // 258)
            r = (R - 1) % 258
            if r > 4:
                sym = max(k for k, base in enumerate(uf_craft.LEN_BASE) if base <= r) + 257
                return b + L[sym] + uf_craft.LEN_EXTRA[sym - 257] + 1
            return b + 2 * r

        while i < len(d):
            c = i // 8
            is_run = False
            if i < n8 and d[i] == 0:
                tail_zero = all(v == 0 for v in d[i:8 * c + 8])
                head_zero = c > 0 and d[8 * c - 1] == 0 and all(v == 0 for v in d[8 * c:i + 1])
                is_run = tail_zero or head_zero
            if is_run:
                run += 1
            else:
                if run:
                    bits += run_bits(run)
                    run = 0
                bits += L[d[i]]
            i += 1
        if run:
            bits += run_bits(run)
        assert len(oracle.compress_ultra_fast(d)) == (bits + 7) // 8 + 4


def test_stored(oracle):
    # compress/mod.rs:241-268 incl. the empty-final-fixed-block quirk at multiples of 65535
    assert oracle.compress_stored(b"") == bytes.fromhex("7801030000000001")
    for n in (1, 65534, 65535, 65536, 131070, 131071):
        d = bytes((i * 7) & 0xFF for i in range(n))
        s = oracle.compress_stored(d)
        assert zlib.decompress(s) == d and oracle.decompress_to_vec(s) == (0, d)
        if n % 65535 == 0:
            assert s[-6:-4] == b"\x03\x00"


def test_differential_vs_zlib(oracle):
    # stand-in for the reference's miniz_oxide / flate2 differential fuzz targets
    rng = random.Random(7)
    for _ in range(150):
        data = cases.payload(rng, rng.randrange(6), rng.choice([0, 1, 5, 100, 1000, 5000, 70000]))
        z = cases.zlib_stream(rng, data)
        assert oracle.decompress_to_vec(z) == (0, data)
        st, out, cons = oracle.inflate_into(z, len(data))
        assert (st, out, cons) == (0, data, len(z))
        if data:
            st, out, _ = oracle.inflate_into(z, len(data) - 1)
            assert st == oracle.STATUS["OutputTooLarge"] and out == data[:-1]


def test_chunking_invariance(oracle):
    # fuzz targets inflate_bytewise3 / inflate_split: the result must not depend on input chunking
    rng = random.Random(9)
    for stream, cap in cases.mixed_zlib_cases(11, 12, [0, 5, 100, 1000]):
        if len(stream) > 3000:
            continue
        whole = oracle.decompress_by_chunks(stream, [len(stream)])
        bytewise = oracle.decompress_by_chunks(stream, iter(lambda: 1, 0))
        k = rng.randrange(1, 9)
        split = oracle.decompress_by_chunks(stream, iter(lambda: k, 0))
        assert whole == bytewise == split


def _adversarial_input(rng: random.Random, n: int) -> bytes:
    """bytes whose zero runs start, end and straddle 8-byte chunk edges, with lengths around the multiples of 258"""
    kind = rng.randrange(5)
    if kind == 0:
        return bytes(rng.choice([0, 0, 0, 1, 255, 7]) for _ in range(n))
    if kind == 1:
        return bytes(rng.getrandbits(8) for _ in range(n))
    out = bytearray()
    while len(out) < n:
        r = rng.random()
        if r < 0.45:
            out += bytes(rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 23, 24, 25, 256, 257, 258, 259, 260, 261, 262, 263,
                                     264, 265, 266, 267, 515, 516, 517, 518, 519, 520, 521, 522, 523, 524, 525,
                                     rng.randrange(1, 1200)]))
        elif r < 0.8:
            out += bytes(rng.choice([1, 255, 2, 254, 128, 0]) for _ in range(rng.randrange(1, 12)))
        else:
            out += bytes([rng.getrandbits(8) | 1]) * rng.randrange(1, 20)
    return bytes(out[:n])


def test_ultrafast_second_restatement(oracle):
    """the reference holds no vector for the ultra-fast encoder's bytes, so the C oracle is diffed
    against an independent pure-Python transliteration of ultrafast.rs (tests/uf_reference_py.py) on >= 10^4 random
    and adversarial inputs; every output must also inflate with zlib."""
    import uf_reference_py as P

    # the one table the transliteration derives itself: RFC 1951 3.2.5 against LENGTH_TO_SYMBOL / LENGTH_TO_LEN_EXTRA
    ls, le = oracle.const_table("length_to_symbol", 256), oracle.const_table("length_to_len_extra", 256)
    for length in range(3, 259):
        sym, extra, _ = P._length_symbol(length)
        assert (ls[length - 3], le[length - 3]) == (sym, extra)
    rng = random.Random(2024)
    count = 0
    for size_hi, reps in ((40, 5000), (300, 4000), (3000, 1200), (30000, 60)):
        for _ in range(reps):
            d = _adversarial_input(rng, rng.randrange(size_hi + 1))
            z = oracle.compress_ultra_fast(d)
            assert z == P.compress_calls([d]), f"restatements disagree on a {len(d)}-byte input"
            if count % 16 == 0:
                assert zlib.decompress(z) == d
            count += 1
    assert count >= 10000


def test_ultrafast_multi_call(oracle):
    """write_data call boundaries change the bytes (ultrafast.rs:97-99: the run counter and the 8-byte chunking
    restart with every call): the oracle's new / write_data / finish against the Python transliteration over random
    call patterns, and against the single-call output where the two must agree (cuts on 8-byte edges outside zero
    runs)."""
    import uf_reference_py as P

    rng = random.Random(77)
    differ = 0
    for _ in range(1500):
        d = _adversarial_input(rng, rng.randrange(1, 2500))
        cuts = sorted(rng.randrange(len(d) + 1) for _ in range(rng.randrange(0, 6)))
        calls = [d[a:b] for a, b in zip([0] + cuts, cuts + [len(d)])]
        z = oracle.compress_ultra_fast_calls(calls)
        assert z == P.compress_calls(calls)
        assert zlib.decompress(z) == d
        differ += z != oracle.compress_ultra_fast(d)
    assert differ > 100  # the call pattern really is visible in the output
    d = bytes(rng.choice([1, 2, 3]) for _ in range(4096))
    assert oracle.compress_ultra_fast_calls([d[:1024], d[1024:2048], d[2048:]]) == oracle.compress_ultra_fast(d)
    assert oracle.compress_ultra_fast_calls([]) == oracle.compress_ultra_fast(b"")
    assert oracle.compress_ultra_fast_calls([b"", b""]) == oracle.compress_ultra_fast(b"")


def test_adler32_against_zlib(oracle):
    """the blocked adler32 (the CPU baseline's checksum) is RFC 1950 arithmetic: zlib.adler32 on block edges"""
    rng = random.Random(1)
    for n in (0, 1, 31, 32, 33, 63, 64, 5535, 5536, 5537, 5551, 5552, 5553, 11072, 100003):
        for d in (bytes(rng.getrandbits(8) for _ in range(n)), bytes([255]) * n):
            assert oracle.adler32(d) == zlib.adler32(d)
            assert oracle.adler32(d, 0xfff0fff0 % 65521 | (65520 << 16)) == zlib.adler32(d, 0xfff0fff0 % 65521 | (65520 << 16))


def test_synth_tiles_match_product_generator(oracle, emul_lib):
    """bench.py's CPU legs generate their input with the oracle's own tile generator; it must be the generator the
    product uses on the device and on the host (csrc/synth.cuh), byte for byte"""
    from fdeflate_b200 import synth_tiles_host

    for first, n, w, h, seed in ((0, 3, 256, 256, 2024), (1000, 2, 1024, 37, 5), (7, 4, 5, 3, 1), (2 ** 40, 1, 300, 1, 99)):
        a = synth_tiles_host(first, n, w, h, seed, emul_lib).reshape(n, -1)
        assert np.array_equal(a, oracle.synth_tiles(first, n, w, h, seed, 3))
