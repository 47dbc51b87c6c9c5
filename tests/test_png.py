"""PNG row filters and the PNG image-data path.

CPU part: the oracle (oracle/png_filter_oracle.c, PNG specification section 9) is pinned against golden PNG files
written by two independent encoders -- Pillow and OpenCV/libpng (tools/make_png_golden.py, tests/golden/png) --
whose decoded pixels are recorded in manifest.json; the kernels run on the SIMT emulator against the oracle.
GPU part (-m gpu): the same through libfdeflate_b200.so, plus files written by our encoder decoded by Pillow when
it is installed."""
import hashlib
import json
import random
import zlib
from pathlib import Path

import numpy as np
import pytest

import cases

GOLD = Path(__file__).resolve().parent / "golden" / "png"
MANIFEST = json.loads((GOLD / "manifest.json").read_text())


def _pixels_hash(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_oracle_unfilter_matches_independent_decoders(oracle):
    """zlib (system) inflates the IDAT stream, the oracle undoes the filters: the pixels must be the ones Pillow /
    OpenCV decoded from the same file.  All five filter types occur in the fixtures."""
    from fdeflate_b200 import png

    types = set()
    for name, want in sorted(MANIFEST.items()):
        info, z = png.parse((GOLD / name).read_bytes())
        filtered = zlib.decompress(z)
        assert len(filtered) == info.height * (1 + info.stride)
        types |= {filtered[y * (1 + info.stride)] for y in range(info.height)}
        bad, raw = oracle.png_unfilter(filtered, info.height, info.stride, info.bpp)
        assert bad == 0
        a = png._to_array(info, np.frombuffer(raw, dtype=np.uint8))
        assert list(a.shape) == want["shape"] and str(a.dtype) == want["dtype"], name
        assert _pixels_hash(a) == want["sha256"], name
    assert types == {0, 1, 2, 3, 4}


def _random_images(seed, count):
    rng = random.Random(seed)
    out = []
    for k in range(count):
        bpp = rng.choice((1, 2, 3, 4, 6, 8))
        w = rng.choice((1, 2, 3, 7, 31, 32, 33, 100))
        h = rng.choice((1, 2, 5, 31, 32, 33, 40))
        stride = w * bpp if k % 5 else max(1, w * bpp - rng.randrange(bpp))  # (sub-byte depths: stride % bpp != 0)
        kind = k % 3
        if kind == 0:
            raw = bytes(rng.getrandbits(8) for _ in range(h * stride))
        elif kind == 1:
            raw = bytes((3 * (i % stride) + 5 * (i // stride)) & 0xff for i in range(h * stride))
        else:
            raw = bytes(rng.choice((0, 0, 0, 1, 255, 7)) for _ in range(h * stride))
        out.append((raw, (h, stride, bpp)))
    return out


def test_oracle_filter_roundtrip_all_modes(oracle):
    for raw, (h, stride, bpp) in _random_images(1, 40):
        for mode in range(6):
            f = oracle.png_filter(raw, h, stride, bpp, mode)
            assert len(f) == h * (1 + stride)
            if mode < 5:
                assert all(f[y * (1 + stride)] == mode for y in range(h))
            assert oracle.png_unfilter(f, h, stride, bpp) == (0, raw)


def _check_filters(ctx, oracle, images):
    geo = [g for _, g in images]
    for mode in range(6):
        st, outs = ctx.png_filter_batch([r for r, _ in images], geo, mode)
        assert (st == 0).all()
        for (raw, (h, s, b)), o in zip(images, outs):
            assert o == oracle.png_filter(raw, h, s, b, mode), f"filter mode {mode} differs from the oracle ({h}x{s} bpp {b})"
        st, back = ctx.png_unfilter_batch(outs, geo)
        assert (st == 0).all()
        assert back == [r for r, _ in images]
    # a filter-type byte out of range
    raw, (h, s, b) = images[3]
    f = bytearray(oracle.png_filter(raw, h, s, b, 1))
    f[(h - 1) * (1 + s)] = 5
    st, _ = ctx.png_unfilter_batch([bytes(f)], [(h, s, b)])
    assert st[0] == 19
    st, _ = ctx.png_unfilter_batch([bytes(f)], [(h, s, 9)])
    assert st[0] == 20


@pytest.mark.emul
def test_filter_kernels_on_emulator(emul_ctx, oracle):
    _check_filters(emul_ctx, oracle, _random_images(2, 24))


def _golden_files():
    return [(name, (GOLD / name).read_bytes()) for name in sorted(MANIFEST)]


def _check_decode_golden(ctx, files):
    from fdeflate_b200 import png

    arrays = png.decode_batch([d for _, d in files], ctx)
    for (name, _), a in zip(files, arrays):
        want = MANIFEST[name]
        assert list(a.shape) == want["shape"] and str(a.dtype) == want["dtype"], name
        assert _pixels_hash(a) == want["sha256"], name
    return arrays


@pytest.mark.emul
def test_png_decode_golden_on_emulator(emul_ctx):
    _check_decode_golden(emul_ctx, _golden_files()[::4])


def _check_encode_roundtrip(ctx, arrays):
    from fdeflate_b200 import png

    for mode in (0, 4, 5):
        files = png.encode_batch(arrays, ctx, filter_mode=mode)
        for f, a in zip(files, arrays):
            info, z = png.parse(f)
            assert z[:2] == b"\x78\x01" and zlib.decompress(z)[0] == (mode if mode < 5 else zlib.decompress(z)[0])
        back = png.decode_batch(files, ctx)
        for a, b in zip(arrays, back):
            assert a.dtype == b.dtype and np.array_equal(a, b)
    return files


@pytest.mark.emul
def test_png_encode_roundtrip_on_emulator(emul_ctx):
    rng = np.random.default_rng(3)
    arrays = [rng.integers(0, 256, (9, 13, 4), dtype=np.uint8), rng.integers(0, 256, (33, 5), dtype=np.uint8),
              rng.integers(0, 65536, (4, 6, 3), dtype=np.uint16), np.zeros((40, 40, 3), dtype=np.uint8)]
    _check_encode_roundtrip(emul_ctx, arrays)


def _check_crc32(ctx):
    rng = random.Random(9)
    items = [b"", b"a", b"123456789", bytes(2047), bytes(rng.getrandbits(8) for _ in range(2048)),
             bytes(rng.getrandbits(8) for _ in range(2049)), bytes(rng.getrandbits(8) for _ in range(70001)), bytes(300000)]
    got = ctx.crc32_batch(items)
    assert int(got[2]) == 0xcbf43926  # the check value of CRC-32/ISO-HDLC
    assert [int(x) for x in got] == [zlib.crc32(i) for i in items]
    seed = zlib.crc32(b"IDAT")
    assert [int(x) for x in ctx.crc32_batch(items, seed)] == [zlib.crc32(i, seed) for i in items]


def _check_crc_detection(ctx):
    from fdeflate_b200 import png

    name = max(MANIFEST, key=lambda k: (GOLD / k).stat().st_size)
    good = (GOLD / name).read_bytes()
    png.decode_batch([good], ctx, crc="device")
    bad = bytearray(good)
    bad[good.index(b"IDAT") + 40] ^= 4  # inside the IDAT payload
    for mode in ("device", "host"):
        with pytest.raises(png.PngError, match="bad CRC"):
            png.decode_batch([good, bytes(bad)], ctx, crc=mode)


@pytest.mark.emul
def test_crc32_on_emulator(emul_ctx):
    _check_crc32(emul_ctx)
    _check_crc_detection(emul_ctx)


@pytest.mark.gpu
def test_crc32_on_gpu(gpu_ctx):
    _check_crc32(gpu_ctx)
    _check_crc_detection(gpu_ctx)
    rng = np.random.default_rng(2)
    items = [rng.integers(0, 256, int(n), dtype=np.uint8).tobytes() for n in rng.integers(0, 3_000_000, 40)]
    assert [int(x) for x in gpu_ctx.crc32_batch(items)] == [zlib.crc32(i) for i in items]


def _check_decode_files(ctx):
    """the library's own container walk: same pixels as the recorded ones; broken files get the right status"""
    from fdeflate_b200 import png

    files = _golden_files()
    arrays = png.decode_files_batch([d for _, d in files], ctx)
    for (name, _), a in zip(files, arrays):
        want = MANIFEST[name]
        assert list(a.shape) == want["shape"] and str(a.dtype) == want["dtype"] and _pixels_hash(a) == want["sha256"], name
    # a file cut into many small IDAT chunks (what libpng writes) must give the same pixels: gather path
    name, good = max(files, key=lambda f: len(f[1]))
    info, z = png.parse(good)
    ihdr_end = 8 + 12 + 13
    pieces = [z[i:i + 97] for i in range(0, len(z), 97)] + [b""]
    multi = good[:ihdr_end] + b"".join(png._chunk(b"IDAT", p) for p in pieces) + png._chunk(b"tEXt", b"k\0v") + png._chunk(b"IEND", b"")
    a = png.decode_files_batch([multi, good], ctx)
    assert _pixels_hash(a[0]) == MANIFEST[name]["sha256"] and _pixels_hash(a[1]) == MANIFEST[name]["sha256"]
    # statuses through the C ABI
    import numpy as np
    from fdeflate_b200.api import _ptr

    bad_crc = bytearray(good); bad_crc[good.index(b"IDAT") + 40] ^= 4
    interlaced = bytearray(good); interlaced[8 + 8 + 12] = 1
    import binascii, struct
    interlaced[8 + 8 + 13:8 + 8 + 17] = struct.pack(">I", binascii.crc32(bytes(interlaced[12:8 + 8 + 13])))
    split_idat = good[:ihdr_end] + png._chunk(b"IDAT", z[:50]) + png._chunk(b"tEXt", b"k\0v") + png._chunk(b"IDAT", z[50:]) + png._chunk(b"IEND", b"")
    cases_ = [good, bytes(bad_crc), bytes(interlaced), b"not a png at all, but long enough to look at", good[:-12], split_idat,
              good[:ihdr_end] + png._chunk(b"IEND", b"")]
    base, off, lens = ctx._pack(cases_, align=1)
    n = len(cases_)
    w, h, d, c, s_ = (np.zeros(n, dtype=np.uint32) for _ in range(5))
    st = np.zeros(n, dtype=np.int32)
    assert ctx.lib.L.fdb_png_probe_batch(_ptr(base), _ptr(off), _ptr(lens), _ptr(w), _ptr(h), _ptr(d), _ptr(c), _ptr(s_), _ptr(st), n) == 0
    assert list(st) == [0, 0, 23, 21, 21, 21, 21]
    raw_off = np.arange(n, dtype=np.uint64) * np.uint64(int(h[0]) * int(s_[0]) + 64)
    raw = np.zeros(int(raw_off[-1]) + int(h[0]) * int(s_[0]) + 64, dtype=np.uint8)
    raw_cap = np.full(n, int(h[0]) * int(s_[0]), dtype=np.uint64)
    assert ctx.lib.L.fdb_png_decode_files_batch(ctx._h, _ptr(base), _ptr(off), _ptr(lens), _ptr(raw), _ptr(raw_off), _ptr(raw_cap), _ptr(st), n) == 0
    assert list(st) == [0, 22, 23, 21, 21, 21, 21]
    # a slot one byte short (or a header that claims more pixels than the slot holds): OutputTooLarge for that file only,
    # and nothing is written to its slot
    raw[:] = 0xAB
    raw_cap[0] -= 1
    assert ctx.lib.L.fdb_png_decode_files_batch(ctx._h, _ptr(base), _ptr(off), _ptr(lens), _ptr(raw), _ptr(raw_off), _ptr(raw_cap), _ptr(st), n) == 0
    assert list(st) == [17, 22, 23, 21, 21, 21, 21]
    assert (raw[: int(raw_off[1])] == 0xAB).all()
    huge = bytearray(good); huge[16:24] = struct.pack(">II", 60000, 60000)  # a header that claims 14 GB of pixels
    huge[29:33] = struct.pack(">I", binascii.crc32(bytes(huge[12:29])))
    base2, off2, lens2 = ctx._pack([bytes(huge), good], align=1)
    st2 = np.zeros(2, dtype=np.int32)
    cap2 = np.full(2, int(h[0]) * int(s_[0]), dtype=np.uint64)
    assert ctx.lib.L.fdb_png_decode_files_batch(ctx._h, _ptr(base2), _ptr(off2), _ptr(lens2), _ptr(raw), _ptr(raw_off), _ptr(cap2), _ptr(st2), 2) == 0
    assert list(st2) == [17, 0]


@pytest.mark.emul
def test_png_decode_files_on_emulator(emul_ctx):
    _check_decode_files(emul_ctx)


@pytest.mark.gpu
def test_png_decode_files_on_gpu(gpu_ctx):
    _check_decode_files(gpu_ctx)


def _check_encode_files(ctx, arrays):
    """files written by the library's own framing: the same bytes as the Python framing, and they decode back"""
    from fdeflate_b200 import png

    for mode in (1, 5):
        files = png.encode_files_batch(arrays, ctx, filter_mode=mode)
        assert files == png.encode_batch(arrays, ctx, filter_mode=mode)
        for f in files:
            png.parse(f)  # signature, chunk order, every CRC (host check)
        for a, b in zip(arrays, png.decode_files_batch(files, ctx)):
            assert a.dtype == b.dtype and np.array_equal(a, b)
    return files


@pytest.mark.emul
def test_png_encode_files_on_emulator(emul_ctx):
    rng = np.random.default_rng(8)
    _check_encode_files(emul_ctx, [rng.integers(0, 256, (9, 13, 4), dtype=np.uint8), rng.integers(0, 256, (33, 5), dtype=np.uint8),
                                   rng.integers(0, 65536, (4, 6, 3), dtype=np.uint16), np.zeros((40, 40, 2), dtype=np.uint8)])


@pytest.mark.gpu
def test_png_encode_files_on_gpu(gpu_ctx):
    rng = np.random.default_rng(8)
    y, x = np.mgrid[0:300, 0:517]
    photo = np.stack([(x + y) % 256, (2 * x) % 256, (x * y >> 5) % 256], -1).astype(np.uint8)
    arrays = [rng.integers(0, 256, (9, 13, 4), dtype=np.uint8), photo, rng.integers(0, 65536, (40, 60), dtype=np.uint16),
              rng.integers(0, 256, (700, 900, 4), dtype=np.uint8), np.zeros((400, 400, 3), dtype=np.uint8)] + \
             [rng.integers(0, 8, (64, 64, 4), dtype=np.uint8) for _ in range(300)]
    files = _check_encode_files(gpu_ctx, arrays)
    try:
        import io

        from PIL import Image
    except ImportError:
        return
    for f, a in list(zip(files, arrays))[:8]:
        if a.dtype == np.uint8:
            assert np.array_equal(np.asarray(Image.open(io.BytesIO(f))), a)


def test_png_container_errors():
    from fdeflate_b200 import png

    good = (GOLD / sorted(MANIFEST)[0]).read_bytes()
    with pytest.raises(png.PngError):
        png.parse(b"not a png")
    with pytest.raises(png.PngError):
        png.parse(good[:-5])
    bad = bytearray(good)
    bad[40] ^= 1  # inside a chunk: its CRC no longer matches
    with pytest.raises(png.PngError):
        png.parse(bytes(bad))


# ---- GPU ------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_filter_kernels_on_gpu(gpu_ctx, oracle):
    _check_filters(gpu_ctx, oracle, _random_images(2, 120))
    # tile-sized images, every mode, against the oracle
    from fdeflate_b200 import synth_tiles_host

    rng = np.random.default_rng(4)
    big = [(rng.integers(0, 256, 256 * 1024, dtype=np.uint8).tobytes(), (256, 1024, 4)),
           (bytes(300 * 999), (300, 999, 3)), (rng.integers(0, 3, 128 * 4096, dtype=np.uint8).tobytes(), (128, 4096, 8))]
    _check_filters(gpu_ctx, oracle, big + _random_images(5, 4))
    tiles = synth_tiles_host(0, 8, 256, 256, 1, gpu_ctx.lib)  # already PNG-filtered rows (Sub / Paeth)
    st, raws = gpu_ctx.png_unfilter_batch([t.tobytes() for t in tiles], [(256, 1024, 4)] * 8)
    assert (st == 0).all()
    for t, r in zip(tiles, raws):
        assert oracle.png_unfilter(t.tobytes(), 256, 1024, 4) == (0, r)


@pytest.mark.gpu
def test_png_decode_golden_on_gpu(gpu_ctx):
    _check_decode_golden(gpu_ctx, _golden_files())


@pytest.mark.gpu
def test_png_encode_roundtrip_on_gpu(gpu_ctx):
    rng = np.random.default_rng(3)
    y, x = np.mgrid[0:300, 0:517]
    photo = np.stack([(x + y) % 256, (2 * x) % 256, (x * y >> 5) % 256, np.full_like(x, 255)], -1).astype(np.uint8)
    arrays = [rng.integers(0, 256, (9, 13, 4), dtype=np.uint8), rng.integers(0, 256, (33, 5), dtype=np.uint8),
              rng.integers(0, 65536, (40, 60, 3), dtype=np.uint16), np.zeros((400, 400, 3), dtype=np.uint8), photo,
              rng.integers(0, 256, (1024, 1024, 4), dtype=np.uint8), photo[..., :2].copy()]
    files = _check_encode_roundtrip(gpu_ctx, arrays)
    try:
        import io

        from PIL import Image
    except ImportError:
        return
    for f, a in zip(files, arrays):
        if a.dtype == np.uint8:  # an independent decoder reads what we wrote
            assert np.array_equal(np.asarray(Image.open(io.BytesIO(f))), a)


def _encode_streams(ctx, raws, geometry, mode):
    """fdb_png_encode_batch -> the zlib stream of every image (statuses must be Ok)."""
    from fdeflate_b200.api import _ptr

    n = len(raws)
    raw_base, raw_off, _ = ctx._pack(raws)
    h = np.array([g[0] for g in geometry], dtype=np.uint32)
    s = np.array([g[1] for g in geometry], dtype=np.uint32)
    b = np.array([g[2] for g in geometry], dtype=np.uint32)
    caps = np.array([ctx.ultrafast_bound(int(hh) * (1 + int(ss))) for hh, ss in zip(h, s)], dtype=np.uint64)
    out_off = np.zeros(n, dtype=np.uint64)
    out_off[1:] = np.cumsum(caps[:-1])
    out = np.zeros(int(out_off[-1] + caps[-1]), dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    rc = ctx.lib.L.fdb_png_encode_batch(ctx._h, _ptr(raw_base), _ptr(raw_off), _ptr(h), _ptr(s), _ptr(b), mode, _ptr(out),
                                        _ptr(out_off), _ptr(caps), _ptr(out_len), _ptr(status), n)
    assert rc == 0 and (status == 0).all(), (rc, list(status))
    return [out[int(out_off[i]): int(out_off[i]) + int(out_len[i])].tobytes() for i in range(n)]


def _check_fused_filter_deflate(ctx, oracle, big: bool):
    """PNG encode with one filter type per image (modes 0..4) computes the filtered bytes inside the encoder
    (deflate_png.cuh): the stream must be the oracle's ultra-fast deflate of the oracle's filtered image, byte for byte,
    for row lengths that put the type bytes and row wraps at every position of the encoder's 16-byte loads."""
    rng = random.Random(11)
    geometry = [(1, 1, 1), (3, 2, 1), (5, 3, 3), (7, 14, 2), (4, 15, 3), (6, 16, 4), (9, 17, 1), (3, 63, 3), (5, 64, 4), (4, 65, 8),
                (2, 100, 6), (33, 31, 4), (1, 5000, 4), (17, 256, 8)]
    if big:
        geometry += [(256, 1024, 4), (100, 3001, 3), (64, 2048, 8), (300, 700, 2), (1024, 1, 1)]
    raws = []
    for k, (h, s, bpp) in enumerate(geometry):
        if k % 3 == 0:
            raws.append(cases.sparse_bytes(rng, h * s))  # zero runs across rows
        elif k % 3 == 1:
            raws.append(bytes(rng.getrandbits(8) for _ in range(h * s)))
        else:  # smooth: small residuals after filtering
            raws.append(bytes((3 * (i % s) + 7 * (i // s) + rng.randrange(3)) & 0xFF for i in range(h * s)))
    for mode in range(5):
        got = _encode_streams(ctx, raws, geometry, mode)
        for i, (h, s, bpp) in enumerate(geometry):
            want = oracle.compress_ultra_fast(oracle.png_filter(raws[i], h, s, bpp, mode))
            assert got[i] == want, "mode %d image %d (%d rows of %d, bpp %d)" % (mode, i, h, s, bpp)


@pytest.mark.emul
def test_png_encode_fused_filter_on_emulator(emul_lib, oracle, monkeypatch):
    import fdeflate_b200 as F

    monkeypatch.setenv("FDB_PNG_FUSED", "1")  # read when a context is created (the default is the two-kernel path)
    _check_fused_filter_deflate(F.Context(0, emul_lib), oracle, big=False)


@pytest.mark.gpu
def test_png_encode_fused_filter_on_gpu(gpu_ctx, oracle, monkeypatch):
    import fdeflate_b200 as F

    monkeypatch.setenv("FDB_PNG_FUSED", "1")
    fused = F.Context(0)
    _check_fused_filter_deflate(fused, oracle, big=True)
    # and the two-kernel path (filter kernel, then the encoder: the default) gives the same streams
    rng = random.Random(12)
    geometry = [(64, 1024, 4), (37, 333, 3), (5, 16, 4)]
    raws = [cases.sparse_bytes(rng, h * s) for h, s, _ in geometry]
    for mode in range(5):
        assert _encode_streams(fused, raws, geometry, mode) == _encode_streams(gpu_ctx, raws, geometry, mode)
