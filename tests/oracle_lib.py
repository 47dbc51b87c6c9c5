"""ctypes binding of the CPU oracle (oracle/libfdeflate_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (fdeflate_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent.parent / "oracle"

STATUS_NAMES = [
    "Ok",
    "BadZlibHeader",
    "InsufficientInput",
    "InvalidBlockType",
    "InvalidUncompressedBlockLength",
    "InvalidHlit",
    "InvalidHdist",
    "InvalidCodeLengthRepeat",
    "BadCodeLengthHuffmanTree",
    "BadLiteralLengthHuffmanTree",
    "BadDistanceHuffmanTree",
    "InvalidLiteralLengthCode",
    "InvalidDistanceCode",
    "InputStartsWithRun",
    "DistanceTooFarBack",
    "WrongChecksum",
    "ExtraInput",
    "OutputTooLarge",
]
STATUS = {n: i for i, n in enumerate(STATUS_NAMES)}
IGNORE_ADLER32 = 1

_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)


def build(native: bool = False) -> Path:
    """Compile the oracle with gcc (building the checker is not using it)."""
    target = "native" if native else "all"
    subprocess.run(["make", "-C", str(ORACLE_DIR), target], check=True, capture_output=True)
    return ORACLE_DIR / ("libfdeflate_oracle_native.so" if native else "libfdeflate_oracle.so")


_LIBS: dict[bool, C.CDLL] = {}


def lib(native: bool = False) -> C.CDLL:
    if native in _LIBS:
        return _LIBS[native]
    path = ORACLE_DIR / ("libfdeflate_oracle_native.so" if native else "libfdeflate_oracle.so")
    src_m = max(os.path.getmtime(ORACLE_DIR / f) for f in ("fdeflate_oracle.c", "fdeflate_oracle.h", "png_filter_oracle.c"))
    if not path.exists() or os.path.getmtime(path) < src_m:
        build(native)
    L = C.CDLL(str(path))
    sz = C.c_size_t
    L.fdo_png_unfilter.restype = sz
    L.fdo_png_unfilter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    L.fdo_png_filter.restype = None
    L.fdo_png_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.fdo_adler32.restype = C.c_uint32
    L.fdo_adler32.argtypes = [C.c_uint32, C.c_void_p, sz]
    L.fdo_inflate_into.restype = C.c_int
    L.fdo_inflate_into.argtypes = [C.c_void_p, sz, C.c_void_p, sz, C.c_uint32, C.POINTER(sz), C.POINTER(sz)]
    L.fdo_decompress_to_vec.restype = C.c_int
    L.fdo_decompress_to_vec.argtypes = [C.c_void_p, sz, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(sz)]
    L.fdo_free.argtypes = [C.c_void_p]
    L.fdo_ultrafast_bound.restype = sz
    L.fdo_ultrafast_bound.argtypes = [sz]
    L.fdo_compress_ultra_fast.restype = sz
    L.fdo_compress_ultra_fast.argtypes = [C.c_void_p, sz, C.c_void_p, sz]
    L.fdo_stored_bound.restype = sz
    L.fdo_stored_bound.argtypes = [sz]
    L.fdo_compress_stored.restype = sz
    L.fdo_compress_stored.argtypes = [C.c_void_p, sz, C.c_void_p, sz]
    L.fdo_decompressor_new.restype = C.c_void_p
    L.fdo_decompressor_free.argtypes = [C.c_void_p]
    L.fdo_decompressor_ignore_adler32.argtypes = [C.c_void_p]
    L.fdo_decompressor_is_done.restype = C.c_int
    L.fdo_decompressor_is_done.argtypes = [C.c_void_p]
    L.fdo_decompressor_read.restype = C.c_int
    L.fdo_decompressor_read.argtypes = [C.c_void_p, C.c_void_p, sz, C.c_void_p, sz, sz, C.POINTER(sz), C.POINTER(sz)]
    L.fdo_build_table.restype = C.c_int
    L.fdo_build_table.argtypes = [C.c_void_p, sz, C.c_void_p, sz, C.c_void_p, C.c_void_p, sz, C.c_void_p,
                                  C.POINTER(sz), C.c_int, C.c_int]
    L.fdo_build_tables.restype = C.c_int
    L.fdo_build_tables.argtypes = [sz, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(sz), C.c_void_p,
                                   C.POINTER(sz)]
    for name, ty in [("fdo_huffman_lengths", C.c_uint8), ("fdo_huffman_codes", C.c_uint16),
                     ("fdo_ultrafast_header", C.c_uint8), ("fdo_litlen_table_entries", C.c_uint32),
                     ("fdo_distance_table_entries", C.c_uint32), ("fdo_length_to_symbol", C.c_uint16),
                     ("fdo_length_to_len_extra", C.c_uint8)]:
        getattr(L, name).restype = C.POINTER(ty)
    L.fdo_inflate_batch.restype = C.c_double
    L.fdo_inflate_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, sz, C.c_uint32, C.c_int]
    L.fdo_compress_ultra_fast_batch.restype = C.c_double
    L.fdo_compress_ultra_fast_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, sz, C.c_int]
    L.fdo_hardware_threads.restype = C.c_int
    L.fdo_ultrafast_new.restype = C.c_void_p
    L.fdo_ultrafast_new.argtypes = [C.c_void_p, sz]
    L.fdo_ultrafast_write_data.restype = None
    L.fdo_ultrafast_write_data.argtypes = [C.c_void_p, C.c_void_p, sz]
    L.fdo_ultrafast_finish.restype = sz
    L.fdo_ultrafast_finish.argtypes = [C.c_void_p]
    L.fdo_synth_tiles.restype = None
    L.fdo_synth_tiles.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int]
    _LIBS[native] = L
    return L


def _buf(b) -> tuple[C.c_void_p, int, object]:
    a = np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b)
    return C.c_void_p(a.ctypes.data if a.size else 0), a.size, a


def adler32(data, start: int = 1) -> int:
    p, n, keep = _buf(data)
    return lib().fdo_adler32(start, p, n)


def inflate_into(data, maxlen: int, flags: int = 0):
    """decompress_to_vec_bounded semantics. Returns (status, output_bytes, consumed)."""
    p, n, keep = _buf(data)
    out = np.zeros(max(maxlen, 1), dtype=np.uint8)
    olen = C.c_size_t(0)
    cons = C.c_size_t(0)
    st = lib().fdo_inflate_into(p, n, C.c_void_p(out.ctypes.data), maxlen, flags, C.byref(olen), C.byref(cons))
    return st, out[: olen.value].tobytes(), cons.value


def decompress_to_vec(data, flags: int = 0):
    """decompress_to_vec semantics. Returns (status, output_bytes or None)."""
    p, n, keep = _buf(data)
    outp = C.c_void_p()
    olen = C.c_size_t(0)
    st = lib().fdo_decompress_to_vec(p, n, flags, C.byref(outp), C.byref(olen))
    if st != 0:
        return st, None
    res = C.string_at(outp, olen.value)
    lib().fdo_free(outp)
    return st, res


def compress_ultra_fast(data) -> bytes:
    p, n, keep = _buf(data)
    cap = lib().fdo_ultrafast_bound(n)
    out = np.zeros(cap, dtype=np.uint8)
    m = lib().fdo_compress_ultra_fast(p, n, C.c_void_p(out.ctypes.data), cap)
    assert m > 0
    return out[:m].tobytes()


def compress_ultra_fast_calls(calls) -> bytes:
    """UltraFastCompressor::new, one write_data per element of `calls`, finish (ultrafast.rs:70-181)."""
    total = sum(len(c) for c in calls)
    cap = lib().fdo_ultrafast_bound(total) + 16 * (len(calls) + 1)
    out = np.zeros(cap, dtype=np.uint8)
    h = lib().fdo_ultrafast_new(C.c_void_p(out.ctypes.data), cap)
    for c in calls:
        p, n, _keep = _buf(c)
        lib().fdo_ultrafast_write_data(h, p, n)
    m = lib().fdo_ultrafast_finish(h)
    assert m > 0
    return out[:m].tobytes()


def synth_tiles(first_tile: int, n_tiles: int, width: int, height: int, seed: int, nthreads: int = 1, native: bool = False) -> np.ndarray:
    """the synthetic PNG-filtered tiles of SURVEY 8d, generated by the oracle (uint8 [n_tiles, height * (1 + 4 * width)])"""
    out = np.zeros((n_tiles, height * (1 + 4 * width)), dtype=np.uint8)
    lib(native).fdo_synth_tiles(C.c_void_p(out.ctypes.data), first_tile, n_tiles, width, height, seed, nthreads)
    return out


def compress_stored(data) -> bytes:
    p, n, keep = _buf(data)
    cap = lib().fdo_stored_bound(n)
    out = np.zeros(cap, dtype=np.uint8)
    m = lib().fdo_compress_stored(p, n, C.c_void_p(out.ctypes.data), cap)
    assert m > 0
    return out[:m].tobytes()


class Decompressor:
    """Mirror of the reference's streaming Decompressor (decompress.rs:96-342) over the oracle."""

    def __init__(self):
        self._d = lib().fdo_decompressor_new()

    def __del__(self):
        try:
            lib().fdo_decompressor_free(self._d)
        except Exception:
            pass

    def ignore_adler32(self):
        lib().fdo_decompressor_ignore_adler32(self._d)

    def is_done(self) -> bool:
        return bool(lib().fdo_decompressor_is_done(self._d))

    def read(self, data, output: np.ndarray, output_position: int):
        """Returns (status, consumed, produced); output is a writable uint8 numpy array."""
        p, n, keep = _buf(data)
        cons = C.c_size_t(0)
        prod = C.c_size_t(0)
        st = lib().fdo_decompressor_read(self._d, p, n, C.c_void_p(output.ctypes.data if output.size else 0),
                                         output.size, output_position, C.byref(cons), C.byref(prod))
        return st, cons.value, prod.value


def decompress_by_chunks(data: bytes, chunks, out_size: int = 1_000_000):
    """Restatement of src/decompress/tests/test_utils.rs:47-87 (adler ignored, big fixed output)."""
    d = Decompressor()
    d.ignore_adler32()
    out = np.zeros(out_size, dtype=np.uint8)
    in_pos = out_pos = 0
    it = iter(chunks)
    iterations = 0
    while not d.is_done():
        iterations += 1
        if iterations > 5000:
            return "TooManyIterations", None
        chunk = next(it, 0)
        end = min(in_pos + chunk, len(data))
        st, c, p = d.read(data[in_pos:end], out, out_pos)
        if st != 0:
            return STATUS_NAMES[st], None
        in_pos += c
        out_pos += p
        if out_pos == out.size and c == 0 and not d.is_done():
            return "OutputTooLarge", None
    return "Ok", out[:out_pos].tobytes()


def build_table(lengths, entries, primary_size: int, is_distance: bool, double_literal: bool):
    """huffman.rs:18-184. Returns (ok, primary u32 array, secondary u16 array, codes u16 array)."""
    lengths = np.asarray(lengths, dtype=np.uint8)
    entries = np.asarray(entries, dtype=np.uint32)
    codes = np.zeros(max(len(lengths), 1), dtype=np.uint16)
    primary = np.zeros(primary_size, dtype=np.uint32)
    secondary = np.zeros(4096, dtype=np.uint16)
    slen = C.c_size_t(0)
    ok = lib().fdo_build_table(C.c_void_p(lengths.ctypes.data), len(lengths),
                               C.c_void_p(entries.ctypes.data if entries.size else 0), entries.size,
                               C.c_void_p(codes.ctypes.data), C.c_void_p(primary.ctypes.data), primary_size,
                               C.c_void_p(secondary.ctypes.data), C.byref(slen), int(is_distance),
                               int(double_literal))
    return bool(ok), primary, secondary[: slen.value], codes


def build_tables(hlit: int, code_lengths):
    """decompress.rs:561-606. Returns (status, litlen u32[4096], dist u32[512], secondary, dist_secondary)."""
    cl = np.asarray(code_lengths, dtype=np.uint8)
    assert cl.size == 320
    lit = np.zeros(4096, dtype=np.uint32)
    dist = np.zeros(512, dtype=np.uint32)
    sec = np.zeros(4096, dtype=np.uint16)
    dsec = np.zeros(4096, dtype=np.uint16)
    n1 = C.c_size_t(0)
    n2 = C.c_size_t(0)
    st = lib().fdo_build_tables(hlit, C.c_void_p(cl.ctypes.data), C.c_void_p(lit.ctypes.data),
                                C.c_void_p(dist.ctypes.data), C.c_void_p(sec.ctypes.data), C.byref(n1),
                                C.c_void_p(dsec.ctypes.data), C.byref(n2))
    return st, lit, dist, sec[: n1.value], dsec[: n2.value]


def const_table(name: str, n: int) -> np.ndarray:
    p = getattr(lib(), "fdo_" + name)()
    return np.ctypeslib.as_array(p, shape=(n,)).copy()


def hardware_threads() -> int:
    return lib().fdo_hardware_threads()


def inflate_batch(in_base: np.ndarray, in_off, in_len, out_base: np.ndarray, out_off, out_cap, flags=0,
                  nthreads=1, native=False):
    """Multi-threaded CPU batch inflate. Returns (seconds, out_len u64[n], status i32[n])."""
    n = len(in_off)
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    in_len = np.ascontiguousarray(in_len, dtype=np.uint64)
    out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
    out_cap = np.ascontiguousarray(out_cap, dtype=np.uint64)
    out_len = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    secs = lib(native).fdo_inflate_batch(in_base.ctypes.data, in_off.ctypes.data, in_len.ctypes.data,
                                         out_base.ctypes.data, out_off.ctypes.data, out_cap.ctypes.data,
                                         out_len.ctypes.data, status.ctypes.data, n, flags, nthreads)
    return secs, out_len, status


def compress_ultra_fast_batch(in_base: np.ndarray, in_off, in_len, out_base: np.ndarray, out_off, out_cap,
                              nthreads=1, native=False):
    n = len(in_off)
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    in_len = np.ascontiguousarray(in_len, dtype=np.uint64)
    out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
    out_cap = np.ascontiguousarray(out_cap, dtype=np.uint64)
    out_len = np.zeros(n, dtype=np.uint64)
    secs = lib(native).fdo_compress_ultra_fast_batch(in_base.ctypes.data, in_off.ctypes.data, in_len.ctypes.data,
                                                     out_base.ctypes.data, out_off.ctypes.data, out_cap.ctypes.data,
                                                     out_len.ctypes.data, n, nthreads)
    return secs, out_len


# ---- PNG row filters (oracle/png_filter_oracle.c; PNG specification section 9) -------------------
def png_unfilter(filtered, h: int, stride: int, bpp: int):
    """-> (bad_row, raw bytes): bad_row = 0, or 1 + index of the first row with a filter type > 4"""
    src = np.frombuffer(bytes(filtered), dtype=np.uint8)
    assert src.size == h * (1 + stride)
    raw = np.zeros(h * stride, dtype=np.uint8)
    bad = lib().fdo_png_unfilter(raw.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), h, stride, bpp)
    return int(bad), raw.tobytes()


def png_filter(raw, h: int, stride: int, bpp: int, mode: int) -> bytes:
    src = np.frombuffer(bytes(raw), dtype=np.uint8)
    assert src.size == h * stride
    out = np.zeros(h * (1 + stride), dtype=np.uint8)
    lib().fdo_png_filter(out.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), h, stride, bpp, mode)
    return out.tobytes()
