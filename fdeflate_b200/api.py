"""Host-side mirror of the reference crate's public API (reference src/lib.rs:29-36) over the C ABI.

Names, argument meaning and error behaviour follow image-rs/fdeflate 0.4.0-dev:

    decompress_to_vec / decompress_to_vec_bounded      src/decompress.rs:1079, :1111
    Decompressor {new, read, is_done, ignore_adler32}  src/decompress.rs:96-342
    DecompressionError / BoundedDecompressionError     src/decompress.rs:13-48, :1090-1102
    compress_to_vec_ultra_fast, UltraFastCompressor    src/compress/mod.rs:313, src/compress/ultrafast.rs:9-181
    Compressor::new(w, 0, zlib)  ("stored")            src/compress/mod.rs:69-101, :241-268

plus the batch entry points this project adds (Context.inflate_batch, ...).  All compute happens in
the CUDA library; this file only packs buffers and maps status codes to exceptions.  The Rust shim
a maintainer would write against the same C ABI is shown in INTEGRATION.md (no Rust toolchain exists
in this image, so the host side is mirrored in Python and C++ instead).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from . import _native
from ._native import FLAG_GENERAL_ONLY, FLAG_IGNORE_ADLER32, NativeLib

STATUS_NAMES = [
    "Ok", "BadZlibHeader", "InsufficientInput", "InvalidBlockType", "InvalidUncompressedBlockLength",
    "InvalidHlit", "InvalidHdist", "InvalidCodeLengthRepeat", "BadCodeLengthHuffmanTree",
    "BadLiteralLengthHuffmanTree", "BadDistanceHuffmanTree", "InvalidLiteralLengthCode", "InvalidDistanceCode",
    "InputStartsWithRun", "DistanceTooFarBack", "WrongChecksum", "ExtraInput", "OutputTooLarge",
    "OutputBufferTooSmall", "PngBadFilterType", "PngBadGeometry", "PngBadFile", "PngBadCrc", "PngUnsupported",
]
ST_OK, ST_INSUFFICIENT_INPUT, ST_OUTPUT_TOO_LARGE = 0, 2, 17


class DecompressionError(Exception):
    """reference src/decompress.rs:13-48; `.kind` is the variant name, `.code` its 1-based index."""

    def __init__(self, code: int):
        self.code = int(code)
        self.kind = STATUS_NAMES[self.code] if 0 <= self.code < len(STATUS_NAMES) else f"Unknown({code})"
        super().__init__(self.kind)

    def __eq__(self, other):
        return isinstance(other, DecompressionError) and other.code == self.code

    def __hash__(self):
        return hash(self.code)


class BoundedDecompressionError(Exception):
    """reference src/decompress.rs:1090-1102: DecompressionError{inner} | OutputTooLarge{partial_output}."""

    def __init__(self, inner: DecompressionError | None = None, partial_output: bytes | None = None):
        self.inner = inner
        self.partial_output = partial_output
        self.kind = "OutputTooLarge" if inner is None else "DecompressionError"
        super().__init__(self.kind if inner is None else f"DecompressionError({inner.kind})")


class FdbError(RuntimeError):
    """batch-level failure (CUDA error / bad argument) reported by the C ABI"""


def _align16(x: int) -> int:
    return (x + 15) & ~15


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data if a.size else 0)


class Context:
    """One fdb_ctx: bound to one GPU, not re-entrant (include/fdeflate_b200.h)."""

    def __init__(self, device: int = 0, lib: NativeLib | None = None):
        self.lib = lib if lib is not None else _native.default_lib()
        h = C.c_void_p()
        rc = self.lib.L.fdb_create(device, C.byref(h))
        if rc != 0 or not h:
            raise FdbError(f"fdb_create(device={device}) failed with code {rc}: no usable CUDA device "
                           f"(fdeflate_b200 has no CPU fallback)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self.lib.L.fdb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise FdbError(f"{what} failed ({rc}): {self.lib.L.fdb_last_error(self._h).decode()}")

    @property
    def launch_count(self) -> int:
        return int(self.lib.L.fdb_launch_count(self._h))

    def set_pipeline_chunk(self, nbytes: int = 0):
        """slot span per chunk of the host-buffer pipeline (0 = default 128 MiB); a tuning knob"""
        self._check(self.lib.L.fdb_set_pipeline_chunk(self._h, nbytes), "fdb_set_pipeline_chunk")

    def last_general_count(self, stream: int = 0) -> int:
        """streams of the last inflate batch that the fast path handed to the general kernel"""
        return int(self.lib.L.fdb_last_general_count(self._h, stream))

    # ---- raw packed host-buffer calls ---------------------------------------------------------
    def set_split_large(self, on: bool = True):
        """device-pointer calls: decode / encode long streams with many warps each (see fdeflate_b200.h)"""
        self._check(self.lib.L.fdb_set_split_large(self._h, 1 if on else 0), "fdb_set_split_large")

    def set_split_threshold(self, inflate_stream_bytes: int = 0, deflate_input_bytes: int = 0):
        """sizes from which a stream is decoded / encoded by many warps (0 = default: 128 KiB / 256 KiB, host-buffer deflate calls 1 MiB)"""
        self._check(self.lib.L.fdb_set_split_threshold(self._h, inflate_stream_bytes, deflate_input_bytes),
                    "fdb_set_split_threshold")

    def set_split_scratch(self, nbytes: int):
        """size of the lane-record pool of span-by-span device-pointer inflate calls (fdb_set_split_scratch)"""
        self._check(self.lib.L.fdb_set_split_scratch(self._h, nbytes), "fdb_set_split_scratch")

    def last_split_spans(self, stream: int = 0) -> int:
        """spans the long streams of the most recent inflate batch were cut into (0 = one warp per stream)"""
        return int(self.lib.L.fdb_last_split_spans(self._h, stream))

    def inflate_packed(self, in_base: np.ndarray, in_off, in_len, out_base: np.ndarray, out_off, out_cap,
                       flags: int = 0):
        n = len(in_off)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        in_len = np.ascontiguousarray(in_len, dtype=np.uint64)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        out_cap = np.ascontiguousarray(out_cap, dtype=np.uint64)
        out_len = np.zeros(n, dtype=np.uint64)
        consumed = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        rc = self.lib.L.fdb_inflate_batch(self._h, _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out_base),
                                          _ptr(out_off), _ptr(out_cap), _ptr(out_len), _ptr(consumed), _ptr(status),
                                          n, flags)
        self._check(rc, "fdb_inflate_batch")
        return out_len, consumed, status

    def _deflate_packed(self, fn, in_base, in_off, in_len, out_base, out_off, out_cap):
        n = len(in_off)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        in_len = np.ascontiguousarray(in_len, dtype=np.uint64)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        out_cap = np.ascontiguousarray(out_cap, dtype=np.uint64)
        out_len = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        rc = fn(self._h, _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out_base), _ptr(out_off), _ptr(out_cap),
                _ptr(out_len), _ptr(status), n)
        self._check(rc, "fdb_deflate_*_batch")
        return out_len, status

    def deflate_ultrafast_packed(self, in_base, in_off, in_len, out_base, out_off, out_cap):
        return self._deflate_packed(self.lib.L.fdb_deflate_ultrafast_batch, in_base, in_off, in_len, out_base, out_off,
                                    out_cap)

    def deflate_stored_packed(self, in_base, in_off, in_len, out_base, out_off, out_cap):
        return self._deflate_packed(self.lib.L.fdb_deflate_stored_batch, in_base, in_off, in_len, out_base, out_off,
                                    out_cap)

    # ---- convenience batch calls on lists of bytes ----------------------------------------------
    @staticmethod
    def _pack(items: Sequence[bytes], align: int = 16):
        lens = np.array([len(b) for b in items], dtype=np.uint64)
        offs = np.zeros(len(items), dtype=np.uint64)
        pos = 0
        for i, l in enumerate(lens):
            offs[i] = pos
            pos = (pos + int(l) + align - 1) // align * align if align > 1 else pos + int(l)
        base = np.zeros(max(pos, 1), dtype=np.uint8)
        for b, o in zip(items, offs):
            if len(b):
                base[int(o): int(o) + len(b)] = np.frombuffer(b, dtype=np.uint8)
        return base, offs, lens

    def inflate_batch(self, streams: Sequence[bytes], out_caps: Sequence[int], flags: int = 0, align: int = 16):
        """-> (status int32[n], outputs list[bytes], consumed uint64[n]).  outputs[i] holds out_len[i] bytes."""
        if len(streams) == 0:
            return np.zeros(0, np.int32), [], np.zeros(0, np.uint64)
        in_base, in_off, in_len = self._pack(streams, align)
        caps = np.asarray(out_caps, dtype=np.uint64)
        out_off = np.zeros(len(caps), dtype=np.uint64)
        pos = 0
        for i, c in enumerate(caps):
            out_off[i] = pos
            pos = (pos + int(c) + align - 1) // align * align if align > 1 else pos + int(c)
        out_base = np.zeros(max(pos, 1), dtype=np.uint8)
        out_len, consumed, status = self.inflate_packed(in_base, in_off, in_len, out_base, out_off, caps, flags)
        outs = [out_base[int(o): int(o) + int(l)].tobytes() for o, l in zip(out_off, out_len)]
        return status, outs, consumed

    def _deflate_batch(self, packed_fn, bound_fn, inputs: Sequence[bytes], align: int):
        if len(inputs) == 0:
            return []
        in_base, in_off, in_len = self._pack(inputs, align)
        caps = np.array([bound_fn(len(b)) for b in inputs], dtype=np.uint64)
        out_off = np.zeros(len(caps), dtype=np.uint64)
        pos = 0
        for i, c in enumerate(caps):
            out_off[i] = pos
            pos = (pos + int(c) + align - 1) // align * align if align > 1 else pos + int(c)
        out_base = np.zeros(max(pos, 1), dtype=np.uint8)
        out_len, status = packed_fn(in_base, in_off, in_len, out_base, out_off, caps)
        if (status != 0).any():
            raise FdbError(f"deflate status {status[status != 0][:4]}")
        return [out_base[int(o): int(o) + int(l)].tobytes() for o, l in zip(out_off, out_len)]

    def deflate_ultrafast_batch(self, inputs: Sequence[bytes], align: int = 16) -> list[bytes]:
        return self._deflate_batch(self.deflate_ultrafast_packed, self.lib.L.fdb_deflate_ultrafast_bound, inputs, align)

    def deflate_stored_batch(self, inputs: Sequence[bytes], align: int = 16) -> list[bytes]:
        return self._deflate_batch(self.deflate_stored_packed, self.lib.L.fdb_deflate_stored_bound, inputs, align)

    # ---- streaming decoders (fdb_stream_*): many in-flight Decompressor::read state machines, state on the device ----
    def stream_open(self, n: int = 1) -> np.ndarray:
        ids = np.zeros(n, dtype=np.uint32)
        self._check(self.lib.L.fdb_stream_open_batch(self._h, _ptr(ids), n), "fdb_stream_open_batch")
        return ids

    def stream_close(self, ids) -> None:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        if self._h:
            self._check(self.lib.L.fdb_stream_close_batch(self._h, _ptr(ids), len(ids)), "fdb_stream_close_batch")

    def stream_read_packed(self, ids, in_base, in_off, in_len, out_base, out_off, out_room, flags: int = 0):
        """one launch advances every decoder in `ids`; -> (produced uint64[n], status int32[n])"""
        n = len(ids)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        in_off, in_len, out_off, out_room = (np.ascontiguousarray(a, dtype=np.uint64) for a in (in_off, in_len, out_off, out_room))
        produced = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        rc = self.lib.L.fdb_stream_read_batch(self._h, _ptr(ids), _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out_base),
                                              _ptr(out_off), _ptr(out_room), _ptr(produced), _ptr(status), n, flags)
        self._check(rc, "fdb_stream_read_batch")
        return produced, status

    def stream_read(self, ids, datas: Sequence[bytes], rooms: Sequence[int], flags: int = 0):
        """-> (status int32[n], outputs list[bytes]) for one step of the decoders `ids`"""
        in_base, in_off, in_len = self._pack(datas, 1)
        rooms = np.asarray(rooms, dtype=np.uint64)
        out_off = np.zeros(len(rooms), dtype=np.uint64)
        out_off[1:] = np.cumsum(rooms[:-1])
        out_base = np.zeros(max(int(rooms.sum()), 1), dtype=np.uint8)
        produced, status = self.stream_read_packed(ids, in_base, in_off, in_len, out_base, out_off, rooms, flags)
        return status, [out_base[int(o): int(o) + int(p)].tobytes() for o, p in zip(out_off, produced)]

    # ---- device-pointer calls (ints / torch data_ptr()); enqueue only -----------------------------
    def inflate_device(self, d_in, d_in_off, d_in_len, d_out, d_out_off, d_out_cap, d_out_len, d_consumed, d_status,
                       n: int, flags: int = 0, stream: int = 0):
        rc = self.lib.L.fdb_inflate_batch_device(self._h, d_in, d_in_off, d_in_len, d_out, d_out_off, d_out_cap,
                                                 d_out_len, d_consumed, d_status, n, flags, stream)
        self._check(rc, "fdb_inflate_batch_device")

    def deflate_ultrafast_device(self, d_in, d_in_off, d_in_len, d_out, d_out_off, d_out_cap, d_out_len, d_status,
                                 n: int, stream: int = 0):
        rc = self.lib.L.fdb_deflate_ultrafast_batch_device(self._h, d_in, d_in_off, d_in_len, d_out, d_out_off,
                                                           d_out_cap, d_out_len, d_status, n, stream)
        self._check(rc, "fdb_deflate_ultrafast_batch_device")

    def deflate_stored_device(self, d_in, d_in_off, d_in_len, d_out, d_out_off, d_out_cap, d_out_len, d_status,
                              n: int, stream: int = 0):
        rc = self.lib.L.fdb_deflate_stored_batch_device(self._h, d_in, d_in_off, d_in_len, d_out, d_out_off,
                                                        d_out_cap, d_out_len, d_status, n, stream)
        self._check(rc, "fdb_deflate_stored_batch_device")

    # ---- PNG row filters (png_filter.cuh) ------------------------------------------------------
    def png_unfilter_device(self, d_filtered, d_filtered_off, d_raw, d_raw_off, d_height, d_stride, d_bpp, d_status,
                            n: int, stream: int = 0):
        rc = self.lib.L.fdb_png_unfilter_batch_device(self._h, d_filtered, d_filtered_off, d_raw, d_raw_off, d_height,
                                                      d_stride, d_bpp, d_status, n, stream)
        self._check(rc, "fdb_png_unfilter_batch_device")

    def png_filter_device(self, d_raw, d_raw_off, d_filtered, d_filtered_off, d_height, d_stride, d_bpp, mode: int,
                          d_status, n: int, stream: int = 0):
        rc = self.lib.L.fdb_png_filter_batch_device(self._h, d_raw, d_raw_off, d_filtered, d_filtered_off, d_height,
                                                    d_stride, d_bpp, mode, d_status, n, stream)
        self._check(rc, "fdb_png_filter_batch_device")

    def png_encode_device(self, d_raw, d_raw_off, d_height, d_stride, d_bpp, mode: int, d_out, d_out_off, d_out_cap, d_out_len,
                          d_filter_status, d_status, n: int, stream: int = 0):
        """filter (mode 0..4) + ultra-fast deflate of raw images in one kernel, device pointers (fdb_png_encode_batch_device)"""
        rc = self.lib.L.fdb_png_encode_batch_device(self._h, d_raw, d_raw_off, d_height, d_stride, d_bpp, mode, d_out, d_out_off,
                                                    d_out_cap, d_out_len, d_filter_status, d_status, n, stream)
        self._check(rc, "fdb_png_encode_batch_device")

    def _png_batch(self, unfilter: bool, images: Sequence[bytes], geometry: Sequence[tuple], mode: int = 0):
        """images[i] with geometry[i] = (height, stride, bpp) -> (status[n], list of bytes)"""
        n = len(images)
        h = np.array([g[0] for g in geometry], dtype=np.uint32)
        s = np.array([g[1] for g in geometry], dtype=np.uint32)
        b = np.array([g[2] for g in geometry], dtype=np.uint32)
        filt = h.astype(np.uint64) * (1 + s.astype(np.uint64))
        raw = h.astype(np.uint64) * s.astype(np.uint64)
        in_sz, out_sz = (filt, raw) if unfilter else (raw, filt)
        for i, im in enumerate(images):
            if len(im) != int(in_sz[i]):
                raise ValueError(f"image {i}: {len(im)} bytes, geometry says {int(in_sz[i])}")
        in_base, in_off, _ = self._pack(images)
        out_off = np.zeros(n, dtype=np.uint64)
        if n > 1:
            out_off[1:] = np.cumsum((out_sz[:-1] + np.uint64(15)) & ~np.uint64(15))
        out_base = np.zeros(int(out_off[-1] + out_sz[-1]) + 16 if n else 16, dtype=np.uint8)
        status = np.zeros(n, dtype=np.int32)
        if unfilter:
            rc = self.lib.L.fdb_png_unfilter_batch(self._h, _ptr(in_base), _ptr(in_off), _ptr(out_base), _ptr(out_off),
                                                   _ptr(h), _ptr(s), _ptr(b), _ptr(status), n)
        else:
            rc = self.lib.L.fdb_png_filter_batch(self._h, _ptr(in_base), _ptr(in_off), _ptr(out_base), _ptr(out_off),
                                                 _ptr(h), _ptr(s), _ptr(b), mode, _ptr(status), n)
        self._check(rc, "fdb_png_*_batch")
        outs = [out_base[int(out_off[i]): int(out_off[i]) + int(out_sz[i])].tobytes() for i in range(n)]
        return status, outs

    def png_unfilter_batch(self, filtered: Sequence[bytes], geometry: Sequence[tuple]):
        return self._png_batch(True, filtered, geometry)

    def png_filter_batch(self, raw: Sequence[bytes], geometry: Sequence[tuple], mode: int = 4):
        return self._png_batch(False, raw, geometry, mode)

    def crc32_batch(self, items: Sequence[bytes], seed: int = 0) -> np.ndarray:
        """CRC-32 (zlib's crc32) of every item, continued from `seed`, computed on the device -> uint32[n]"""
        n = len(items)
        if n == 0:
            return np.zeros(0, dtype=np.uint32)
        base, off, lens = self._pack(items)
        crc = np.zeros(n, dtype=np.uint32)
        rc = self.lib.L.fdb_crc32_batch(self._h, _ptr(base), _ptr(off), _ptr(lens), seed & 0xffffffff, _ptr(crc), n)
        self._check(rc, "fdb_crc32_batch")
        return crc

    def synth_tiles_device(self, d_out: int, first_tile: int, n_tiles: int, width: int, height: int, seed: int,
                           stream: int = 0):
        rc = self.lib.L.fdb_synth_tiles_device(self._h, d_out, first_tile, n_tiles, width, height, seed, stream)
        self._check(rc, "fdb_synth_tiles_device")

    def ultrafast_bound(self, n: int) -> int:
        return int(self.lib.L.fdb_deflate_ultrafast_bound(n))

    def stored_bound(self, n: int) -> int:
        return int(self.lib.L.fdb_deflate_stored_bound(n))


def synth_tiles_host(first_tile: int, n_tiles: int, width: int, height: int, seed: int,
                     lib: NativeLib | None = None) -> np.ndarray:
    """Synthetic PNG-filtered RGBA tiles, host version of the device generator."""
    lib = lib if lib is not None else _native.default_lib()
    tb = int(lib.L.fdb_synth_tile_bytes(width, height))
    out = np.zeros(n_tiles * tb, dtype=np.uint8)
    rc = lib.L.fdb_synth_tiles_host(_ptr(out), first_tile, n_tiles, width, height, seed)
    if rc != 0:
        raise FdbError("fdb_synth_tiles_host failed")
    return out.reshape(n_tiles, tb)


# ---- module-level mirror of the reference API -----------------------------------------------------
_default_ctx: dict[int, Context] = {}


class MultiContext(Context):
    """One fdb_multi: the GPUs of one box behind one handle (include/fdeflate_b200.h, fdb_multi_*).  Batches are
    partitioned by stream (byte-balanced, no collective) and every device runs the ordinary host-buffer call on
    its share from its own host thread; arguments and results are those of Context.  Only the host-buffer batch
    calls exist on a device set."""

    def __init__(self, devices: Sequence[int], lib: NativeLib | None = None):
        self.lib = lib if lib is not None else _native.default_lib()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self.lib.L.fdb_multi_create(devs, len(devices), C.byref(h))
        if rc != 0 or not h:
            raise FdbError(f"fdb_multi_create(devices={list(devices)}) failed with code {rc}: no usable CUDA device "
                           f"(fdeflate_b200 has no CPU fallback)")
        self._m = h
        self._h = None
        self.devices = list(devices)

    def close(self):
        if getattr(self, "_m", None):
            self.lib.L.fdb_multi_destroy(self._m)
            self._m = None

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise FdbError(f"{what} failed ({rc}): {self.lib.L.fdb_multi_last_error(self._m).decode()}")

    def inflate_packed(self, in_base, in_off, in_len, out_base, out_off, out_cap, flags: int = 0):
        n = len(in_off)
        in_off, in_len, out_off, out_cap = (np.ascontiguousarray(a, dtype=np.uint64) for a in (in_off, in_len, out_off, out_cap))
        out_len = np.zeros(n, dtype=np.uint64)
        consumed = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        rc = self.lib.L.fdb_multi_inflate_batch(self._m, _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out_base), _ptr(out_off),
                                                _ptr(out_cap), _ptr(out_len), _ptr(consumed), _ptr(status), n, flags)
        self._check(rc, "fdb_multi_inflate_batch")
        return out_len, consumed, status

    def _multi_deflate(self, fn, what, in_base, in_off, in_len, out_base, out_off, out_cap):
        n = len(in_off)
        in_off, in_len, out_off, out_cap = (np.ascontiguousarray(a, dtype=np.uint64) for a in (in_off, in_len, out_off, out_cap))
        out_len = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        rc = fn(self._m, _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out_base), _ptr(out_off), _ptr(out_cap), _ptr(out_len),
                _ptr(status), n)
        self._check(rc, what)
        return out_len, status

    def deflate_ultrafast_packed(self, in_base, in_off, in_len, out_base, out_off, out_cap):
        return self._multi_deflate(self.lib.L.fdb_multi_deflate_ultrafast_batch, "fdb_multi_deflate_ultrafast_batch", in_base, in_off,
                                   in_len, out_base, out_off, out_cap)

    def deflate_stored_packed(self, in_base, in_off, in_len, out_base, out_off, out_cap):
        return self._multi_deflate(self.lib.L.fdb_multi_deflate_stored_batch, "fdb_multi_deflate_stored_batch", in_base, in_off,
                                   in_len, out_base, out_off, out_cap)

    def last_partition(self, n: int) -> np.ndarray:
        owner = np.zeros(n, dtype=np.uint32)
        if self.lib.L.fdb_multi_last_partition(self._m, _ptr(owner), n) != 0:
            raise FdbError("fdb_multi_last_partition: no call of that size yet")
        return owner


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def decompress_to_vec_bounded(data: bytes, maxlen: int, ctx: Context | None = None, flags: int = 0) -> bytes:
    """reference src/decompress.rs:1111-1144.  Raises BoundedDecompressionError."""
    ctx = ctx or default_context()
    cap = min(maxlen, max(1024, 4 * len(data)))
    while True:
        status, outs, _ = ctx.inflate_batch([bytes(data)], [cap], flags)
        st = int(status[0])
        if st == ST_OK:
            return outs[0]
        if st == ST_OUTPUT_TOO_LARGE:
            if cap >= maxlen:
                raise BoundedDecompressionError(partial_output=outs[0])
            cap = min(maxlen, cap * 4)  # the reference grows its Vec and carries on (:1132-1134)
            continue
        raise BoundedDecompressionError(inner=DecompressionError(st))


def decompress_to_vec(data: bytes, ctx: Context | None = None) -> bytes:
    """reference src/decompress.rs:1079-1087.  Raises DecompressionError."""
    try:
        return decompress_to_vec_bounded(data, 1 << 62, ctx)
    except BoundedDecompressionError as e:
        if e.inner is not None:
            raise e.inner from None
        raise


def compress_to_vec_ultra_fast(data: bytes, ctx: Context | None = None) -> bytes:
    """reference src/compress/mod.rs:313-317."""
    ctx = ctx or default_context()
    return ctx.deflate_ultrafast_batch([bytes(data)])[0]


def compress_to_vec_stored(data: bytes, ctx: Context | None = None) -> bytes:
    """compress_to_vec_with_level(data, 0), reference src/compress/mod.rs:299-303 with level 0."""
    ctx = ctx or default_context()
    return ctx.deflate_stored_batch([bytes(data)])[0]


class Decompressor:
    """The reference's streaming decoder (src/decompress.rs:96-342) with the read() contract of :158-184.  The state
    machine lives on the device (fdb_stream_*, csrc/inflate_general.cuh: K3StreamState): every call resumes at the
    token boundary the last one stopped at, so the work of a call is proportional to the bytes of that call whatever
    the chunking (byte-wise feeding is linear in the stream length).  read() takes all of `data` (what cannot be parsed
    yet is kept by the context) and returns (len(data), bytes written); when the output is full, call again with more
    room (data may be empty).  Many decoders advance in one launch through Context.stream_read."""

    def __init__(self, ctx: Context | None = None):
        self._ctx = ctx or default_context()
        self._id = self._ctx.stream_open(1)
        self._done = False
        self._flags = 0

    def close(self):
        if self._id is not None:
            try:
                self._ctx.stream_close(self._id)
            finally:
                self._id = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ignore_adler32(self):
        self._flags |= FLAG_IGNORE_ADLER32

    def is_done(self) -> bool:
        return self._done

    def read(self, data: bytes, output: np.ndarray, output_position: int):
        if self._done:
            return 0, 0  # :185-187
        if output_position > output.size:
            raise IndexError("output_position out of bounds")  # the reference panics (:189)
        data = bytes(data)
        in_base = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, dtype=np.uint8)
        room = output.size - output_position
        z = np.zeros(1, dtype=np.uint64)
        produced, status = self._ctx.stream_read_packed(self._id, in_base, z, np.array([len(data)], dtype=np.uint64), output,
                                                        np.array([output_position], dtype=np.uint64),
                                                        np.array([room], dtype=np.uint64), self._flags)
        st = int(status[0])
        if st > 0:
            raise DecompressionError(st)
        if st == ST_OK:
            self._done = True
        return len(data), int(produced[0])


class UltraFastCompressor:
    """reference src/compress/ultrafast.rs:9-181.  `writer` needs a .write(bytes) method.

    The reference's output depends on write_data call boundaries (the zero-run state and the 8-byte
    chunking are local to each call).  Every call is therefore compressed as its own
    stream in ONE device batch at finish(), and the token bits are spliced on the host, which
    reproduces the reference byte for byte for any call pattern."""

    def __init__(self, writer, ctx: Context | None = None):
        self._w = writer
        self._ctx = ctx or default_context()
        self._calls: list[bytes] = []

    def write_data(self, data: bytes):
        self._calls.append(bytes(data))

    def finish(self):
        calls = self._calls if self._calls else [b""]
        streams = self._ctx.deflate_ultrafast_batch(calls)
        if len(streams) == 1:
            self._w.write(streams[0])
            return self._w
        self._w.write(_splice_ultrafast(streams, calls))
        return self._w


_HEADER_BITS = 53 * 8 + 5


def _splice_ultrafast(streams: Sequence[bytes], calls: Sequence[bytes]) -> bytes:
    import zlib

    acc = int.from_bytes(streams[0][:54], "little") & ((1 << _HEADER_BITS) - 1)
    nbits = _HEADER_BITS
    for s in streams:
        body = int.from_bytes(s[:-4], "little")
        end = body.bit_length() - 12  # the EOB code (12 bits) ends at the highest set bit
        tok = (body >> _HEADER_BITS) & ((1 << (end - _HEADER_BITS)) - 1)
        acc |= tok << nbits
        nbits += end - _HEADER_BITS
    acc |= 2303 << nbits
    nbits += 12
    out = acc.to_bytes((nbits + 7) // 8, "little")
    adler = 1
    for c in calls:
        adler = zlib.adler32(c, adler)
    return out + adler.to_bytes(4, "big")


class Compressor:
    """reference src/compress/mod.rs:47-215 restricted to level 0 ("stored").  Levels 1-9 are the
    reference's sequential LZ77 encoders and are out of scope of the accelerated path."""

    def __init__(self, writer, level: int, zlib: bool, ctx: Context | None = None):
        if level != 0:
            raise NotImplementedError("only level 0 (stored) is on the accelerated path")
        self._w = writer
        self._zlib = zlib
        self._ctx = ctx or default_context()
        self._data = bytearray()

    def write_data(self, data: bytes):
        self._data += bytes(data)  # stored blocks split at 65535-byte boundaries of the concatenation

    def finish(self):
        s = self._ctx.deflate_stored_batch([bytes(self._data)])[0]
        self._w.write(s if self._zlib else s[2:-4])
        return self._w
