"""Batch PNG image data on the GPU: container framing on the host, everything per byte on the device.

The hot path of image-rs/fdeflate sits inside a PNG codec (reference README.md:11): decode = collect the IDAT
chunks -> inflate -> undo the row filters; encode = row filters -> (ultra-fast) deflate -> IDAT.  This module is
that caller for whole batches: `decode_batch` / `encode_batch` parse and write the
chunk structure on the host (signature, IHDR, IDAT, IEND, CRC-32 per chunk) and hand the zlib streams and the
pixels to `fdb_png_decode_batch` / `fdb_png_encode_batch` (include/fdeflate_b200.h), which keep the intermediate
filtered image on the device.  Non-interlaced images only; palette images decode to their index plane.
"""
from __future__ import annotations

import binascii
import struct
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from .api import Context, FdbError, STATUS_NAMES, _ptr, default_context

SIGNATURE = b"\x89PNG\r\n\x1a\n"
_CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}
_DEPTHS = {0: (1, 2, 4, 8, 16), 2: (8, 16), 3: (1, 2, 4, 8), 4: (8, 16), 6: (8, 16)}


class PngError(ValueError):
    pass


@dataclass(frozen=True)
class PngInfo:
    width: int
    height: int
    bit_depth: int
    color_type: int

    @property
    def channels(self) -> int:
        return _CHANNELS[self.color_type]

    @property
    def bpp(self) -> int:  # bytes per complete pixel, rounded up to 1 (PNG 9.2)
        return max(1, self.channels * self.bit_depth // 8)

    @property
    def stride(self) -> int:  # bytes per raw row
        return (self.width * self.channels * self.bit_depth + 7) // 8


def parse(png: bytes, crc_todo: list | None = None) -> tuple[PngInfo, bytes]:
    """-> (IHDR fields, the zlib stream = all IDAT payloads concatenated).  Checks the signature, the chunk CRCs,
    the chunk order IHDR .. IDAT .. IEND and the IHDR values.  With `crc_todo` the chunk CRCs are not computed
    here: (type + body, stored CRC, chunk type) is appended for every chunk, to be checked in one device batch."""
    if png[:8] != SIGNATURE:
        raise PngError("not a PNG file")
    pos, info, idat, seen_end = 8, None, [], False
    while pos < len(png):
        if pos + 12 > len(png):
            raise PngError("truncated chunk")
        (length,), ctype = struct.unpack(">I", png[pos:pos + 4]), png[pos + 4:pos + 8]
        body = png[pos + 8:pos + 8 + length]
        if len(body) != length or pos + 12 + length > len(png):
            raise PngError("truncated chunk")
        (crc,) = struct.unpack(">I", png[pos + 8 + length:pos + 12 + length])
        if crc_todo is not None:
            crc_todo.append((png[pos + 4:pos + 8 + length], crc, ctype))
        elif binascii.crc32(ctype + body) & 0xffffffff != crc:
            raise PngError(f"bad CRC in {ctype!r} chunk")
        pos += 12 + length
        if info is None:
            if ctype != b"IHDR" or length != 13:
                raise PngError("first chunk is not IHDR")
            w, h, depth, color, comp, filt, interlace = struct.unpack(">IIBBBBB", body)
            if color not in _CHANNELS or depth not in _DEPTHS[color] or w == 0 or h == 0 or comp or filt:
                raise PngError("invalid IHDR")
            if interlace:
                raise PngError("interlaced images are not supported")
            info = PngInfo(w, h, depth, color)
        elif ctype == b"IDAT":
            idat.append(body)
        elif ctype == b"IEND":
            seen_end = True
            break
    if info is None or not idat or not seen_end:
        raise PngError("missing IHDR, IDAT or IEND")
    return info, b"".join(idat)


def _chunk(ctype: bytes, body: bytes) -> bytes:
    return struct.pack(">I", len(body)) + ctype + body + struct.pack(">I", binascii.crc32(ctype + body) & 0xffffffff)


def _to_array(info: PngInfo, raw: np.ndarray) -> np.ndarray:
    if info.bit_depth == 8:
        a = raw.reshape(info.height, info.width, info.channels)
        return a[..., 0] if info.channels == 1 else a
    if info.bit_depth == 16:
        a = raw.view(">u2").astype(np.uint16).reshape(info.height, info.width, info.channels)
        return a[..., 0] if info.channels == 1 else a
    return raw.reshape(info.height, info.stride)  # packed sub-byte samples, as stored


def decode_batch(pngs: Sequence[bytes], ctx: Context | None = None, crc: str = "device") -> list[np.ndarray]:
    """PNG files -> pixel arrays (h, w[, channels]) uint8 / uint16; sub-byte depths come back as packed rows.
    crc = "device": all chunk CRCs of the batch in one fdb_crc32_batch call; "host": binascii while parsing;
    "skip": not checked."""
    ctx = ctx or default_context()
    if crc not in ("device", "host", "skip"):
        raise ValueError("crc must be 'device', 'host' or 'skip'")
    todo: list | None = None if crc == "host" else []
    infos, streams = zip(*(parse(p, todo) for p in pngs)) if pngs else ((), ())
    n = len(infos)
    if n == 0:
        return []
    if crc == "device" and todo:
        got = ctx.crc32_batch([t[0] for t in todo])
        for (_, want, ctype), g in zip(todo, got):
            if int(g) != want:
                raise PngError(f"bad CRC in {ctype!r} chunk")
    idat_base, idat_off, idat_len = ctx._pack(streams)
    h = np.array([i.height for i in infos], dtype=np.uint32)
    s = np.array([i.stride for i in infos], dtype=np.uint32)
    b = np.array([i.bpp for i in infos], dtype=np.uint32)
    raw_sz = h.astype(np.uint64) * s.astype(np.uint64)
    raw_off = np.zeros(n, dtype=np.uint64)
    raw_off[1:] = np.cumsum((raw_sz[:-1] + np.uint64(15)) & ~np.uint64(15))
    raw = np.zeros(int(raw_off[-1] + raw_sz[-1]) + 16, dtype=np.uint8)
    status = np.zeros(n, dtype=np.int32)
    rc = ctx.lib.L.fdb_png_decode_batch(ctx._h, _ptr(idat_base), _ptr(idat_off), _ptr(idat_len), _ptr(raw), _ptr(raw_off),
                                        _ptr(h), _ptr(s), _ptr(b), _ptr(status), n)
    ctx._check(rc, "fdb_png_decode_batch")
    out = []
    for i, info in enumerate(infos):
        if status[i] != 0:
            raise PngError(f"image {i}: {STATUS_NAMES[status[i]] if status[i] < len(STATUS_NAMES) else status[i]}")
        out.append(_to_array(info, raw[int(raw_off[i]): int(raw_off[i]) + int(raw_sz[i])].copy()))
    return out


def encode_batch(images: Sequence[np.ndarray], ctx: Context | None = None, filter_mode: int = 5) -> list[bytes]:
    """uint8 / uint16 arrays (h, w) or (h, w, 2|3|4) -> PNG files: row filters (mode 0..4 fixed, 5 adaptive) and
    ultra-fast deflate on the GPU, one IDAT chunk per image."""
    ctx = ctx or default_context()
    n = len(images)
    if n == 0:
        return []
    infos, raws = [], []
    for a in images:
        a = np.asarray(a)
        if a.dtype not in (np.uint8, np.uint16) or a.ndim not in (2, 3) or a.size == 0:
            raise PngError("expected a non-empty uint8 / uint16 array of shape (h, w) or (h, w, channels)")
        ch = 1 if a.ndim == 2 else a.shape[2]
        color = {1: 0, 2: 4, 3: 2, 4: 6}.get(ch)
        if color is None:
            raise PngError("1, 2, 3 or 4 channels")
        infos.append(PngInfo(a.shape[1], a.shape[0], 8 * a.dtype.itemsize, color))
        raws.append(np.ascontiguousarray(a.astype(">u2") if a.dtype == np.uint16 else a).tobytes())
    raw_base, raw_off, _ = ctx._pack(raws)
    h = np.array([i.height for i in infos], dtype=np.uint32)
    s = np.array([i.stride for i in infos], dtype=np.uint32)
    b = np.array([i.bpp for i in infos], dtype=np.uint32)
    caps = np.array([ctx.ultrafast_bound(int(hh) * (1 + int(ss))) for hh, ss in zip(h, s)], dtype=np.uint64)
    out_off = np.zeros(n, dtype=np.uint64)
    out_off[1:] = np.cumsum(caps[:-1])
    out = np.zeros(int(out_off[-1] + caps[-1]), dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    rc = ctx.lib.L.fdb_png_encode_batch(ctx._h, _ptr(raw_base), _ptr(raw_off), _ptr(h), _ptr(s), _ptr(b), filter_mode,
                                        _ptr(out), _ptr(out_off), _ptr(caps), _ptr(out_len), _ptr(status), n)
    ctx._check(rc, "fdb_png_encode_batch")
    for i in range(n):
        if status[i] != 0:
            raise FdbError(f"image {i}: status {int(status[i])}")
    zs = [out[int(out_off[i]): int(out_off[i]) + int(out_len[i])].tobytes() for i in range(n)]
    # the IDAT chunk CRCs (over "IDAT" + payload) in one device batch; IHDR / IEND are a few bytes each
    crcs = ctx.crc32_batch(zs, seed=binascii.crc32(b"IDAT"))
    files = []
    for info, z, c in zip(infos, zs, crcs):
        ihdr = struct.pack(">IIBBBBB", info.width, info.height, info.bit_depth, info.color_type, 0, 0, 0)
        idat = struct.pack(">I", len(z)) + b"IDAT" + z + struct.pack(">I", int(c))
        files.append(SIGNATURE + _chunk(b"IHDR", ihdr) + idat + _chunk(b"IEND", b""))
    return files


MAX_IMAGE_BYTES = 1 << 32  # decode_files_batch: largest raw image (height * stride from the untrusted IHDR) it allocates for


def decode_files_batch(pngs: Sequence[bytes], ctx: Context | None = None, max_image_bytes: int = MAX_IMAGE_BYTES) -> list[np.ndarray]:
    """Same result as decode_batch, but the container is handled by the library too (`fdb_png_probe_batch` +
    `fdb_png_decode_files_batch`: chunk walk in C++, chunk CRCs / IDAT gathering / inflate / unfilter on the device),
    so nothing per file or per byte happens in Python."""
    ctx = ctx or default_context()
    n = len(pngs)
    if n == 0:
        return []
    base, off, lens = ctx._pack(pngs, align=1)
    w, h, depth, color, stride = (np.zeros(n, dtype=np.uint32) for _ in range(5))
    status = np.zeros(n, dtype=np.int32)
    rc = ctx.lib.L.fdb_png_probe_batch(_ptr(base), _ptr(off), _ptr(lens), _ptr(w), _ptr(h), _ptr(depth), _ptr(color),
                                       _ptr(stride), _ptr(status), n)
    if rc != 0:
        raise FdbError("fdb_png_probe_batch failed")
    for i in range(n):
        if status[i] != 0:
            raise PngError(f"image {i}: {STATUS_NAMES[status[i]]}")
    raw_sz = h.astype(np.uint64) * stride.astype(np.uint64)
    for i in range(n):
        if int(raw_sz[i]) > max_image_bytes:
            raise PngError(f"image {i}: {int(raw_sz[i])} bytes of pixels exceed max_image_bytes = {max_image_bytes}")
    raw_off = np.zeros(n, dtype=np.uint64)
    raw_off[1:] = np.cumsum((raw_sz[:-1] + np.uint64(15)) & ~np.uint64(15))
    raw = np.zeros(int(raw_off[-1] + raw_sz[-1]) + 16, dtype=np.uint8)
    rc = ctx.lib.L.fdb_png_decode_files_batch(ctx._h, _ptr(base), _ptr(off), _ptr(lens), _ptr(raw), _ptr(raw_off),
                                              _ptr(raw_sz), _ptr(status), n)
    ctx._check(rc, "fdb_png_decode_files_batch")
    out = []
    for i in range(n):
        if status[i] != 0:
            raise PngError(f"image {i}: {STATUS_NAMES[status[i]] if status[i] < len(STATUS_NAMES) else status[i]}")
        info = PngInfo(int(w[i]), int(h[i]), int(depth[i]), int(color[i]))
        out.append(_to_array(info, raw[int(raw_off[i]): int(raw_off[i]) + int(raw_sz[i])].copy()))
    return out


def encode_files_batch(images: Sequence[np.ndarray], ctx: Context | None = None, filter_mode: int = 5) -> list[bytes]:
    """Same files as encode_batch, written by the library (`fdb_png_encode_files_batch`: row filter, ultra-fast deflate
    and the IDAT CRC on the device, the chunk framing in C++), nothing per image in Python but slicing the result."""
    ctx = ctx or default_context()
    n = len(images)
    if n == 0:
        return []
    raws, w, h, depth, color = [], [], [], [], []
    for a in images:
        a = np.asarray(a)
        if a.dtype not in (np.uint8, np.uint16) or a.ndim not in (2, 3) or a.size == 0:
            raise PngError("expected a non-empty uint8 / uint16 array of shape (h, w) or (h, w, channels)")
        ch = 1 if a.ndim == 2 else a.shape[2]
        ct = {1: 0, 2: 4, 3: 2, 4: 6}.get(ch)
        if ct is None:
            raise PngError("1, 2, 3 or 4 channels")
        raws.append(np.ascontiguousarray(a.astype(">u2") if a.dtype == np.uint16 else a).tobytes())
        w.append(a.shape[1]); h.append(a.shape[0]); depth.append(8 * a.dtype.itemsize); color.append(ct)
    raw_base, raw_off, _ = ctx._pack(raws)
    w, h, depth, color = (np.array(x, dtype=np.uint32) for x in (w, h, depth, color))
    caps = np.array([ctx.lib.L.fdb_png_file_bound(int(a), int(b), int(c), int(d)) for a, b, c, d in zip(w, h, depth, color)],
                    dtype=np.uint64)
    caps = (caps + np.uint64(15)) & ~np.uint64(15)
    f_off = np.zeros(n, dtype=np.uint64)
    f_off[1:] = np.cumsum(caps[:-1])
    files = np.zeros(int(f_off[-1] + caps[-1]), dtype=np.uint8)
    f_len = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    rc = ctx.lib.L.fdb_png_encode_files_batch(ctx._h, _ptr(raw_base), _ptr(raw_off), _ptr(w), _ptr(h), _ptr(depth), _ptr(color),
                                              filter_mode, _ptr(files), _ptr(f_off), _ptr(caps), _ptr(f_len), _ptr(status), n)
    ctx._check(rc, "fdb_png_encode_files_batch")
    out = []
    for i in range(n):
        if status[i] != 0:
            raise FdbError(f"image {i}: status {int(status[i])}")
        out.append(files[int(f_off[i]): int(f_off[i]) + int(f_len[i])].tobytes())
    return out
