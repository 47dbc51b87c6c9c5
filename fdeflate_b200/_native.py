"""ctypes binding of the C ABI declared in include/fdeflate_b200.h.

The product library is fdeflate_b200/libfdeflate_b200.so (CUDA, sm_100a), built in-tree by
`make -C fdeflate_b200/csrc` or `__graft_entry__.build()`.  There is no CPU fallback: if the
library is missing, or no CUDA device is visible when a context is created, this module raises.
(`NativeLib(path)` exists so the test-suite can bind the test-only SIMT-emulator build of the same
sources; nothing in the package passes a path.)
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
DEFAULT_LIB = PKG_DIR / "libfdeflate_b200.so"

EXPORTS = [
    "fdb_create", "fdb_destroy", "fdb_last_error", "fdb_version",
    "fdb_inflate_batch_device", "fdb_inflate_batch",
    "fdb_deflate_ultrafast_bound", "fdb_deflate_ultrafast_batch_device", "fdb_deflate_ultrafast_batch",
    "fdb_deflate_stored_bound", "fdb_deflate_stored_batch_device", "fdb_deflate_stored_batch",
    "fdb_synth_tile_bytes", "fdb_synth_tiles_host", "fdb_synth_tiles_device", "fdb_launch_count", "fdb_last_general_count",
    "fdb_set_pipeline_chunk", "fdb_last_split_spans", "fdb_set_split_large", "fdb_set_split_threshold", "fdb_set_split_scratch", "fdb_png_unfilter_batch_device", "fdb_png_filter_batch_device", "fdb_png_encode_batch_device", "fdb_png_unfilter_batch", "fdb_png_filter_batch", "fdb_png_decode_batch", "fdb_png_encode_batch", "fdb_crc32_batch_device", "fdb_crc32_batch", "fdb_png_probe_batch", "fdb_png_decode_files_batch", "fdb_png_file_bound", "fdb_png_encode_files_batch",
    "fdb_stream_open_batch", "fdb_stream_read_batch", "fdb_stream_close_batch",
    "fdb_multi_create", "fdb_multi_destroy", "fdb_multi_device_count", "fdb_multi_last_error", "fdb_multi_inflate_batch",
    "fdb_multi_deflate_ultrafast_batch", "fdb_multi_deflate_stored_batch", "fdb_multi_last_partition",
]

FLAG_IGNORE_ADLER32 = 1
FLAG_GENERAL_ONLY = 2
FLAG_SPLIT_LARGE = 4


class NativeLibraryMissing(RuntimeError):
    pass


class NativeLib:
    def __init__(self, path: str | Path | None = None):
        path = Path(path) if path is not None else DEFAULT_LIB
        if not path.exists():
            raise NativeLibraryMissing(
                f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                f"g.build()' or make -C fdeflate_b200/csrc). fdeflate_b200 has no CPU fallback.")
        self.path = path
        L = C.CDLL(str(path))
        self.L = L
        vp, sz, u32, u64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64
        L.fdb_create.restype = C.c_int
        L.fdb_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.fdb_destroy.restype = None
        L.fdb_destroy.argtypes = [vp]
        L.fdb_last_error.restype = C.c_char_p
        L.fdb_last_error.argtypes = [vp]
        L.fdb_version.restype = C.c_char_p
        L.fdb_inflate_batch_device.restype = C.c_int
        L.fdb_inflate_batch_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, u32, vp]
        L.fdb_inflate_batch.restype = C.c_int
        L.fdb_inflate_batch.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, u32]
        for name in ("fdb_deflate_ultrafast_batch_device", "fdb_deflate_stored_batch_device"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
        for name in ("fdb_deflate_ultrafast_batch", "fdb_deflate_stored_batch"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, sz]
        for name in ("fdb_deflate_ultrafast_bound", "fdb_deflate_stored_bound"):
            f = getattr(L, name)
            f.restype = sz
            f.argtypes = [sz]
        L.fdb_synth_tile_bytes.restype = sz
        L.fdb_synth_tile_bytes.argtypes = [u32, u32]
        L.fdb_synth_tiles_host.restype = C.c_int
        L.fdb_synth_tiles_host.argtypes = [vp, u64, u64, u32, u32, u64]
        L.fdb_synth_tiles_device.restype = C.c_int
        L.fdb_synth_tiles_device.argtypes = [vp, vp, u64, u64, u32, u32, u64, vp]
        L.fdb_launch_count.restype = u64
        L.fdb_launch_count.argtypes = [vp]
        L.fdb_set_pipeline_chunk.restype = C.c_int
        L.fdb_set_pipeline_chunk.argtypes = [vp, sz]
        L.fdb_last_general_count.restype = C.c_int64
        L.fdb_last_general_count.argtypes = [vp, vp]
        L.fdb_set_split_large.restype = C.c_int
        L.fdb_set_split_large.argtypes = [vp, C.c_int]
        L.fdb_png_unfilter_batch_device.restype = C.c_int
        L.fdb_png_unfilter_batch_device.argtypes = [vp] * 9 + [sz, vp]
        L.fdb_png_filter_batch_device.restype = C.c_int
        L.fdb_png_filter_batch_device.argtypes = [vp] * 8 + [C.c_uint32, vp, sz, vp]
        L.fdb_png_encode_batch_device.restype = C.c_int
        L.fdb_png_encode_batch_device.argtypes = [vp] * 6 + [C.c_uint32] + [vp] * 6 + [sz, vp]
        L.fdb_png_unfilter_batch.restype = C.c_int
        L.fdb_png_unfilter_batch.argtypes = [vp] * 9 + [sz]
        L.fdb_png_filter_batch.restype = C.c_int
        L.fdb_png_filter_batch.argtypes = [vp] * 8 + [C.c_uint32, vp, sz]
        L.fdb_png_decode_batch.restype = C.c_int
        L.fdb_png_decode_batch.argtypes = [vp] * 10 + [sz]
        L.fdb_png_encode_batch.restype = C.c_int
        L.fdb_png_encode_batch.argtypes = [vp] * 6 + [C.c_uint32] + [vp] * 5 + [sz]
        L.fdb_crc32_batch_device.restype = C.c_int
        L.fdb_crc32_batch_device.argtypes = [vp, vp, vp, vp, C.c_uint32, vp, sz, vp]
        L.fdb_crc32_batch.restype = C.c_int
        L.fdb_crc32_batch.argtypes = [vp, vp, vp, vp, C.c_uint32, vp, sz]
        L.fdb_png_probe_batch.restype = C.c_int
        L.fdb_png_probe_batch.argtypes = [vp] * 9 + [sz]
        L.fdb_png_decode_files_batch.restype = C.c_int
        L.fdb_png_decode_files_batch.argtypes = [vp] * 8 + [sz]
        L.fdb_png_file_bound.restype = sz
        L.fdb_png_file_bound.argtypes = [C.c_uint32] * 4
        L.fdb_png_encode_files_batch.restype = C.c_int
        L.fdb_png_encode_files_batch.argtypes = [vp] * 7 + [C.c_uint32] + [vp] * 5 + [sz]
        L.fdb_set_split_threshold.restype = C.c_int
        L.fdb_set_split_threshold.argtypes = [vp, sz, sz]
        L.fdb_set_split_scratch.restype = C.c_int
        L.fdb_set_split_scratch.argtypes = [vp, sz]
        L.fdb_last_split_spans.restype = C.c_int64
        L.fdb_last_split_spans.argtypes = [vp, vp]
        L.fdb_stream_open_batch.restype = C.c_int
        L.fdb_stream_open_batch.argtypes = [vp, vp, sz]
        L.fdb_stream_close_batch.restype = C.c_int
        L.fdb_stream_close_batch.argtypes = [vp, vp, sz]
        L.fdb_stream_read_batch.restype = C.c_int
        L.fdb_stream_read_batch.argtypes = [vp] * 10 + [sz, u32]
        L.fdb_multi_create.restype = C.c_int
        L.fdb_multi_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.fdb_multi_destroy.restype = None
        L.fdb_multi_destroy.argtypes = [vp]
        L.fdb_multi_device_count.restype = C.c_int
        L.fdb_multi_device_count.argtypes = [vp]
        L.fdb_multi_last_error.restype = C.c_char_p
        L.fdb_multi_last_error.argtypes = [vp]
        L.fdb_multi_inflate_batch.restype = C.c_int
        L.fdb_multi_inflate_batch.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, u32]
        for name in ("fdb_multi_deflate_ultrafast_batch", "fdb_multi_deflate_stored_batch"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, sz]
        L.fdb_multi_last_partition.restype = C.c_int
        L.fdb_multi_last_partition.argtypes = [vp, vp, sz]

    def version(self) -> str:
        return self.L.fdb_version().decode()


_default: NativeLib | None = None


def default_lib() -> NativeLib:
    global _default
    if _default is None:
        _default = NativeLib()
    return _default
