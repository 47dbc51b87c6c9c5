"""fdeflate_b200 -- B200-native batch zlib codec, drop-in for the hot path of image-rs/fdeflate.

The compute lives in fdeflate_b200/libfdeflate_b200.so (hand-written CUDA for sm_100a behind the C ABI
of include/fdeflate_b200.h).  Importing the package never touches a CPU implementation: creating a
Context without the CUDA library or without a GPU raises.
"""
from ._native import FLAG_GENERAL_ONLY, FLAG_IGNORE_ADLER32, FLAG_SPLIT_LARGE, NativeLib, NativeLibraryMissing
from .api import (
    STATUS_NAMES,
    BoundedDecompressionError,
    Compressor,
    Context,
    DecompressionError,
    Decompressor,
    FdbError,
    MultiContext,
    UltraFastCompressor,
    compress_to_vec_stored,
    compress_to_vec_ultra_fast,
    decompress_to_vec,
    decompress_to_vec_bounded,
    default_context,
    synth_tiles_host,
)

__all__ = [
    "FLAG_GENERAL_ONLY", "FLAG_IGNORE_ADLER32", "FLAG_SPLIT_LARGE", "NativeLib", "NativeLibraryMissing", "STATUS_NAMES",
    "BoundedDecompressionError", "Compressor", "Context", "DecompressionError", "Decompressor", "FdbError", "MultiContext",
    "UltraFastCompressor", "compress_to_vec_stored", "compress_to_vec_ultra_fast", "decompress_to_vec",
    "decompress_to_vec_bounded", "default_context", "synth_tiles_host",
]
