"""The C-ABI library loads and exports every symbol include/fdeflate_b200.h declares.
No compute calls: this runs on machines without a GPU."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "fdeflate_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("fdb_inflate_batch", "fdb_inflate_batch_device", "fdb_deflate_ultrafast_batch",
              "fdb_deflate_ultrafast_batch_device", "fdb_deflate_stored_batch", "fdb_create", "fdb_destroy"):
        assert s in syms


def test_cuda_library_exports_every_declared_symbol():
    from fdeflate_b200 import NativeLib, _native

    assert _native.DEFAULT_LIB.exists(), "CUDA library not built: run __graft_entry__.build()"
    lib = NativeLib()
    for s in declared_symbols():
        assert hasattr(lib.L, s), f"{s} is declared in the header but not exported"
    assert set(_native.EXPORTS) == set(declared_symbols())
    assert "CUDA" in lib.version()


def test_bounds_are_pure_functions():
    from fdeflate_b200 import NativeLib

    lib = NativeLib()
    assert lib.L.fdb_deflate_ultrafast_bound(0) >= 60
    assert lib.L.fdb_deflate_ultrafast_bound(262400) >= 54 + 262400 * 12 // 8 + 6
    assert lib.L.fdb_deflate_stored_bound(65535) >= 2 + 5 + 65535 + 2 + 4
    assert lib.L.fdb_synth_tile_bytes(256, 256) == 262400


def test_no_cpu_fallback_without_library(tmp_path):
    from fdeflate_b200 import NativeLib, NativeLibraryMissing

    with pytest.raises(NativeLibraryMissing):
        NativeLib(tmp_path / "missing.so")


def test_product_never_imports_the_oracle():
    """the oracle and the emulator build are checkers: nothing under fdeflate_b200/ may load or call them"""
    banned = ("oracle_lib", "libfdeflate_oracle", "fdeflate_oracle.h", "fdo_", "import oracle", "libfdb_emul.so")
    for p in (ROOT / "fdeflate_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp"):
            t = p.read_text()
            for b in banned:
                if b == "libfdb_emul.so" and p.name == "simt.h":
                    continue  # named in a comment that says the package never loads it
                assert b not in t, f"{p} references {b}"
