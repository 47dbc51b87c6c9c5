"""Compiles the C++ mirror of the reference API (include/fdeflate_b200.hpp) with g++ and runs its
harness: against the emulator build on CPU, against the CUDA library with -m gpu."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _build_and_run(lib: Path, tmp_path):
    exe = tmp_path / "test_cpp_api"
    subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), "-I", str(ROOT / "oracle"),
           str(ROOT / "tests" / "cpp" / "test_cpp_api.cpp"), "-o", str(exe),
           str(lib), str(ROOT / "oracle" / "libfdeflate_oracle.so"),
           f"-Wl,-rpath,{lib.parent}", f"-Wl,-rpath,{ROOT / 'oracle'}", "-pthread"]
    subprocess.run(cmd, check=True, capture_output=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp api ok" in r.stdout


@pytest.mark.emul
def test_cpp_mirror_on_emulator(emul_lib, tmp_path):
    _build_and_run(emul_lib.path, tmp_path)


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(tmp_path):
    _build_and_run(ROOT / "fdeflate_b200" / "libfdeflate_b200.so", tmp_path)
