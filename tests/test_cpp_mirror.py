"""The C++ host mirror (include/nafgpu.hpp) runs the reference's own decoder tests (tests/cpp/test_decoder.cpp restates
nafcodec/tests/decoder/{dna,fastq,protein}.rs): compiled with g++ and linked against the emulator build (CPU tier) or
the sm_100a library (GPU tier)."""
import os
import subprocess

import pytest

from _harness import BACKENDS, library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("backend", BACKENDS)
def test_reference_decoder_tests_in_cpp(backend, tmp_path):
    lib = library(backend).path
    exe = str(tmp_path / "test_decoder")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_decoder.cpp"), "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference decoder tests passed" in r.stdout
