"""Shared parity harness: run an archive through a nafgpu library (real CUDA build or the CPU SIMT emulator build of
the same kernel sources) and compare every record field with the CPU oracle."""
import io
import os
import subprocess

import numpy as np
import pytest

import _oracle as O
import nafcodec_b200 as N
from nafcodec_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_DIR = os.path.join(ROOT, "tests", "emul")
EMUL_LIB = os.path.join(EMUL_DIR, "_build", "libnafgpu_emul.so")

_libs = {}


def emul_library():
    if "emul" not in _libs:
        override = os.environ.get("NAFGPU_EMUL_LIB")       # e.g. an AddressSanitizer build of the emulator library (tests/emul/README)
        if override:
            _libs["emul"] = _ffi.Library(override)
        else:
            subprocess.run(["make", "-s", "-j8", "-C", EMUL_DIR], check=True)
            _libs["emul"] = _ffi.Library(EMUL_LIB)
    return _libs["emul"]


def cuda_library():
    if "cuda" not in _libs:
        _libs["cuda"] = _ffi.default_library()      # raises if the product .so is missing: no fallback
    return _libs["cuda"]


def library(kind):
    return emul_library() if kind == "emul" else cuda_library()


BACKENDS = [pytest.param("emul", id="emul"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


def decode_soa(lib, data: bytes, **fields) -> N.ArchiveResult:
    a = N.parse_archive(data, lib)
    want = N.decoder._want_bits(fields.get("id", True), fields.get("comment", True), fields.get("sequence", True),
                                fields.get("quality", True), fields.get("mask", True))
    return N.shared_context(0, lib).decode([a], want)[0]


def assert_same_as_oracle(res: N.ArchiveResult, d: O.Decoded, label=""):
    assert res.n_records == d.n, label
    n = d.n
    # None-ness per record (mod.rs:356-399)
    idp = np.arange(n) < (res.n_ids if res.ids is not None else 0)
    cop = np.arange(n) < (res.n_comments if res.comments is not None else 0)
    lep = np.arange(n) < (res.n_lengths if res.lengths is not None else 0)
    assert np.array_equal(idp, d.id_present.astype(bool)), label + " id presence"
    assert np.array_equal(cop, d.com_present.astype(bool)), label + " comment presence"
    assert np.array_equal(lep, d.len_present.astype(bool)), label + " length presence"
    assert np.array_equal(lep & (res.sequence is not None), d.seq_present.astype(bool)), label + " sequence presence"
    assert np.array_equal(lep & (res.quality is not None), d.qual_present.astype(bool)), label + " quality presence"
    nl = int(lep.sum())
    if nl:
        assert np.array_equal(res.lengths[:nl], d.lengths[:nl]), label + " lengths"
        assert np.array_equal(res.record_offsets[:nl + 1], np.concatenate([[0], np.cumsum(d.lengths[:nl])]).astype(np.uint64)), label
    if res.sequence is not None and nl:
        assert res.sequence == d.sequence, label + " sequence bytes" + _first_diff(res.sequence, d.sequence)
    if res.quality is not None and nl:
        assert res.quality == d.quality, label + " quality bytes" + _first_diff(res.quality, d.quality)
    ni = int(idp.sum())
    for i in range(ni):
        assert res.id_bytes(i) == d.id(i), f"{label} id {i}"
    nc = int(cop.sum())
    for i in range(nc):
        assert res.comment_bytes(i) == d.comment(i), f"{label} comment {i}"


def _first_diff(a, b):
    if a == b:
        return ""
    if len(a) != len(b):
        return f" (len {len(a)} vs {len(b)})"
    x = np.frombuffer(a, np.uint8)
    y = np.frombuffer(b, np.uint8)
    k = int(np.nonzero(x != y)[0][0])
    return f" (first diff at {k}: got {a[max(0, k - 8):k + 8]!r} want {b[max(0, k - 8):k + 8]!r}; {int((x != y).sum())} bytes differ)"


def check_parity(kind, data: bytes, label="", **fields):
    lib = library(kind)
    res = decode_soa(lib, data, **fields)
    d = O.decode(data, **fields)
    assert_same_as_oracle(res, d, label)
    return res, d


def records(kind, data: bytes, **fields):
    return list(N.Decoder(io.BytesIO(data), _library=library(kind), **fields))
