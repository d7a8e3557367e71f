"""CPU tier: the C-ABI library loads, exports every symbol include/nafgpu.h declares, the host-only entry points work,
and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import nafcodec_b200 as N
from nafcodec_b200 import _ffi
from conftest import read_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nafgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nafgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_ffi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _ffi.default_library()            # built by __graft_entry__.build(); raises ImportError if missing
    for name in declared_symbols():
        assert hasattr(lib.dll, name), name


def test_host_parser_known_answers():
    lib = _ffi.default_library()
    a = N.parse_archive(read_golden("NZ_AAEN01000029.naf"), lib)
    h = a.header
    assert (h.format_version, h.sequence_type, h.flags, chr(h.name_separator), h.line_length, h.number_of_sequences) == (1, 0, 0x3E, " ", 80, 30)
    secs = [(s.present, s.original_size, s.compressed_size) for s in a.sections]
    assert secs == [(1, 540, 122), (1, 2308, 212), (1, 120, 120), (1, 21525, 15), (1, 5488676, 1330710), (0, 0, 0)]
    # parser.rs:141-152 header known answer (no sections follow: flags say there are -> UnexpectedEof, like nom Incomplete)
    v = C.c_uint64()
    for n, enc in [(0, "00"), (127, "7f"), (128, "8100"), (129, "8101"), (34359738367, "ffffffff7f"), (34359738368, "818080808000")]:
        b = bytes.fromhex(enc)
        assert lib.dll.nafgpu_variable_u64(b, len(b), C.byref(v)) == len(b) and v.value == n
    with pytest.raises(N.NafParseError):
        N.parse_archive(b"\x01\xF9\xED\x01\x00 \x00\x00", lib)
    with pytest.raises(N.NafIoError):
        N.parse_archive(b"\x01\xF9\xEC\x01", lib)
    assert lib.strerror(-8).startswith("no CUDA device")


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(N.NafDeviceError):
        N.Decoder(os.path.join(ROOT, "tests", "golden", "phix.naf")).read()


def test_empty_input_is_unexpected_eof():
    # decoder/mod.rs:470-476
    import io
    with pytest.raises(N.NafIoError) as e:
        N.Decoder(io.BytesIO(b""))
    assert e.value.status == _ffi.ERR_UNEXPECTED_EOF
    with pytest.raises(FileNotFoundError):          # nafcodec-py test_decoder.py:123-125
        N.Decoder("")
    with pytest.raises(IsADirectoryError):          # test_decoder.py:127-130
        N.Decoder(os.path.dirname(__file__))
