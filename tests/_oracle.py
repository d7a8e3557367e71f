"""ctypes wrapper over oracle/libnaforacle.so -- the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (nafcodec_b200) never does.
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libnaforacle.so")

SEC_NAMES = ["id", "comment", "length", "mask", "sequence", "quality"]
DNA, RNA, PROTEIN, TEXT = 0, 1, 2, 3


class _Sec(C.Structure):
    _fields_ = [("present", C.c_int32), ("_pad", C.c_int32), ("original_size", C.c_uint64),
                ("compressed_size", C.c_uint64), ("offset", C.c_uint64)]


class Layout(C.Structure):
    _fields_ = [("format_version", C.c_int32), ("sequence_type", C.c_int32), ("flags", C.c_uint32),
                ("name_separator", C.c_int32), ("line_length", C.c_uint64), ("number_of_sequences", C.c_uint64),
                ("header_size", C.c_uint64), ("sec", _Sec * 6)]


class _Records(C.Structure):
    _fields_ = [("n_records", C.c_uint64),
                ("ids", C.c_void_p), ("id_off", C.c_void_p), ("id_present", C.c_void_p),
                ("comments", C.c_void_p), ("com_off", C.c_void_p), ("com_present", C.c_void_p),
                ("sequence", C.c_void_p), ("seq_off", C.c_void_p), ("seq_present", C.c_void_p),
                ("quality", C.c_void_p), ("qual_off", C.c_void_p), ("qual_present", C.c_void_p),
                ("lengths", C.c_void_p), ("len_present", C.c_void_p)]


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def lib():
    global _lib
    if _lib is None:
        src_mtime = max(os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in ("naf_oracle.c", "naf_oracle.h"))
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < src_mtime:
            build()
        L = C.CDLL(LIB_PATH)
        L.nafo_zstd_version.restype = C.c_char_p
        L.nafo_last_error.restype = C.c_char_p
        L.nafo_time_decode.restype = C.c_double
        L.nafo_time_decode_many.restype = C.c_double
        L.nafo_mask_runs.restype = C.c_int64
        L.nafo_synth_mask.restype = C.c_uint64
        L.nafo_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().nafo_last_error().decode())


def _buf(b):
    return (C.c_uint8 * len(b)).from_buffer_copy(b) if len(b) else (C.c_uint8 * 1)()


def variable_u64(b: bytes):
    out = C.c_uint64()
    k = lib().nafo_variable_u64(_buf(b), C.c_size_t(len(b)), C.byref(out))
    if k < 0:
        raise OracleError(k, "variable_u64")
    return out.value, k


def write_variable_length(n: int) -> bytes:
    out = (C.c_uint8 * 16)()
    k = lib().nafo_write_variable_length(C.c_uint64(n), out)
    return bytes(out[:k])


def parse(data: bytes) -> Layout:
    L = Layout()
    _check(lib().nafo_parse(_buf(data), C.c_size_t(len(data)), C.byref(L)))
    return L


def parse_header(data: bytes) -> Layout:
    L = Layout()
    _check(lib().nafo_parse_header(_buf(data), C.c_size_t(len(data)), C.byref(L)))
    return L


def zstd_decompress(data: bytes) -> bytes:
    dst = C.c_void_p()
    n = C.c_size_t()
    _check(lib().nafo_zstd_decompress(_buf(data), C.c_size_t(len(data)), C.byref(dst), C.byref(n)))
    out = C.string_at(dst, n.value)
    lib().nafo_free(dst)
    return out


def zstd_compress(data: bytes, level=3, checksum=False, window_log=0) -> bytes:
    """One magicless zstd frame (optionally with a content checksum): shapes NAF writers never emit, third parties may."""
    dst = C.c_void_p()
    n = C.c_size_t()
    _check(lib().nafo_zstd_compress(_buf(data), C.c_size_t(len(data)), C.c_int(level), C.c_int(int(checksum)), C.c_int(window_log),
                                    C.byref(dst), C.byref(n)))
    out = C.string_at(dst, n.value)
    lib().nafo_free(dst)
    return out


def section_bytes(data: bytes, name: str) -> Optional[bytes]:
    L = parse(data)
    s = L.sec[SEC_NAMES.index(name)]
    if not s.present:
        return None
    return zstd_decompress(data[s.offset:s.offset + s.compressed_size])


def mask_runs(mask: bytes, total: int) -> List[int]:
    cap = len(mask) + 1
    runs = (C.c_uint64 * cap)()
    k = lib().nafo_mask_runs(_buf(mask), C.c_size_t(len(mask)), C.c_uint64(total), runs, C.c_size_t(cap))
    return list(runs[:k])


@dataclass
class Decoded:
    """Flat record-indexed result (numpy views copied out of the oracle's buffers)."""
    n: int
    ids: Optional[bytes]
    id_off: np.ndarray
    id_present: np.ndarray
    comments: Optional[bytes]
    com_off: np.ndarray
    com_present: np.ndarray
    sequence: Optional[bytes]
    seq_off: np.ndarray
    seq_present: np.ndarray
    quality: Optional[bytes]
    qual_off: np.ndarray
    qual_present: np.ndarray
    lengths: np.ndarray
    len_present: np.ndarray

    def _field(self, blob, off, present, i):
        if not present[i]:
            return None
        return blob[int(off[i]):int(off[i + 1])]

    def id(self, i):
        return self._field(self.ids, self.id_off, self.id_present, i)

    def comment(self, i):
        return self._field(self.comments, self.com_off, self.com_present, i)

    def seq(self, i):
        return self._field(self.sequence, self.seq_off, self.seq_present, i)

    def qual(self, i):
        return self._field(self.quality, self.qual_off, self.qual_present, i)

    def length(self, i):
        return int(self.lengths[i]) if self.len_present[i] else None


def _np(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(n, dtype=dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,)).view(dtype).copy()


def decode(data: bytes, id=True, comment=True, sequence=True, quality=True, mask=True) -> Decoded:
    r = _Records()
    _check(lib().nafo_decode(_buf(data), C.c_size_t(len(data)), int(id), int(comment), int(sequence),
                             int(quality), int(mask), C.byref(r)))
    n = r.n_records

    def blob(ptr, off):
        total = int(off[n]) if n else 0
        return C.string_at(ptr, total) if ptr and total else b""

    id_off = _np(r.id_off, n + 1, np.uint64)
    com_off = _np(r.com_off, n + 1, np.uint64)
    seq_off = _np(r.seq_off, n + 1, np.uint64)
    qual_off = _np(r.qual_off, n + 1, np.uint64)
    d = Decoded(n, blob(r.ids, id_off), id_off, _np(r.id_present, n, np.uint8),
                blob(r.comments, com_off), com_off, _np(r.com_present, n, np.uint8),
                blob(r.sequence, seq_off), seq_off, _np(r.seq_present, n, np.uint8),
                blob(r.quality, qual_off), qual_off, _np(r.qual_present, n, np.uint8),
                _np(r.lengths, n, np.uint64), _np(r.len_present, n, np.uint8))
    lib().nafo_free_records(C.byref(r))
    return d


def _pack(strings):
    if strings is None:
        return None, None, None
    off = np.zeros(len(strings) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in strings], dtype=np.uint64)
    blob = b"".join(strings)
    return blob, _buf(blob), off.ctypes.data_as(C.POINTER(C.c_uint64))


def encode(ids=None, comments=None, sequences=None, qualities=None, mask_runs_=None, sequence_type=DNA,
           level=0, flush_per_record=True, line_length=60, name_separator=" ") -> bytes:
    """Restatement of Encoder::push/write (+ Mask section). Each field: list of bytes, or None."""
    n = max(len(x) for x in (ids, comments, sequences, qualities) if x is not None)
    keep = []
    args = []
    for f in (ids, comments, sequences, qualities):
        blob, cb, off = _pack(f)
        keep.append((blob, cb, off))
        args += [cb, off] if f is not None else [None, None]
    if mask_runs_ is not None:
        mr = np.asarray(mask_runs_, dtype=np.uint64)
        mptr, mn = mr.ctypes.data_as(C.POINTER(C.c_uint64)), len(mr)
    else:
        mr, mptr, mn = None, None, 0
    out = C.c_void_p()
    out_len = C.c_size_t()
    _check(lib().nafo_encode(C.c_int(sequence_type), C.c_int(level), C.c_int(int(flush_per_record)),
                             C.c_uint64(line_length), C.c_int(ord(name_separator)), C.c_uint64(n),
                             *args, mptr, C.c_uint64(mn), C.byref(out), C.byref(out_len)))
    res = C.string_at(out, out_len.value)
    lib().nafo_free(out)
    return res


def set_encoder_workers(n: int):
    """Generator-only: libzstd worker threads for encoders created afterwards (0 = the reference's single-threaded pattern)."""
    lib().nafo_set_encoder_workers(C.c_int(n))


def encode_blobs(n, ids=None, comments=None, sequences=None, qualities=None, mask_runs_=None, sequence_type=DNA, level=0,
                 flush_per_record=True, line_length=60, name_separator=" ") -> bytes:
    """encode() for fields that already are (uint8 blob, uint64 offsets[n + 1]) numpy pairs (millions of records)."""
    args, keep = [], []
    for f in (ids, comments, sequences, qualities):
        if f is None:
            args += [None, None]
        else:
            blob, off = np.ascontiguousarray(f[0], dtype=np.uint8), np.ascontiguousarray(f[1], dtype=np.uint64)
            keep.append((blob, off))
            args += [blob.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.POINTER(C.c_uint64))]
    if mask_runs_ is not None:
        mr = np.asarray(mask_runs_, dtype=np.uint64)
        mptr, mn = mr.ctypes.data_as(C.POINTER(C.c_uint64)), len(mr)
    else:
        mr, mptr, mn = None, None, 0
    out = C.c_void_p()
    out_len = C.c_size_t()
    _check(lib().nafo_encode(C.c_int(sequence_type), C.c_int(level), C.c_int(int(flush_per_record)), C.c_uint64(line_length),
                             C.c_int(ord(name_separator)), C.c_uint64(n), *args, mptr, C.c_uint64(mn), C.byref(out), C.byref(out_len)))
    res = C.string_at(out, out_len.value)
    lib().nafo_free(out)
    return res


def synth_chromosome(seed, n) -> np.ndarray:
    dst = np.empty(n, dtype=np.uint8)
    lib().nafo_synth_chromosome(C.c_uint64(seed), C.c_uint64(n), dst.ctypes.data_as(C.c_void_p))
    return dst


def synth_fastq(seed, n_reads, read_len=150):
    """(ids blob, ids offsets, sequence blob, quality blob) of the cfg4 shape, generated in C."""
    ids = np.empty(24 * n_reads + 32, dtype=np.uint8)
    ids_off = np.empty(n_reads + 1, dtype=np.uint64)
    seq = np.empty(n_reads * read_len, dtype=np.uint8)
    qual = np.empty(n_reads * read_len, dtype=np.uint8)
    lib().nafo_synth_fastq(C.c_uint64(seed), C.c_uint64(n_reads), C.c_uint64(read_len), ids.ctypes.data_as(C.c_void_p),
                           ids_off.ctypes.data_as(C.c_void_p), seq.ctypes.data_as(C.c_void_p), qual.ctypes.data_as(C.c_void_p))
    return ids[:int(ids_off[-1])], ids_off, seq, qual


def synth_dna(seed, n, gc=0.5, families=2, repeat_len=5000, copies=7, iupac_rate=1e-5,
              gap_count=0, gap_len=0, telomere=0) -> bytes:
    dst = np.empty(n, dtype=np.uint8)
    lib().nafo_synth_dna(C.c_uint64(seed), C.c_uint64(n), C.c_double(gc), C.c_int(families), C.c_uint64(repeat_len),
                         C.c_int(copies), C.c_double(iupac_rate), C.c_uint64(gap_count), C.c_uint64(gap_len),
                         C.c_uint64(telomere), dst.ctypes.data_as(C.c_void_p))
    return dst.tobytes()


def synth_mask(seed, total, mean_u=2000.0, mean_m=300.0, leading_zero=True) -> List[int]:
    cap = int(total // 4 + 64)
    runs = np.zeros(cap, dtype=np.uint64)
    k = lib().nafo_synth_mask(C.c_uint64(seed), C.c_uint64(total), C.c_double(mean_u), C.c_double(mean_m),
                              C.c_int(int(leading_zero)), runs.ctypes.data_as(C.c_void_p), C.c_uint64(cap))
    return runs[:k].tolist()


def apply_mask(seq: bytes, runs) -> bytes:
    a = np.frombuffer(seq, dtype=np.uint8).copy()
    r = np.asarray(runs, dtype=np.uint64)
    lib().nafo_apply_mask(a.ctypes.data_as(C.c_void_p), C.c_uint64(len(a)), r.ctypes.data_as(C.c_void_p), C.c_uint64(len(r)))
    return a.tobytes()


def time_decode(data: bytes, quality=True, mask=True, iters=1):
    """Seconds for `iters` full CPU decodes (1 thread) and the ASCII sequence bytes of one decode."""
    nb = C.c_uint64()
    t = lib().nafo_time_decode(_buf(data), C.c_size_t(len(data)), int(quality), int(mask), int(iters), C.byref(nb))
    if t < 0:
        raise OracleError(-1, lib().nafo_last_error().decode())
    return t, nb.value


def time_decode_many(archives, threads: int, quality=True, mask=True):
    """Seconds for one CPU decode of every archive on `threads` native threads (one archive per core at a time), and the
    ASCII sequence bytes decoded."""
    n = len(archives)
    keep = [_buf(a) for a in archives]
    ptrs = (C.c_void_p * n)(*[C.cast(k, C.c_void_p) for k in keep])
    lens = (C.c_size_t * n)(*[len(a) for a in archives])
    nb = C.c_uint64()
    t = lib().nafo_time_decode_many(ptrs, lens, C.c_size_t(n), C.c_int(threads), int(quality), int(mask), C.byref(nb))
    if t < 0:
        raise OracleError(-1, lib().nafo_last_error().decode())
    return t, nb.value


def format_text(data: bytes, fmt: str = "auto", line_length=None, comment=True, mask=True) -> bytes:
    """FASTA / FASTQ text of an archive, restated on the CPU from the oracle's records.  The reference crate carries
    line_length / name_separator (data.rs:198-236) but prints nothing itself; the text is pinned by the fixtures'
    source files (masked.fna, LuxC.faa, phix.fastq; tests/test_oracle_golden.py)."""
    lay = parse_header(data)
    want_q = fmt != "fasta"
    d = decode(data, comment=comment, quality=want_q, mask=mask)
    has_q = bool(lay.flags & 0x01)
    fastq = fmt == "fastq" or (fmt == "auto" and has_q)
    if fastq and not has_q:
        raise ValueError("FASTQ output needs qualities")
    W = int(lay.line_length) if line_length is None else int(line_length)
    sep = bytes([lay.name_separator & 0xFF])
    out = []
    for i in range(d.n):
        name, com, seq = d.id(i) or b"", d.comment(i) or b"", d.seq(i) or b""
        head = (b"@" if fastq else b">") + name + ((sep + com) if com else b"") + b"\n"
        out.append(head)
        if fastq:
            q = d.qual(i) or b""
            out += [seq, b"\n+\n", q, b"\n"]
        elif W == 0:
            out += [seq, b"\n"]
        else:
            for k in range(0, len(seq), W):
                out += [seq[k:k + W], b"\n"]
    return b"".join(out)
