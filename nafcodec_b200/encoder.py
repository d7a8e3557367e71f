"""Host-side mirror of the reference ENCODER surface (nafcodec/src/encoder/mod.rs:63-384; Python binding
nafcodec-py/nafcodec/lib.rs:466-597, lib.pyi:69-87) over the encode-side entry point of the C ABI.

What runs on the device (nafgpu_pack, csrc/naf_pack.cu): the 4-bit IUPAC packing with the odd-length carry
(encoder/writer.rs:31-90), the length words (encoder/mod.rs:37-44) and -- with `mask=True`, an extension: the reference's
mask writer is commented out (encoder/mod.rs:240) and its SequenceWriter rejects lower case -- the extraction of the
soft-mask runs.  zstd compression stays on the CPU as in the reference (`_zstd.StreamEncoder`, the reference's call pattern:
one flush per record for the comment, sequence and quality streams), so for upper-case input the archive is byte-identical
to what the reference encoder's restatement in oracle/ writes (tests/test_encoder.py).
"""
import ctypes as C
import os
from typing import List, Optional

import numpy as np

from . import _ffi
from ._zstd import StreamEncoder
from .data import Flag, Record, SequenceType
from .errors import raise_for_status

_SEQTYPES = {"dna": SequenceType.Dna, "rna": SequenceType.Rna, "protein": SequenceType.Protein, "text": SequenceType.Text}


def write_variable_length(n: int) -> bytes:          # encoder/mod.rs:22-35
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append(0x80 | (n & 0x7F))
        n >>= 7
    return bytes(reversed(out))


def pack_sequences(sequences: List[bytes], sequence_type: int = 0, extract_mask: bool = False, device: int = 0, _library=None):
    """nafgpu_pack on the records' residues: returns (packed bytes, length-stream bytes, mask-stream bytes or None, mask runs)."""
    from .decoder import shared_context
    lib = _library or _ffi.default_library()
    ctx = shared_context(device, lib)
    lengths = np.array([len(s) for s in sequences], dtype=np.uint64)
    blob = b"".join(sequences)
    src = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob or b"\0")
    inp = _ffi.PackInput(C.cast(src, C.c_void_p), lengths.ctypes.data_as(C.c_void_p), len(sequences), len(blob), int(sequence_type), int(extract_mask))
    res = _ffi.PackResult()
    with ctx._lock:
        rc = lib.dll.nafgpu_pack(ctx._ctx, C.byref(inp), C.byref(res))
        if rc == _ffi.ERR_INVALID_DATA and res.first_invalid != _ffi.NO_RECORD:
            raise ValueError(f"invalid sequence: unexpected character at residue {res.first_invalid}")      # Error::InvalidSequence -> ValueError (lib.rs:56-58)
        raise_for_status(lib, rc, ctx._ctx)
        packed = C.string_at(res.packed, res.packed_size) if res.packed_size else b""
        words = C.string_at(res.length_words, res.length_size) if res.length_size else b""
        mask = (C.string_at(res.mask, res.mask_size) if res.mask_size else b"") if extract_mask else None
        return packed, words, mask, int(res.n_mask_runs)


class Encoder:
    """An encoder for Nucleotide Archive Format files (surface of nafcodec.Encoder; records are buffered until `close`,
    like the reference buffers its compressed streams in temporary storage until `Encoder::write`)."""

    def __init__(self, file, sequence_type: str = "dna", *, id: bool = False, comment: bool = False, sequence: bool = False,
                 quality: bool = False, compression_level: int = 0, mask: bool = False, line_length: int = 60,
                 name_separator: str = " ", device: int = 0, _library=None):
        if sequence_type not in _SEQTYPES:
            raise ValueError(f"invalid sequence type: {sequence_type!r}")
        self._type = _SEQTYPES[sequence_type]
        self._file = file
        self._own = not hasattr(file, "write")
        self._fh = open(os.fspath(file), "wb") if self._own else file
        self._id, self._comment, self._sequence, self._quality = id, comment, sequence, quality
        self._mask = mask and sequence and self._type.is_nucleotide()
        self._level, self._line_length, self._sep = compression_level, line_length, name_separator
        self._device, self._library = device, _library
        self._records: List[Record] = []
        self._closed = False

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        if exc_type is None:
            self.close()
        return False

    def write(self, record: Record) -> None:
        """Encoder::push (encoder/mod.rs:250-327): the same field checks, raised as ValueError like the binding (lib.rs:56-66)."""
        if self._id and record.id is None:
            raise ValueError("missing record field: id")
        if self._comment and record.comment is None:
            raise ValueError("missing record field: comment")
        if self._sequence and record.sequence is None:
            raise ValueError("missing record field: sequence")
        if self._quality and record.quality is None:
            raise ValueError("missing record field: quality")
        length = record.length
        if self._sequence:
            if length is not None and length != len(record.sequence):
                raise ValueError("invalid sequence length")
            length = len(record.sequence)
        if self._quality:
            if length is not None and length != len(record.quality):
                raise ValueError("invalid sequence length")
        if self._sequence and self._type.is_nucleotide():
            # SequenceWriter::encode rejects the record that carries the bad character (encoder/mod.rs:283-286): checked on the
            # host per record so that the error is raised by the `write` that caused it; the device reports the same
            # position again at `close` (nafgpu_pack_result.first_invalid), which the tests compare
            # (bytes.translate with a delete table: one C loop over the record; anything left over is not in the alphabet)
            if record.sequence.encode("latin-1", "replace").translate(None, _valid_bytes(self._type == SequenceType.Rna, self._mask)):
                raise ValueError("invalid sequence: unexpected sequence character")
        self._records.append(record)

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        self._fh.write(self._build())
        if self._own:
            self._fh.close()

    # -- Encoder::write (encoder/mod.rs:334-384) ---------------------------------------------------------------------------
    def _build(self) -> bytes:
        recs = self._records
        nucl = self._type.is_nucleotide()
        e_len = StreamEncoder(self._level)                                  # always created (encoder/mod.rs:188)
        e_id = StreamEncoder(self._level) if self._id else None
        e_com = StreamEncoder(self._level) if self._comment else None
        e_seq = StreamEncoder(self._level) if self._sequence else None
        e_qual = StreamEncoder(self._level) if self._quality else None
        packed = words = mask = None
        if self._sequence and nucl:
            packed, words, mask, _ = pack_sequences([r.sequence.encode("ascii") for r in recs], int(self._type), self._mask,
                                                    self._device, self._library)
        elif self._sequence or self._quality:
            lens = [len(r.sequence if self._sequence else r.quality) for r in recs]
            words = b"".join((b"\xff\xff\xff\xff" * (l // 0xFFFFFFFF)) + (l % 0xFFFFFFFF).to_bytes(4, "little") for l in lens)
        pos = wpos = 0
        for r in recs:
            if self._sequence or self._quality:
                l = len(r.sequence) if self._sequence else len(r.quality)
                nw = 4 * (l // 0xFFFFFFFF + 1)
                e_len.write(words[wpos:wpos + nw])
                wpos += nw
            if e_id is not None:
                e_id.write(r.id.encode("utf-8") + b"\0")
            if e_com is not None:
                e_com.write(r.comment.encode("utf-8") + b"\0")
                e_com.flush()                                               # encoder/mod.rs:271
            if e_seq is not None:
                l = len(r.sequence)
                if nucl:
                    if l:                                                   # the bytes this record completes: [pos / 2, (pos + l) / 2)
                        e_seq.write(packed[pos // 2:(pos + l) // 2])
                        e_seq.flush()                                       # writer.rs:88
                    e_seq.written += l - ((pos + l) // 2 - pos // 2)        # WriteCounter counts residues (counter.rs:25-34)
                else:
                    e_seq.write(r.sequence.encode("utf-8"))
                pos += l
                e_seq.flush()                                               # encoder/mod.rs:298
            if e_qual is not None:
                e_qual.write(r.quality.encode("utf-8"))
                e_qual.flush()                                              # encoder/mod.rs:319
        flags = (Flag.Id if self._id else 0) | (Flag.Comment if self._comment else 0)
        if self._sequence:
            flags |= Flag.Sequence | Flag.Length
        if self._quality:
            flags |= Flag.Quality | Flag.Length
        if self._mask:
            flags |= Flag.Mask
        out = bytearray(b"\x01\xf9\xec")
        out += bytes([1]) if self._type == SequenceType.Dna else bytes([2, int(self._type)])       # v1 iff DNA (encoder/mod.rs:167-171)
        out += bytes([int(flags), ord(self._sep)]) + write_variable_length(self._line_length) + write_variable_length(len(recs))

        def section(orig, enc):                                            # write_block! (encoder/mod.rs:357-374)
            comp = enc.finish()
            return write_variable_length(orig) + write_variable_length(len(comp)) + comp

        if e_id is not None:
            out += section(e_id.written, e_id)
        if e_com is not None:
            out += section(e_com.written, e_com)
        out += section(e_len.written, e_len)                               # ALWAYS written (encoder/mod.rs:378)
        if self._mask:
            e_mask = StreamEncoder(self._level)
            e_mask.write(mask)
            out += section(len(mask), e_mask)
        if e_seq is not None:
            if nucl and pos % 2:                                           # into_inner pads the cached nibble (writer.rs:21-28)
                e_seq.write(packed[pos // 2:pos // 2 + 1])
                e_seq.written -= 1
            e_seq.flush()
            out += section(e_seq.written, e_seq)
        if e_qual is not None:
            out += section(e_qual.written, e_qual)
        return bytes(out)


_VALID_DNA = set(b"ACGTRYSWKMBDHVN-")
_VALID_RNA = set(b"ACGURYSWKMBDHVN-")
_VALID_BYTES = {}


def _valid_bytes(rna: bool, mask: bool) -> bytes:
    """The alphabet SequenceWriter::encode accepts (encoder/writer.rs:31-90), lower case too when the mask is recorded."""
    key = (rna, mask)
    if key not in _VALID_BYTES:
        ok = _VALID_RNA if rna else _VALID_DNA
        if mask:
            ok = ok | {c + 32 for c in ok if 65 <= c <= 90}
        _VALID_BYTES[key] = bytes(sorted(ok))
    return _VALID_BYTES[key]
