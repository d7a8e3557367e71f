"""FASTA / FASTQ text formatted on the device against the source texts shipped beside the reference's
fixtures and against the oracle's CPU formatter, byte for byte."""
import numpy as np
import pytest

import _cases as K
import _oracle as O
import nafcodec_b200 as N
from nafcodec_b200 import _ffi
from _harness import BACKENDS, library
from conftest import read_golden

SOURCES = [("masked.naf", "masked.fna"), ("LuxC.naf", "LuxC.faa"), ("phix.naf", "phix.fastq")]


def source_text(name):
    """The fixture's source text; data/masked.fna was saved without the newline after its last line, which a formatter
    (upstream unnaf included) always writes."""
    t = read_golden(name)
    return t if t.endswith(b"\n") else t + b"\n"


@pytest.mark.parametrize("naf,text", SOURCES)
def test_oracle_formatter_reproduces_the_fixture_sources(naf, text):
    """Pins the CPU formatter: the texts the fixtures were made from come back byte for byte."""
    assert O.format_text(read_golden(naf)) == source_text(text)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("naf,text", SOURCES)
def test_fixture_sources_byte_for_byte(backend, naf, text):
    got = N.to_text(read_golden(naf), _library=library(backend))
    assert got == source_text(text)


@pytest.mark.parametrize("backend", BACKENDS)
def test_text_options_against_the_oracle(backend):
    lib = library(backend)
    for name in ["masked.naf", "phix.naf", "CP040672.naf", "LuxC.naf", "NZ_AAEN01000029.naf"]:
        data = read_golden(name)
        for ll in (None, 0, 1, 7, 16, 17, 60, 100000):
            assert N.to_fasta(data, line_length=ll, _library=lib) == O.format_text(data, "fasta", ll), (name, ll)
        assert N.to_fasta(data, mask=False, comment=False, _library=lib) == O.format_text(data, "fasta", comment=False, mask=False), name
    assert N.to_fastq(read_golden("phix.naf"), _library=lib) == O.format_text(read_golden("phix.naf"), "fastq")
    with pytest.raises(ValueError):
        N.to_fastq(read_golden("masked.naf"), _library=lib)           # no quality section


@pytest.mark.parametrize("backend", BACKENDS)
def test_generated_archives_and_batches(backend):
    lib = library(backend)
    big = 400_000 if backend == "emul" else 6_000_000
    arcs = [
        K.genome(7, big, level=3),                                          # one long record: every chunk is mid-record
        K.multi_record_dna(3, 300, 900, level=3, empty_every=7),            # ragged records, odd lengths, empty records
        K.fastq_reads(5, 2000 if backend == "emul" else 50000),             # many short records per chunk
        O.encode(ids=[b"", b"x", b""], comments=[b"", b"", b"c"], sequences=[b"", b"A", b""], level=3, line_length=0),
        O.encode(ids=[b"only"], sequences=[b"ACGT" * 40], level=3, line_length=80, name_separator="|"),
        K.text_archive([b"MKV", b"", b"protein text with spaces " * 50], sequence_type=O.PROTEIN),
    ]
    for i, a in enumerate(arcs):
        for fmt in ("auto", "fasta"):
            assert N.to_text(a, fmt, _library=lib) == O.format_text(a, fmt), (i, fmt)
    got = N.to_text(arcs, "auto", _library=lib)                             # one job for all of them
    assert got == [O.format_text(a, "auto") for a in arcs]


@pytest.mark.parametrize("backend", BACKENDS)
def test_text_raises_at_invalid_utf8(backend):
    lib = library(backend)
    arc = O.encode(sequence_type=O.TEXT, ids=[b"a", b"b"], sequences=[b"ok", b"\xff\xfe"], level=3)
    with pytest.raises(N.NafUnicodeError):
        N.to_fasta(arc, _library=lib)


@pytest.mark.parametrize("backend", BACKENDS)
def test_text_of_odd_archives(backend):
    """Archives without ids / with empty records / whose header claims more records than the streams hold: absent fields are
    empty in the text, exactly as in the CPU formatter."""
    lib = library(backend)
    arcs = [
        O.encode(sequences=[b"ACGT" * 10, b"GG"], level=3),                                     # no ids: '>' alone
        O.encode(ids=[b"x"], sequences=[b""], level=3, line_length=5),                          # an empty record: header line only
        O.encode(ids=[b"a", b"b", b"c"], sequences=[b"", b"", b""], mask_runs_=[0, 0, 5]),
        O.encode(ids=[b"r"], sequences=[b"ACGU" * 30], sequence_type=O.RNA, level=3, line_length=7),
        O.encode(sequences=[b"ACGT", b"TT"], qualities=[b"IIII", b"##"], level=3),              # FASTQ without ids
    ]
    claimed = bytearray(O.encode(ids=[b"r1", b"r2"], comments=[b"c1", b"c2"], sequences=[b"ACGT", b"GG"]))
    claimed[O.parse(bytes(claimed)).header_size - 1] = 5                                        # 5 records claimed, 2 stored
    arcs.append(bytes(claimed))
    for i, a in enumerate(arcs):
        assert N.to_text(a, _library=lib) == O.format_text(a), i
    with pytest.raises(ValueError):
        N.to_fasta(O.encode(ids=[b"only", b"ids"]), _library=lib)                               # no sequence to print
