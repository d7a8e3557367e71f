"""Edge cases of the container and of the record iterator, bit-exact against the oracle (both backends), plus
hypothesis-driven random archives (odd lengths, empty records, zero-length mask units, every field subset)."""
import io

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import _cases as K
import _oracle as O
import nafcodec_b200 as N
from _harness import BACKENDS, check_parity, library, records

pytestmark = pytest.mark.parametrize("backend", BACKENDS)


def test_zero_records_and_empty_sequences(backend):
    check_parity(backend, O.encode(ids=[], sequences=[]) if False else O.encode(ids=[b""], sequences=[b""]), "one empty record")
    check_parity(backend, O.encode(ids=[b"a", b"b", b"c"], sequences=[b"", b"", b""], mask_runs_=[0, 0, 5]), "all empty")
    check_parity(backend, O.encode(ids=[b"a", b"b"], sequences=[b"", b"ACGTN"], qualities=[b"", b"IIIII"]), "first empty")
    check_parity(backend, O.encode(sequences=[b"A"]), "single residue, no ids")
    check_parity(backend, O.encode(ids=[b"only", b"ids", b""]), "ids only: unflagged Length section follows (encoder/mod.rs:378)")


def test_header_says_more_records_than_the_streams_hold(backend):
    # fields run out: id/comment/length become None for the later records (CStringReader / LengthReader return None at EOF)
    data = bytearray(O.encode(ids=[b"r1", b"r2"], comments=[b"c1", b"c2"], sequences=[b"ACGT", b"GG"]))
    L = O.parse(bytes(data))
    assert data[L.header_size - 1] == 2           # number_of_sequences varint
    data[L.header_size - 1] = 5
    res, d = check_parity(backend, bytes(data), "5 records claimed, 2 stored")
    assert [res.length(i) for i in range(5)] == [4, 2, None, None, None]
    recs = records(backend, bytes(data))
    assert len(recs) == 5 and recs[2].id is None and recs[2].sequence is None and recs[1].sequence == "GG"


def test_title_is_skipped_and_trailing_bytes_ignored(backend):
    # decoder/mod.rs:191-196 (title parsed and discarded); bytes after the last flagged section are never read
    base = O.encode(ids=[b"x1", b"x2"], sequences=[b"ACGTACGTT", b"NNNNACGT"], mask_runs_=[3, 4, 100])
    L = O.parse(base)
    hdr = bytearray(base[:L.header_size])
    hdr[4] |= 0x40                                # v1: magic(3) version(1) flags(1)
    title = b"a title with \xc3\xa9"
    data = bytes(hdr) + O.write_variable_length(len(title)) + title + base[L.header_size:] + b"trailing garbage \x00\xff"
    res, d = check_parity(backend, data, "title + trailing bytes")
    assert d.id(0) == b"x1"
    dec = N.Decoder(io.BytesIO(data), _library=library(backend))
    assert dec.number_of_sequences == 2 and dec.read().sequence == "ACGtacgTT"


def test_header_properties_and_iterator_protocol(backend):
    # nafcodec-py test_decoder.py:39-47 (len counts down), lib.pyi properties, context manager, read() -> None at the end
    data = O.encode(ids=[b"a", b"b", b"c"], sequences=[b"AC", b"G", b"T"], line_length=70, name_separator="|")
    with N.open(io.BytesIO(data), "r", _library=library(backend)) as dec:
        assert (dec.sequence_type, dec.format_version, dec.line_length, dec.name_separator, dec.number_of_sequences) == ("dna", "v1", 70, "|", 3)
        assert len(dec) == 3
        assert next(dec).id == "a"
        assert len(dec) == 2
        assert [r.sequence for r in dec] == ["G", "T"]
        assert len(dec) == 0 and dec.read() is None
    b = N.DecoderBuilder.from_flags(N.Flag.Id | N.Flag.Quality)            # decoder/mod.rs:93-101 doctest
    b._library = library(backend)
    from conftest import read_golden
    r = b.with_bytes(read_golden("phix.naf")).read()
    assert r.sequence is None and r.quality is not None and r.id is not None and r.comment is None


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(data=st.data())
def test_random_archives(backend, data):
    rng = np.random.default_rng(data.draw(st.integers(0, 2 ** 31)))
    n = data.draw(st.integers(1, 12))
    lens = [data.draw(st.sampled_from([0, 1, 2, 3, 7, 31, 32, 33, 64, 255, 256, 257, 1000, 4097])) for _ in range(n)]
    alphabet = data.draw(st.sampled_from([b"ACGT", b"ACGTN", b"ACGTRYKMSWBDHVN-", b"A", b"AC"]))
    seqs = [K.random_dna(rng, l, alphabet) for l in lens]
    use = dict(ids=data.draw(st.booleans()), comments=data.draw(st.booleans()), qualities=data.draw(st.booleans()))
    total = sum(lens)
    runs = None
    if data.draw(st.booleans()) and total:
        runs, s = [], 0
        while s < total:
            r = data.draw(st.sampled_from([0, 1, 2, 5, 31, 32, 33, 254, 255, 256, 300, 511, 5000]))
            runs.append(r)
            s += r
    kw = dict(sequences=seqs, mask_runs_=runs, level=data.draw(st.sampled_from([1, 3, 19])), flush_per_record=data.draw(st.booleans()))
    if use["ids"]:
        kw["ids"] = [b"id%d" % i for i in range(n)]
    if use["comments"]:
        kw["comments"] = [bytes(rng.integers(32, 127, size=int(rng.integers(0, 40))).astype(np.uint8)) for _ in range(n)]
    if use["qualities"]:
        kw["qualities"] = [bytes(rng.integers(33, 74, size=l).astype(np.uint8)) for l in lens]
    arc = O.encode(**kw)
    fields = {f: data.draw(st.booleans()) for f in ("id", "comment", "sequence", "quality", "mask")}
    check_parity(backend, arc, f"lens={lens} runs={runs} fields={fields}", **fields)
