"""The reference's Python-binding tests, restated against nafcodec_b200 (the same names, arguments and assertions):
  nafcodec-py/nafcodec/tests/test_decoder.py:19-130   _TestDecoder x {TestDecoderHandle (BytesIO), TestDecoderFile (open file)}
  nafcodec-py/nafcodec/tests/test_open.py:23-70       nafcodec.open in the four combinations read/write x file object/filename
and, for DecoderBuilder::buffer_size (nafcodec/src/decoder/mod.rs:104-112; `buffer_size=` in lib.rs), the property the
reference has by construction: the records do not depend on the buffer size.  Here an explicit buffer_size switches the
decoder to windows (nafgpu_job_fetch_window): the archive is decoded into HBM once and the records cross PCIe a window at a
time, which the tests exercise from one record per window up to the whole archive in one."""
import io
import os
import shutil

import numpy as np
import pytest

import _oracle as O
import nafcodec_b200 as nafcodec
from _harness import BACKENDS, library
from _cases import random_dna

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SOURCES = [pytest.param("handle", id="BytesIO"), pytest.param("file", id="file"), pytest.param("path", id="path"),
           pytest.param("windowed", id="buffer_size")]


@pytest.fixture
def get_decoder(request, backend, source):
    handles = []

    def get(filename, **options):
        path = os.path.join(GOLDEN, filename)
        lib = library(backend)
        if source == "handle":                              # TestDecoderHandle._get_decoder (test_decoder.py:101-107)
            with open(path, "rb") as f:
                return nafcodec.Decoder(io.BytesIO(f.read()), _library=lib, **options)
        if source == "file":                                # TestDecoderFile._get_decoder (test_decoder.py:118-121)
            handles.append(open(path, "rb"))
            return nafcodec.Decoder(handles[-1], _library=lib, **options)
        if source == "path":
            return nafcodec.Decoder(path, _library=lib, **options)
        return nafcodec.Decoder(path, buffer_size=4096, _library=lib, **options)

    yield get
    for h in handles:                                       # TestDecoderFile.tearDown
        h.close()


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("source", SOURCES)
class TestDecoder:
    def test_fastq_optional(self, get_decoder):             # test_decoder.py:24-37
        decoder = get_decoder("phix.naf", id=False, sequence=False, comment=False)
        n = 0
        for record in decoder:
            assert record.id is None
            assert record.sequence is None
            assert record.comment is None
            assert record.quality is not None
            n += 1
        assert n == 42
        decoder = get_decoder("phix.naf", id=False, comment=False)
        for record in decoder:
            assert record.id is None
            assert record.comment is None
            assert record.sequence is not None
            assert record.quality is not None

    def test_len(self, get_decoder):                        # test_decoder.py:40-47
        decoder = get_decoder("phix.naf")
        assert len(decoder) == 42
        next(decoder)
        assert len(decoder) == 41
        records = list(decoder)
        assert len(records) == 41
        assert len(decoder) == 0

    def test_fastq(self, get_decoder):                      # test_decoder.py:50-59
        decoder = get_decoder("phix.naf")
        assert decoder.sequence_type == "dna"
        records = list(decoder)
        assert len(records) == 42
        assert records[0].id == "SRR1377138.1"
        assert records[0].sequence[:36] == "NGCTCTTAAACCTGCTATTGAGGCTTGTGGCATTTC"
        assert records[0].quality[:31] == "#8CCCGGGGGGGGGGGGGGGGGGGGGGGGGG"

    def test_dna(self, get_decoder):                        # test_decoder.py:62-72
        decoder = get_decoder("CP040672.naf")
        assert decoder.sequence_type == "dna"
        records = list(decoder)
        assert len(records) == 100
        assert records[0].id == "lcl|NZ_CP040672.1_cds_WP_044801954.1_1"
        assert records[0].sequence.count("A") == 181
        assert records[0].sequence.count("C") == 200
        assert records[0].sequence.count("G") == 210
        assert records[0].sequence.count("T") == 240
        assert records[0].quality is None

    def test_protein(self, get_decoder):                    # test_decoder.py:75-85
        decoder = get_decoder("LuxC.naf")
        assert decoder.sequence_type == "protein"
        records = list(decoder)
        assert len(records) == 12
        assert records[0].id == "sp|P19841|LUXC_PHOPO"
        assert records[0].sequence[:25] == "MCNAEFKGDCMIKKIPMIIGGAERD"
        assert records[0].quality is None
        assert records[5].id == "sp|P29236|LUXC2_PHOLE"
        assert records[5].sequence[:25] == "MIKKIPMIIGGVVQNTSGYGMRELT"
        assert records[5].quality is None

    def test_dna_masked(self, get_decoder):                 # test_decoder.py:88-98
        decoder = get_decoder("masked.naf")
        assert decoder.sequence_type == "dna"
        records = list(decoder)
        assert len(records) == 2
        assert records[0].id == "test1"
        assert records[0].sequence[:657].isupper()
        assert records[0].sequence[657:676].islower()
        assert records[0].sequence[676:1311].isupper()
        assert records[0].sequence[1311:1350].islower()
        assert records[0].quality is None


def test_error_filenotfound():                              # test_decoder.py:123-125
    with pytest.raises(FileNotFoundError):
        nafcodec.Decoder("")


@pytest.mark.skipif(os.name == "nt", reason="Windows error codes differ")
def test_error_isadirectory():                              # test_decoder.py:127-130
    with pytest.raises(IsADirectoryError):
        nafcodec.Decoder(os.path.dirname(__file__))
    with pytest.raises(IsADirectoryError):
        nafcodec.Decoder(os.path.dirname(__file__), buffer_size=4096)


@pytest.mark.parametrize("backend", BACKENDS)
class TestOpen:
    def test_open_read_fileobj(self, backend):              # test_open.py:23-28
        with open(os.path.join(GOLDEN, "LuxC.naf"), "rb") as f:
            with nafcodec.open(f, "r", _library=library(backend)) as decoder:
                assert isinstance(decoder, nafcodec.Decoder)
                assert len(decoder) == 12
                assert len(list(decoder)) == 12

    def test_open_read_filename(self, backend, tmp_path):   # test_open.py:31-39
        dst = tmp_path / "copy.naf"
        with open(os.path.join(GOLDEN, "LuxC.naf"), "rb") as f, open(dst, "wb") as g:
            shutil.copyfileobj(f, g)
        with nafcodec.open(str(dst), "r", _library=library(backend)) as decoder:
            assert isinstance(decoder, nafcodec.Decoder)
            assert len(decoder) == 12
            assert len(list(decoder)) == 12

    def _check_first(self, decoder):
        r = decoder.read()
        assert r.id == "r1"
        assert r.sequence is None
        assert r.comment is None
        assert r.quality is None
        assert r.length is None

    def test_open_write_fileobj(self, backend):             # test_open.py:41-55
        buffer = io.BytesIO()
        with nafcodec.open(buffer, "w", id=True, _library=library(backend)) as encoder:
            assert isinstance(encoder, nafcodec.Encoder)
            encoder.write(nafcodec.Record(id="r1"))
            encoder.write(nafcodec.Record(id="r2"))
            encoder.write(nafcodec.Record(id="r3"))
        buffer.seek(0)
        with nafcodec.open(buffer, "r", _library=library(backend)) as decoder:
            self._check_first(decoder)

    def test_open_write_filename(self, backend, tmp_path):  # test_open.py:57-70
        name = str(tmp_path / "out.naf")
        with nafcodec.open(name, "w", id=True, _library=library(backend)) as encoder:
            assert isinstance(encoder, nafcodec.Encoder)
            encoder.write(nafcodec.Record(id="r1"))
            encoder.write(nafcodec.Record(id="r2"))
            encoder.write(nafcodec.Record(id="r3"))
        with nafcodec.open(name, "r", _library=library(backend)) as decoder:
            self._check_first(decoder)
        with nafcodec.open(name, "r", buffer_size=1, _library=library(backend)) as decoder:
            self._check_first(decoder)


# ---- buffer_size: the records do not depend on it (mod.rs:104-112) ------------------------------------------------------

def _records_equal(a, b):
    return (a.id, a.comment, a.sequence, a.quality, a.length) == (b.id, b.comment, b.sequence, b.quality, b.length)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", ["phix.naf", "masked.naf", "LuxC.naf", "CP040672.naf", "NZ_AAEN01000029.naf"])
@pytest.mark.parametrize("buffer_size", [1, 300, 4096, 1 << 16, 1 << 30])
def test_windows_yield_the_same_records(backend, name, buffer_size):
    lib = library(backend)
    data = open(os.path.join(GOLDEN, name), "rb").read()
    whole = list(nafcodec.Decoder(io.BytesIO(data), _library=lib))
    dec = nafcodec.Decoder(io.BytesIO(data), buffer_size=buffer_size, _library=lib)
    assert len(dec) == len(whole)
    got = list(dec)
    assert len(got) == len(whole)
    for a, b in zip(got, whole):
        assert _records_equal(a, b)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("fields", [dict(sequence=False), dict(id=False, comment=False), dict(mask=False), dict(quality=False, id=False)])
def test_windows_with_skipped_fields(backend, fields):
    lib = library(backend)
    data = open(os.path.join(GOLDEN, "phix.naf"), "rb").read()
    whole = list(nafcodec.Decoder(io.BytesIO(data), _library=lib, **fields))
    got = list(nafcodec.Decoder(io.BytesIO(data), buffer_size=700, _library=lib, **fields))
    assert len(got) == len(whole) == 42
    for a, b in zip(got, whole):
        assert _records_equal(a, b)


@pytest.mark.parametrize("backend", BACKENDS)
def test_window_sizes_and_oracle(backend):
    """The C ABI directly: windows of a synthetic genome (records of very different sizes) against the oracle; a window never
    exceeds max_bytes unless it is a single record; first_bad_record / counts are relative to the window."""
    lib = library(backend)
    rng = np.random.default_rng(5)
    lens = [0, 1, 7, 50_000, 3, 0, 120_001, 64, 9_999, 2]
    seqs = [random_dna(rng, n, b"ACGTACGTNRY") for n in lens]
    data = O.encode(ids=[b"record-%d" % k for k in range(len(lens))], comments=[b"c" * (k * 37 % 200) for k in range(len(lens))],
                    sequences=seqs, mask_runs_=[100, 40_000, 77, 5, 300_000], level=3)
    d = O.decode(data)
    ctx = nafcodec.Context(0, lib)
    a = nafcodec.parse_archive(data, lib)
    ctx.prepare([a])
    ctx.run()
    for max_bytes in (0, 1, 1000, 60_000, 1 << 20):
        i = 0
        while i < d.n:
            w = ctx.fetch_window(0, i, d.n - i, max_bytes)
            assert w.n_records >= 1
            payload = len(w.sequence or b"") + len(w.ids or b"") + len(w.comments or b"")
            if max_bytes and w.n_records > 1:
                assert payload + 32 * w.n_records <= max_bytes
            if max_bytes == 0:
                assert w.n_records == d.n - i
            for j in range(w.n_records):
                assert w.id_bytes(j) == d.id(i + j)
                assert w.comment_bytes(j) == d.comment(i + j)
                assert w.length(j) == int(d.lengths[i + j])
                assert w.sequence_bytes(j) == d.seq(i + j)
            i += w.n_records
    # past the end: an empty window, no error
    w = ctx.fetch_window(0, d.n + 5, 10, 0)
    assert w.n_records == 0
    # a full fetch after windows still works (the whole-result buffer is taken on demand)
    full = ctx.fetch()[0]
    assert full.sequence == d.sequence
    ctx.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_window_reports_bad_utf8_at_the_same_record(backend):
    """reader.rs:108-109: the error is yielded AT the record whose text is not UTF-8, records before it are fine, also when
    the records arrive in windows."""
    lib = library(backend)
    ids = [b"ok%d" % k for k in range(9)]
    ids[6] = b"bad\xff\xfe"
    data = O.encode(ids=ids, sequences=[b"ACGT" * (k + 1) for k in range(9)])
    for bs in (1, 64, 1 << 20):
        dec = nafcodec.Decoder(io.BytesIO(data), buffer_size=bs, _library=lib)
        for k in range(6):
            assert next(dec).id == "ok%d" % k
        with pytest.raises(UnicodeError):
            next(dec)
        assert next(dec).id == "ok7"
