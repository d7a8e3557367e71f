"""Encode side: the device packer / length words / mask-run extraction against the oracle's restatement
of the reference encoder (oracle/naf_oracle.c: encoder/mod.rs:22-384, encoder/writer.rs), and the reference's own encoder
tests restated through the mirror (nafcodec/tests/encoder.rs:31-175, nafcodec-py/nafcodec/tests/test_encoder.py:21-86).
Both backends: "emul" (CPU tier) and "cuda" (-m gpu)."""
import io

import numpy as np
import pytest

import _cases as K
import _oracle as O
import nafcodec_b200 as N
from _harness import BACKENDS, library

pytestmark = pytest.mark.parametrize("backend", BACKENDS)


def _true_runs(seq: bytes):
    """Alternating unmasked / masked run lengths of a sequence (first run unmasked, possibly empty)."""
    if not seq:
        return []
    low = (np.frombuffer(seq, np.uint8) >= 97).astype(np.int8)
    flips = np.flatnonzero(np.diff(np.concatenate([[0], low]))).tolist()
    bounds = flips + [len(seq)]
    return [b - a for a, b in zip([0] + flips, bounds)]


def _random_records(rng, n, alphabet=b"ACGTNRYKMSWBDHV-", lower=0.0, max_len=3000):
    lens = [int(rng.choice([0, 1, 2, 3, 17, 21, 31, 32, 33, 255, 256, 1000])) if rng.random() < 0.5 else int(rng.integers(0, max_len)) for _ in range(n)]
    seqs = []
    for l in lens:
        a = rng.choice(np.frombuffer(alphabet, np.uint8), size=l).astype(np.uint8)
        if lower:
            k = 0
            while k < l:                                         # lower-case runs of geometric length
                run = int(rng.geometric(0.02))
                if rng.random() < lower:
                    seg = a[k:k + run]
                    seg[(seg >= 65) & (seg <= 90)] += 32
                k += run
        seqs.append(a.tobytes())
    return seqs


@pytest.mark.parametrize("seq_type", [O.DNA, O.RNA])
def test_pack_matches_the_reference_writer(backend, seq_type):
    # SequenceWriter: low nibble first, odd-length cache across records, padded last nibble; write_length words
    lib = library(backend)
    rng = np.random.default_rng(11 + seq_type)
    alphabet = b"ACGTNRYKMSWBDHV-" if seq_type == O.DNA else b"ACGUNRYKMSWBDHV-"
    for trial in range(4):
        seqs = _random_records(rng, int(rng.integers(1, 40)), alphabet)
        arc = O.encode(ids=[b"x"] * len(seqs), sequences=seqs, sequence_type=seq_type)
        packed, words, mask, _ = N.pack_sequences(seqs, seq_type, False, _library=lib)
        assert packed == O.section_bytes(arc, "sequence")
        assert words == O.section_bytes(arc, "length")
        assert mask is None


def test_pack_rejects_what_the_reference_rejects(backend):
    lib = library(backend)
    with pytest.raises(ValueError, match="residue 7"):           # writer.rs:49-52 "unexpected sequence character"
        N.pack_sequences([b"ACGT", b"ACGXA"], O.DNA, _library=lib)
    with pytest.raises(ValueError, match="residue 2"):           # lower case is rejected by SequenceWriter::encode
        N.pack_sequences([b"ACgT"], O.DNA, _library=lib)
    with pytest.raises(ValueError, match="residue 0"):           # 'U' in DNA, 'T' in RNA
        N.pack_sequences([b"UACG"], O.DNA, _library=lib)
    with pytest.raises(ValueError, match="residue 3"):
        N.pack_sequences([b"ACGT"], O.RNA, _library=lib)
    assert N.pack_sequences([], O.DNA, _library=lib)[:2] == (b"", b"")
    assert N.pack_sequences([b"", b""], O.DNA, _library=lib)[:2] == (b"", bytes(8))


def test_mask_extraction(backend):
    # the reference cannot write masks (encoder/mod.rs:240); the runs are checked against the Mask section the oracle writes
    # for the same runs, in the format MaskReader reads (reader.rs:196-231)
    lib = library(backend)
    rng = np.random.default_rng(5)
    cases = [[b"acgtACGT"], [b"ACGTacgt"], [b"a"], [b"A"], [b"AC", b"gt", b"", b"nnNN"],
             [b"A" * 254 + b"c" * 255 + b"G" * 256 + b"t" * 600 + b"A"], [bytes([97]) * 70000 + b"ACGT" * 100]]
    cases += [_random_records(rng, 25, b"ACGTN", lower=0.4) for _ in range(3)]
    cases.append(_random_records(rng, 3, b"ACGT", lower=0.5, max_len=200_000))       # several scan tiles
    for seqs in cases:
        cat = b"".join(seqs)
        runs = _true_runs(cat)
        upper = [s.upper() for s in seqs]
        arc = O.encode(ids=[b"x"] * len(seqs), sequences=upper, mask_runs_=runs)
        packed, words, mask, n_runs = N.pack_sequences(seqs, O.DNA, True, _library=lib)
        assert packed == O.section_bytes(arc, "sequence")
        assert mask == O.section_bytes(arc, "mask"), (len(cat), runs[:6])
        assert n_runs == len(runs)
        assert O.mask_runs(mask, len(cat)) == runs


def _roundtrip(backend, records, seq_type="dna", **fields):
    buf = io.BytesIO()
    with N.Encoder(buf, seq_type, _library=library(backend), **fields) as enc:
        for r in records:
            enc.write(r)
    return buf.getvalue()


@pytest.mark.parametrize("fields", [("id",), ("id", "sequence"), ("quality",), ("id", "comment", "sequence", "quality")])
def test_reference_roundtrips(backend, fields):
    # nafcodec/tests/encoder.rs:31-175: odd lengths 17 and 21 exercise the nibble cache; `length` is None when only ids are written
    R = K.reference_roundtrip_records()
    recs = [N.Record(id=R["ids"][i].decode(), comment=R["comments"][i].decode(), sequence=R["sequences"][i].decode(),
                     quality=R["qualities"][i].decode()) for i in range(2)]
    data = _roundtrip(backend, recs, **{f: True for f in fields})
    # byte-identical to the reference encoder's restatement (same zstd call pattern, same level)
    key = {"id": "ids", "comment": "comments", "sequence": "sequences", "quality": "qualities"}
    assert data == O.encode(**{key[f]: R[key[f]] for f in fields})
    got = list(N.Decoder(io.BytesIO(data), _library=library(backend)))
    assert len(got) == 2
    for g, r in zip(got, recs):
        assert g.id == (r.id if "id" in fields else None)
        assert g.comment == (r.comment if "comment" in fields else None)
        assert g.sequence == (r.sequence if "sequence" in fields else None)
        assert g.quality == (r.quality if "quality" in fields else None)
        assert g.length == (len(r.sequence) if ("sequence" in fields or "quality" in fields) else None)


def test_python_binding_encoder_tests(backend):
    # nafcodec-py/nafcodec/tests/test_encoder.py:21-86
    lib = library(backend)
    with pytest.raises(ValueError):
        N.Encoder(io.BytesIO(), sequence_type="dna", sequence=True, _library=lib).write(N.Record(sequence="hello world?!"))
    with pytest.raises(ValueError):
        N.Encoder(io.BytesIO(), sequence_type="dna", sequence=True, _library=lib).write(N.Record())
    with pytest.raises(ValueError):
        N.Encoder(io.BytesIO(), sequence_type="dna", sequence=True, _library=lib).write(N.Record(id="r1"))
    dna = [("r1", "ATTATTAGACAGAGC"), ("r2", "CTATTG"), ("r3", "TTAGTNNNNN")]
    data = _roundtrip(backend, [N.Record(id=i, sequence=s) for i, s in dna], id=True, sequence=True)
    recs = list(N.Decoder(io.BytesIO(data), _library=lib))
    assert [(r.id, r.sequence, r.quality, r.comment) for r in recs] == [(i, s, None, None) for i, s in dna]
    rna = [("r1", "AUUAU", "GGGGG"), ("r2", "CUAUU", "#8A@C"), ("r3", "UUAGU", "CCGGG")]
    buf = io.BytesIO()
    with N.open(buf, "w", sequence_type="rna", id=True, sequence=True, quality=True, _library=lib) as f:      # test_open.py: mode "w"
        for i, s, q in rna:
            f.write(N.Record(id=i, sequence=s, quality=q))
    recs = list(N.Decoder(io.BytesIO(buf.getvalue()), _library=lib))
    assert [(r.id, r.sequence, r.quality, r.comment) for r in recs] == [(i, s, q, None) for i, s, q in rna]
    assert N.Decoder(io.BytesIO(buf.getvalue()), _library=lib).sequence_type == "rna"


def test_archives_are_byte_identical_to_the_reference_encoder(backend):
    rng = np.random.default_rng(3)
    for level in (0, 3, 19):
        seqs = _random_records(rng, 30, b"ACGTN")
        ids = [b"rec%d" % i for i in range(len(seqs))]
        coms = [b"comment %d" % i if i % 2 else b"" for i in range(len(seqs))]
        quals = [bytes(rng.integers(33, 74, size=len(s)).astype(np.uint8)) for s in seqs]
        recs = [N.Record(id=i.decode(), comment=c.decode(), sequence=s.decode(), quality=q.decode()) for i, c, s, q in zip(ids, coms, seqs, quals)]
        data = _roundtrip(backend, recs, id=True, comment=True, sequence=True, quality=True, compression_level=level)
        assert data == O.encode(ids=ids, comments=coms, sequences=seqs, qualities=quals, level=level)
    prot = [b"MCNAEFKGDCMIKKIPMIIGGAERD", b"", b"MKK"]
    data = _roundtrip(backend, [N.Record(id="p%d" % i, sequence=s.decode()) for i, s in enumerate(prot)], "protein", id=True, sequence=True)
    assert data == O.encode(ids=[b"p0", b"p1", b"p2"], sequences=prot, sequence_type=O.PROTEIN)


def test_soft_mask_round_trip(backend):
    # An extension (the reference encoder has no mask writer).  Lower-case runs that end inside a record come back exactly; a
    # run that reaches the end of a record hits the reference DECODER's quirk (decoder/mod.rs:413-416: the tail is carried
    # over, not lower-cased), which the device decoder reproduces and the oracle pins.
    lib = library(backend)
    seqs = ["ACGTacgtACGTNNnnAC", "TTTTggggT", "acgtA", "ACGT"]
    data = _roundtrip(backend, [N.Record(id=f"m{i}", sequence=s) for i, s in enumerate(seqs)], id=True, sequence=True, mask=True)
    assert [r.sequence for r in N.Decoder(io.BytesIO(data), _library=lib)] == seqs
    cat = "".join(seqs).encode()
    assert data == O.encode(ids=[b"m0", b"m1", b"m2", b"m3"], sequences=[s.upper().encode() for s in seqs], mask_runs_=_true_runs(cat))
    tail = ["ACGTacgt", "ACGT"]                                  # the masked run ends with the record
    data = _roundtrip(backend, [N.Record(sequence=s) for s in tail], sequence=True, mask=True)
    want = O.decode(data)
    assert [r.sequence.encode() for r in N.Decoder(io.BytesIO(data), _library=lib)] == [want.seq(0), want.seq(1)] == [b"ACGTACGT", b"ACGT"]
