"""Pins the CPU oracle (oracle/naf_oracle.c) against every golden vector the reference's own tests
hold for the hot path.  CPU only.  Citations are reference file:line."""
import hashlib

import pytest

import _oracle as O
from conftest import read_golden


def parse_fasta(text: bytes):
    recs = []
    for line in text.splitlines():
        if line.startswith(b">"):
            recs.append([line[1:], b""])
        elif recs:
            recs[-1][1] += line.strip()
    return recs


def parse_fastq(text: bytes):
    lines = text.splitlines()
    return [(lines[i][1:], lines[i + 1], lines[i + 3]) for i in range(0, len(lines) - 3, 4)]


# ---- inline known answers -------------------------------------------------------------------

def test_variable_length_known_answers():
    # encoder/mod.rs:403-412
    for n, enc in [(0, "00"), (127, "7f"), (128, "8100"), (129, "8101"),
                   (34359738367, "ffffffff7f"), (34359738368, "818080808000")]:
        assert O.write_variable_length(n).hex() == enc
        assert O.variable_u64(bytes.fromhex(enc)) == (n, len(enc) // 2)


def test_header_known_answer():
    # decoder/parser.rs:141-152
    L = O.parse_header(bytes([0x01, 0xF9, 0xEC, 0x01, 0x3E, 0x20, 0x3C, 0x20]))
    assert chr(L.name_separator) == " " and L.line_length == 60 and L.number_of_sequences == 32


def test_error_empty():
    # decoder/mod.rs:470-476: empty input -> Io(UnexpectedEof)
    with pytest.raises(O.OracleError) as e:
        O.decode(b"")
    assert e.value.code == -1


def test_bad_magic():
    with pytest.raises(O.OracleError) as e:
        O.decode(b"\x01\xF9\xED\x01\x00 \x00\x00")
    assert e.value.code == -3


# ---- fixtures --------------------------------------------------------------------------------

def test_genome_fixture():
    # nafcodec/tests/decoder/dna.rs:8-34
    data = read_golden("NZ_AAEN01000029.naf")
    L = O.parse(data)
    assert (chr(L.name_separator), L.number_of_sequences, L.line_length, L.sequence_type, L.format_version) == (" ", 30, 80, O.DNA, 1)
    assert L.flags == 0x3E
    sizes = {n: (L.sec[i].original_size, L.sec[i].compressed_size, L.sec[i].offset) for i, n in enumerate(O.SEC_NAMES) if L.sec[i].present}
    assert sizes == {"id": (540, 122, 11), "comment": (2308, 212, 137), "length": (120, 120, 351),
                     "mask": (21525, 15, 475), "sequence": (5488676, 1330710, 497)}
    d = O.decode(data)
    assert d.n == 30
    assert d.id(0) == b"NZ_AAEN01000029.1"
    assert d.comment(0) == b"Bacillus anthracis str. CNEVA-9066 map unlocalized plasmid pXO1 cont2250, whole genome shotgun sequence"
    s = d.seq(0)
    assert len(s) == 182777
    assert (s.count(b"A"), s.count(b"C"), s.count(b"G"), s.count(b"T")) == (62115, 28747, 30763, 61152)
    assert d.id(1) == b"NZ_AAEN01000030.3"
    assert d.comment(1) == b"Bacillus anthracis str. CNEVA-9066 map unlocalized plasmid pXO2 cont2251, whole genome shotgun sequence"
    assert d.lengths.tolist() == [182777, 95646, 1087, 265902, 145793, 179124, 136786, 277207, 6997, 24957, 22349, 97690,
                                  258793, 677301, 1053408, 38198, 200528, 506449, 258357, 271364, 403539, 40701, 228247,
                                  8945, 49711, 339, 43309, 4349, 6285, 2538]
    assert hashlib.sha256(d.sequence).hexdigest() == "84242bd01d97b877141329b7283ddbf93414f6ce8e7981ec3b6eb61b9ec6f90b"
    assert d.quality == b"" and not d.qual_present.any()


def test_masked_fixture():
    # nafcodec/tests/decoder/dna.rs:36-63, decoder/mod.rs:486-504
    data = read_golden("masked.naf")
    L = O.parse(data)
    assert (L.number_of_sequences, L.line_length) == (2, 50)
    mask = O.section_bytes(data, "mask")
    assert mask.hex() == "ffff9313ffff7d27ffffd760630dae0effffff72"
    assert O.mask_runs(mask, 3350)[:5] == [657, 19, 635, 39, 725]
    d = O.decode(data)
    assert d.id(0) == b"test1" and d.id(1) == b"test2"
    s0, s1 = d.seq(0), d.seq(1)
    assert len(s0) == 1550 and len(s1) == 1800
    assert s0[:657].isupper() and s0[657:676].islower() and s0[676:1311].isupper() and s0[1311:1350].islower()
    assert s1[:525].isupper() and s1[525:621].islower() and s1[621:720].isupper() and s1[720:733].islower()
    assert hashlib.sha256(d.sequence).hexdigest() == "c921ec989ea0cd2c43789bdb86db0c980698df1ef61271efef2e4060d5e1b6ce"
    fa = parse_fasta(read_golden("masked.fna"))
    assert [r[1] for r in fa] == [s0, s1]
    # dna.rs:65-88 force_nomask
    d2 = O.decode(data, mask=False)
    assert d2.seq(0).isupper() and d2.seq(1).isupper()
    assert d2.sequence == d.sequence.upper()


def test_phix_fixture():
    # nafcodec/tests/decoder/fastq.rs:16-118
    data = read_golden("phix.naf")
    L = O.parse(data)
    assert L.number_of_sequences == 42 and L.flags == 0x3F and L.line_length == 301
    d = O.decode(data)
    assert d.id(0) == b"SRR1377138.1"
    assert d.comment(0) == b"a comment that should not be included in the SAM output"
    assert d.seq(0).startswith(b"NGCTCTTAAACCTGCTATTGAGGCTTGTGGCATTTC")
    assert d.qual(0).startswith(b"#8CCCGGGGGGGGGGGGGGGGGGGGGGGGGG")
    assert d.id(1) == b"SRR1377138.2" and d.comment(1) == b"some lowercase nucleotides"
    assert d.lengths.tolist() == [301] * 40 + [95, 301]          # == line lengths of data/phix.fastq (sum 12436)
    assert hashlib.sha256(d.sequence).hexdigest() == "31adb5c8cf3806ece7b87e044d68faf5fef9180aee20608b3313b973fca83915"
    assert hashlib.sha256(d.quality).hexdigest() == "1ed7cb3cdcc2bf223f9eec7b130dd99ee20f44d1488f98af807ac37fefea75c1"
    fq = parse_fastq(read_golden("phix.fastq"))
    assert len(fq) == 42
    for i, (name, seq, qual) in enumerate(fq):
        assert name == (d.id(i) + b" " + d.comment(i)).rstrip(b" ")     # empty comments: no separator in the text
        assert seq == d.seq(i) and qual == d.qual(i)
    for field in ("id", "sequence", "comment", "quality"):
        dd = O.decode(data, **{field: False})
        assert dd.n == 42
        pres = {"id": dd.id_present, "sequence": dd.seq_present, "comment": dd.com_present, "quality": dd.qual_present}[field]
        assert not pres.any()
        others = {"id": dd.id_present, "sequence": dd.seq_present, "comment": dd.com_present, "quality": dd.qual_present}
        assert all(v.all() for k, v in others.items() if k != field)
    # nafcodec-py test_decoder.py:25-37
    dd = O.decode(data, id=False, sequence=False, comment=False)
    assert not dd.id_present.any() and not dd.seq_present.any() and dd.qual_present.all()


def test_cp040672_fixture():
    # nafcodec-py/nafcodec/tests/test_decoder.py:61-72
    d = O.decode(read_golden("CP040672.naf"))
    assert d.n == 100
    assert d.id(0) == b"lcl|NZ_CP040672.1_cds_WP_044801954.1_1"
    s = d.seq(0)
    assert len(s) == 831 and (s.count(b"A"), s.count(b"C"), s.count(b"G"), s.count(b"T")) == (181, 200, 210, 240)
    assert hashlib.sha256(d.sequence).hexdigest() == "c3bc2d8e85b8429262076a711e9953a5ac84d596adbd3acdbe5fcaf02d926a8a"
    assert not d.qual_present.any()


def test_luxc_fixture():
    # nafcodec/tests/decoder/protein.rs:4-22, decoder/mod.rs:478-483,506-515, test_decoder.py:74-85
    data = read_golden("LuxC.naf")
    L = O.parse(data)
    assert (L.format_version, L.sequence_type, L.number_of_sequences, L.line_length, L.flags) == (2, O.PROTEIN, 12, 60, 0x3A)
    d = O.decode(data)
    assert d.n == 12 and d.id(0) == b"sp|P19841|LUXC_PHOPO" and len(d.seq(0)) == 488
    assert d.seq(0).startswith(b"MCNAEFKGDCMIKKIPMIIGGAERD")
    assert d.id(5) == b"sp|P29236|LUXC2_PHOLE" and d.seq(5).startswith(b"MIKKIPMIIGGVVQNTSGYGMRELT")
    assert hashlib.sha256(d.sequence).hexdigest() == "b3dd0e7c601e2e0d925a7b8d70a157b3783f5742295914843e22f7dd2df5794f"
    fa = parse_fasta(read_golden("LuxC.faa"))
    assert [r[1] for r in fa] == [d.seq(i) for i in range(12)]
    assert not O.decode(data, sequence=False).seq_present.any()


# ---- encoder restatement round trips (nafcodec/tests/encoder.rs:12-175) ----------------------

RECS = dict(ids=[b"r1", b"r2"], comments=[b"record 1", b"record 2"],
            sequences=[b"NGCTCTTAAACCTGCTA", b"NTAATAAGCAATGACGGCAGC"],
            qualities=[b"#8CCCGGGGGGGGGGGG", b"#8AACCFF<FFGGFGE@@@@@"])


@pytest.mark.parametrize("fields", [("ids",), ("ids", "sequences"), ("qualities",), ("ids", "comments", "sequences", "qualities")])
@pytest.mark.parametrize("flush", [True, False])
def test_encoder_roundtrip(fields, flush):
    kw = {k: RECS[k] for k in fields}
    data = O.encode(flush_per_record=flush, **kw)
    d = O.decode(data)
    assert d.n == 2
    for i in range(2):
        assert d.id(i) == (RECS["ids"][i] if "ids" in fields else None)
        assert d.comment(i) == (RECS["comments"][i] if "comments" in fields else None)
        assert d.seq(i) == (RECS["sequences"][i] if "sequences" in fields else None)
        assert d.qual(i) == (RECS["qualities"][i] if "qualities" in fields else None)
        if "sequences" in fields or "qualities" in fields:
            assert d.length(i) == len(RECS["sequences"][i])
        else:
            assert d.length(i) is None          # tests/encoder.rs:55,62


def test_encoder_rna_and_invalid():
    # nafcodec-py test_encoder.py:21-25,36-86
    data = O.encode(sequences=[b"ACGU", b"UUGCA"], qualities=[b"IIII", b"IIIII"], sequence_type=O.RNA)
    L = O.parse(data)
    assert L.format_version == 2 and L.sequence_type == O.RNA
    d = O.decode(data)
    assert d.seq(0) == b"ACGU" and d.seq(1) == b"UUGCA"
    with pytest.raises(O.OracleError) as e:
        O.encode(sequences=[b"ACGX"])
    assert e.value.code == -5
    with pytest.raises(O.OracleError):
        O.encode(sequences=[b"acgt"])           # lowercase is rejected (encoder/writer.rs:31-54)


def test_mask_quirk_tail_not_lowercased():
    """decoder/mod.rs:413-416: a masked unit that reaches the end of a record is carried over WITHOUT
    lower-casing the record's tail.  Pinned because parity is against the reference implementation."""
    seqs = [b"ACGTACGTAC", b"GGGGGGGGGG", b"TTTTTTTTTT"]
    # runs: U4 M3 U1 M12 (covers tail of r0 [8,10), all of r1, [20,22)... ) U10
    data = O.encode(ids=[b"a", b"b", b"c"], sequences=seqs, mask_runs_=[4, 3, 1, 12, 10])
    d = O.decode(data)
    assert d.seq(0) == b"ACGTacgTAC"       # [8,10) inside M12 stays upper: unit reaches the record end
    assert d.seq(1) == b"GGGGGGGGGG"       # fully covered record: untouched
    assert d.seq(2) == b"TTTTTTTTTT"       # remainder Masked(0) then U10
    data = O.encode(ids=[b"a", b"b", b"c"], sequences=seqs, mask_runs_=[4, 3, 1, 14, 8])
    d = O.decode(data)
    assert d.seq(2) == b"ttTTTTTTTT"       # remainder Masked(2) < 10 -> prefix lower-cased


def test_synthetic_mask_roundtrip():
    n = 200000
    seq = O.synth_dna(7, n)
    runs = O.synth_mask(7, n)
    assert sum(runs) == n and runs[0] == 0 and 255 in runs and 510 in runs and 70000 in runs
    data = O.encode(ids=[b"synth"], comments=[b"c"], sequences=[seq], mask_runs_=runs, level=3)
    d = O.decode(data)
    expect = O.apply_mask(seq, runs)
    if len(runs) % 2 == 0:                    # final run is masked and reaches the record end: stays upper (quirk)
        tail = runs[-1]
        expect = expect[:n - tail] + expect[n - tail:].upper()
    assert d.seq(0) == expect
    assert O.decode(data, mask=False).seq(0) == seq


def test_lengths_continuation_words():
    # reader.rs:46-68 / encoder/mod.rs:37-44: 0xFFFFFFFF words continue the sum. Exercised on the raw reader via
    # a hand-built Length section (no 4 GiB sequence needed): ids only + crafted length words is not reachable
    # through the encoder, so check the word arithmetic through a crafted archive without Sequence flag.
    import struct
    words = struct.pack("<IIII", 0xFFFFFFFF, 5, 7, 0xFFFFFFFF)          # -> [4294967300, 7], trailing FFFFFFFF => None
    body = O.encode(ids=[b"x", b"y", b"z"])
    L = O.parse(body)
    # rebuild: flags Id|Length, sections id + length(raw zstd frame made by the oracle encoder path)
    lens_frame = O.encode(qualities=[b""])                              # any archive; we only need a zstd frame maker
    del lens_frame
    import ctypes as C
    # make a magicless frame for `words` by encoding them as a quality stream of a 1-record archive
    arch = O.encode(qualities=[words.replace(b"\x00", b"\x01")])
    del arch
    # simplest faithful path: text archive whose quality IS the words; checked in test_host_lengths for the GPU path
    ids_sec = body[L.sec[0].offset:L.sec[0].offset + L.sec[0].compressed_size]
    q = O.encode(sequence_type=O.TEXT, sequences=[words])               # TEXT sequence section = raw bytes frame
    Lq = O.parse(q)
    frame = q[Lq.sec[4].offset:Lq.sec[4].offset + Lq.sec[4].compressed_size]
    hdr = bytes([0x01, 0xF9, 0xEC, 0x01, 0x28, 0x20]) + O.write_variable_length(60) + O.write_variable_length(3)
    arc = hdr + O.write_variable_length(L.sec[0].original_size) + O.write_variable_length(len(ids_sec)) + ids_sec \
        + O.write_variable_length(len(words)) + O.write_variable_length(len(frame)) + frame
    d = O.decode(arc)
    assert [d.length(i) for i in range(3)] == [4294967300, 7, None]
    assert [d.id(i) for i in range(3)] == [b"x", b"y", b"z"]
