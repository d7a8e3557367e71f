"""Parity of the device path with the CPU oracle, bit-exact, through the C ABI.

Every test runs twice: backend "emul" (CPU tier: the same kernel sources compiled against tests/emul's SIMT emulator)
and backend "cuda" (`-m gpu`: the real sm_100a library on a B200).  Sizes differ per backend."""
import hashlib

import numpy as np
import pytest

import _cases as K
import _oracle as O
import nafcodec_b200 as N
from _harness import BACKENDS, assert_same_as_oracle, check_parity, decode_soa, library, records
from conftest import read_golden

pytestmark = pytest.mark.parametrize("backend", BACKENDS)

FIXTURES = ["masked.naf", "LuxC.naf", "phix.naf", "CP040672.naf", "NZ_AAEN01000029.naf"]
SHA = {"NZ_AAEN01000029.naf": "84242bd01d97b877141329b7283ddbf93414f6ce8e7981ec3b6eb61b9ec6f90b",
       "masked.naf": "c921ec989ea0cd2c43789bdb86db0c980698df1ef61271efef2e4060d5e1b6ce",
       "phix.naf": "31adb5c8cf3806ece7b87e044d68faf5fef9180aee20608b3313b973fca83915",
       "CP040672.naf": "c3bc2d8e85b8429262076a711e9953a5ac84d596adbd3acdbe5fcaf02d926a8a",
       "LuxC.naf": "b3dd0e7c601e2e0d925a7b8d70a157b3783f5742295914843e22f7dd2df5794f"}


@pytest.mark.parametrize("name", FIXTURES)
def test_fixture_all_fields(backend, name):
    res, _ = check_parity(backend, read_golden(name), name)
    assert hashlib.sha256(res.sequence).hexdigest() == SHA[name]          # SURVEY 8c golden digests


@pytest.mark.parametrize("name", ["phix.naf", "masked.naf", "LuxC.naf"])
@pytest.mark.parametrize("skip", ["id", "comment", "sequence", "quality", "mask"])
def test_fixture_field_skipping(backend, name, skip):
    # tests/decoder/fastq.rs:55-118, dna.rs:65-88, decoder/mod.rs:506-515
    check_parity(backend, read_golden(name), f"{name} -{skip}", **{skip: False})


def test_fixture_zstd_boundary(backend):
    """Every section of every fixture: device zstd == libzstd (the pure-zstd boundary the reference never tests alone)."""
    ctx = N.shared_context(0, library(backend))
    for name in FIXTURES:
        data = read_golden(name)
        L = O.parse(data)
        for i in range(6):
            s = L.sec[i]
            if s.present:
                frame = data[s.offset:s.offset + s.compressed_size]
                want = O.zstd_decompress(frame)
                assert ctx.zstd_decompress(frame, len(want)) == want, (name, O.SEC_NAMES[i])


@pytest.mark.parametrize("fields", [("ids",), ("ids", "sequences"), ("qualities",), ("ids", "comments", "sequences", "qualities")])
@pytest.mark.parametrize("flush", [True, False])
def test_reference_roundtrip_records(backend, fields, flush):
    # nafcodec/tests/encoder.rs:31-175 (odd lengths 17 and 21 exercise the nibble carry)
    R = K.reference_roundtrip_records()
    data = O.encode(flush_per_record=flush, **{k: R[k] for k in fields})
    check_parity(backend, data, str(fields))
    recs = records(backend, data)
    assert len(recs) == 2
    for i, r in enumerate(recs):
        assert r.id == (R["ids"][i].decode() if "ids" in fields else None)
        assert r.sequence == (R["sequences"][i].decode() if "sequences" in fields else None)
        assert r.quality == (R["qualities"][i].decode() if "qualities" in fields else None)
        assert r.length == (len(R["sequences"][i]) if ("sequences" in fields or "qualities" in fields) else None)


def test_mask_quirk_and_edges(backend):
    # decoder/mod.rs:402-441 incl. the record-tail quirk (413-416), zero-length units, units spanning several records
    seqs = [b"ACGTACGTAC", b"GGGGGGGGGG", b"TTTTTTTTTT", b"", b"A", b"CCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCC"]
    ids = [b"a", b"b", b"c", b"d", b"e", b"f"]
    for runs in ([4, 3, 1, 12, 10, 100], [4, 3, 1, 14, 8, 100], [0, 5, 0, 0, 5, 0, 3, 200], [72], [0, 72], [10, 0, 10, 0, 10, 0, 10, 50],
                 [9, 1, 9, 1, 9, 1, 1, 41], [31, 1, 32, 8], [255, 255], [1] * 72):
        data = O.encode(ids=ids, sequences=seqs, mask_runs_=runs)
        check_parity(backend, data, f"runs {runs}")
        check_parity(backend, data, f"runs {runs} nomask", mask=False)


def test_dense_mask_spans_several_scan_slices(backend):
    """> 32 KiB of mask bytes: k_naf_scan cuts the mask into slices, one CTA each, every slice deriving its carry-in
    (residue position, run index) from the bytes before it; runs of a few residues mixed with runs > 255 (0xFF
    continuation bytes across slice boundaries)."""
    rng = np.random.default_rng(11)
    n = 360_000 if backend == "emul" else 3_000_000
    runs, total = [], 0
    while total < n:
        r = int(rng.integers(1, 5)) if rng.random() > 0.002 else int(rng.integers(250, 2000))
        r = min(r, n - total)
        runs.append(r); total += r
    seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n))
    cuts = sorted(set(int(x) for x in rng.integers(1, n, size=6)))
    cuts = [0] + cuts + [n]
    seqs = [seq[cuts[i]:cuts[i + 1]] for i in range(len(cuts) - 1)]
    arc = O.encode(ids=[b"r%d" % i for i in range(len(seqs))], sequences=seqs, mask_runs_=runs, level=3)
    assert O.parse(arc).sec[3].original_size > 2 * 32768
    check_parity(backend, arc, "dense mask")


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("level,flush", [(0, True), (3, False), (19, True)])
def test_multi_record_dna(backend, seed, level, flush):
    n, mx = (60, 3000) if backend == "emul" else (400, 20000)
    data = K.multi_record_dna(seed, n, mx, level=level, flush=flush, empty_every=7)
    check_parity(backend, data, f"seed {seed}")
    check_parity(backend, data, f"seed {seed} nomask", mask=False)


def test_fastq_tiny_blocks(backend):
    # cfg4 shape at a size the oracle finishes in seconds: one zstd block per record per stream
    n = 300 if backend == "emul" else 20000
    for lvl in (0, 19) if backend == "emul" else (0, 3):
        data = K.fastq_reads(5, n, level=lvl, with_mask=True)
        check_parity(backend, data, f"fastq level {lvl}")
        check_parity(backend, data, f"fastq level {lvl} -quality", quality=False)


def test_genome_with_gaps(backend):
    # N stretches: RLE blocks and offset-1 matches with huge lengths (SURVEY 8a note)
    n = 400_000 if backend == "emul" else 3_000_000
    data = K.genome(11, n, level=19, gaps=3, gap_len=n // 8, telomere=10_000, records=3)
    check_parity(backend, data, "gaps")


def test_text_protein_rna_and_masked_text(backend):
    prot = [b"MCNAEFKGDCMIKKIPMIIGGAERD" * 7, b"MIKKIPMIIGGVVQNTSGYGMRELT" * 3, b""]
    check_parity(backend, K.text_archive(prot, O.PROTEIN), "protein")
    check_parity(backend, K.text_archive(prot, O.PROTEIN, mask_runs=[10, 20, 30, 40, 1000]), "protein masked")
    check_parity(backend, K.text_archive([b"hello WORLD, this IS text", b"MORE text"], O.TEXT, mask_runs=[6, 5, 3, 100]), "text masked")
    check_parity(backend, O.encode(sequences=[b"ACGU", b"UUGCANNRY"], qualities=[b"IIII", b"IIIIIIIII"], sequence_type=O.RNA), "rna")
    utf = "séquence naïve ☃ 𝄞".encode()
    check_parity(backend, O.encode(ids=[utf, b"plain"], comments=[b"x", utf], sequences=[b"ACGT", b"TTGA"]), "utf8 ids")
    check_parity(backend, K.text_archive([utf, b"abc"], O.TEXT), "utf8 text")


def test_invalid_utf8_raises_at_the_record(backend):
    # reader.rs:108-109: String::from_utf8 per record -> error at THAT record; earlier records are fine
    data = K.text_archive([b"good", b"bad \xff\xfe", b"later"], O.TEXT)
    dec = N.Decoder(__import__("io").BytesIO(data), _library=library(backend))
    assert dec.read().sequence == "good"
    with pytest.raises(N.NafUnicodeError):
        dec.read()
    with pytest.raises(O.OracleError):
        O.decode(data)
    # a multi-byte character split across two records is invalid in both
    e = "é".encode()
    data = K.text_archive([b"ab" + e[:1], e[1:] + b"cd"], O.TEXT)
    with pytest.raises(N.NafUnicodeError):
        records(backend, data)


def test_lengths_continuation_words(backend):
    import struct
    words = struct.pack("<IIII", 0xFFFFFFFF, 5, 7, 0xFFFFFFFF)      # -> [4294967300, 7], dangling continuation -> None
    data = K.crafted_lengths_archive(words, [b"x", b"y", b"z"])
    res, d = check_parity(backend, data, "continuation")
    assert [res.length(i) for i in range(3)] == [4294967300, 7, None]


def test_zstd_boundary_generated(backend):
    """Frames at several levels over data shapes that hit every block / literal / sequence mode."""
    ctx = N.shared_context(0, library(backend))
    rng = np.random.default_rng(42)
    big = 300_000 if backend == "emul" else 2_000_000
    text = (b"lcl|NZ_CP040672.1_cds_WP_%09d.1_%d [gene=abc%d] [protein=hypothetical protein] [location=%d..%d]\n")
    payloads = {
        "empty": b"", "one": b"x", "zeros": bytes(70000), "short_rle": b"a" * 40,
        "uniform16": bytes(rng.integers(0, 16, size=50000).astype(np.uint8)),
        "uniform64": bytes(rng.integers(0, 64, size=50000).astype(np.uint8)),
        "uniform256": bytes(rng.integers(0, 256, size=20000).astype(np.uint8)),
        "skewed": bytes(rng.choice(np.arange(8, dtype=np.uint8), p=[.5, .2, .1, .08, .06, .03, .02, .01], size=big // 2)),
        "text": b"".join(text % (i, i, i % 97, i * 13, i * 13 + 700) for i in range(big // 110)),
        "period3": b"abc" * 30000, "mixed": bytes(rng.integers(0, 4, size=big // 3).astype(np.uint8)) + bytes(200000) + b"xyz" * 5000,
    }
    for name, p in payloads.items():
        for level in (1, 3, 9, 19):
            if backend == "emul" and level == 9:
                continue
            frame = K.zstd_frame(p, level)
            assert O.zstd_decompress(frame) == p
            assert ctx.zstd_decompress(frame, len(p)) == p, (name, level)
        frame = K.zstd_frame(p, 3, flush_every=997)                    # many small blocks, tables carried by repeat modes
        assert ctx.zstd_decompress(frame, len(p)) == p, (name, "flushed")


def _generations(rng, gens, blen, far_every=0, long_every=0):
    """Text in which block k is block k-1 with one byte changed: every match feeds the next one (one dependency chain),
    which is what quality strings and ids look like to the LZ stage.  `far_every` / `long_every` splice in references
    far behind the current position and long matches."""
    b = bytearray(rng.integers(33, 75, size=blen).astype(np.uint8).tobytes())
    first = bytes(rng.integers(33, 75, size=5000).astype(np.uint8))
    q = bytearray(first)
    for g in range(gens):
        b[int(rng.integers(blen))] = int(rng.integers(33, 75))
        q += b
        if far_every and g % far_every == far_every - 1:
            k = int(rng.integers(0, 4000)); q += first[k:k + int(rng.integers(8, 600))]     # far reference
        if long_every and g % long_every == long_every - 1:
            q += bytes(b[:7]) * 4000                                                          # 28 kB periodic match
    return bytes(q)


@pytest.mark.parametrize("lz_small", ["8192", "0"])
def test_ordered_finisher_on_dependency_chains(backend, monkeypatch, lz_small):
    """Sections that are one long LZ dependency chain are handed to k_lz_finish (shared-memory ring); the job stats
    say so, and the bytes must still be libzstd's.  Small jobs run their rounds in one CTA (k_lz_small: up to 8192 matches,
    2048 by default), the others in k_lz_first / k_lz_resolve: both hand over here (frames of 3000 to 40000 matches)."""
    monkeypatch.setenv("NAFGPU_LZ_SMALL", lz_small)
    ctx = N.shared_context(0, library(backend))
    rng = np.random.default_rng(7)
    gens = 3000 if backend == "emul" else 40000
    for blen, far, lng, level in [(100, 0, 0, 3), (300, 0, 0, 19), (100, 50, 0, 3), (150, 40, 700, 19), (3000, 3, 0, 3)]:
        p = _generations(rng, gens if blen < 1000 else gens // 20, blen, far, lng)
        frame = K.zstd_frame(p, level)
        assert O.zstd_decompress(frame) == p
        assert ctx.zstd_decompress(frame, len(p)) == p, (blen, far, lng, level)
        if blen < 1000:                                    # (few, long matches: the rounds stay cheaper than the finisher)
            assert ctx.stats().lz_handover > 0, (blen, far, lng, level)
            assert ctx.stats().lz_unresolved > 0, "chains cross the 64 KB chunks: the second level has work"
    # a chain next to ordinary frames in one job: quality of a FASTQ archive whose reads descend from one another
    qual = _generations(rng, 1500, 150)[5000:]
    n = len(qual) // 150
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=150)) for _ in range(n)]
    arc = O.encode(sequence_type=O.DNA, ids=[b"r%d" % i for i in range(n)], sequences=seqs,
                   qualities=[qual[i * 150:(i + 1) * 150] for i in range(n)], level=3)
    check_parity(backend, arc)
    assert ctx.stats().lz_handover > 0


# libzstd checks that the device path does NOT replicate: none known.  (300 single-bit flips of this archive: 222 decode
# identically on both sides, 77 are rejected by the device, 1 yields invalid UTF-8 that both report at the same record.)
# The list is closed: a flip that libzstd rejects while the device accepts it fails the test unless its message is listed.
_NOT_REPLICATED = ()


def test_corrupt_input_never_hangs(backend):
    """Truncations and byte flips: the device path must return (error or data), never hang or fault."""
    data = bytearray(K.multi_record_dna(9, 20, 2000, level=3))
    lib = library(backend)
    rng = np.random.default_rng(1)
    L = O.parse(bytes(data))
    start = L.sec[0].offset
    accepted_where_libzstd_rejects = 0
    trials = 60
    for trial in range(trials):
        bad = bytearray(data)
        pos = int(rng.integers(start, len(bad)))
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            got = N.shared_context(0, lib).decode([N.parse_archive(bytes(bad), lib)])[0]
        except (N.NafError, ValueError):
            continue
        try:
            d = O.decode(bytes(bad))
        except O.OracleError as e:
            if e.code == -4:                                   # invalid UTF-8: the device reports it at the record (reader.rs:108-109)
                assert got.first_bad_record is not None, f"flip at {pos}: {e}"
                continue
            assert any(k in str(e) for k in _NOT_REPLICATED), f"flip at {pos}: device accepted, libzstd says: {e}"
            accepted_where_libzstd_rejects += 1
            continue
        # both accepted the mutated archive: they must agree
        from _harness import assert_same_as_oracle
        assert_same_as_oracle(got, d, f"flip at {pos}")
    assert accepted_where_libzstd_rejects == 0
    for cut in (len(data) - 1, len(data) // 2, start + 3):
        with pytest.raises((N.NafError, ValueError)):
            N.shared_context(0, lib).decode([N.parse_archive(bytes(data[:cut]), lib)])


def test_batch_of_archives(backend):
    """Independent archives in one set of launches (cfg5 shape, small)."""
    lib = library(backend)
    n = 5 if backend == "emul" else 24
    size = 30_000 if backend == "emul" else 400_000
    arcs = [K.genome(100 + i, size + 1111 * i, level=3 if i % 2 else 19, records=1 + i % 3) for i in range(n)]
    arcs.append(read_golden("phix.naf"))
    arcs.append(read_golden("LuxC.naf"))
    res = N.decode_batch(arcs, _library=lib)
    from _harness import assert_same_as_oracle
    for i, (r, a) in enumerate(zip(res, arcs)):
        assert_same_as_oracle(r, O.decode(a), f"archive {i}")


def test_pipeline_lanes(backend):
    """Several contexts driven from host threads (the e2e path of bench.py) give the same results as one context."""
    lib = library(backend)
    arcs = [K.genome(300 + i, 25_000 + 999 * i, level=3) for i in range(5)] + [read_golden("phix.naf")]
    arcs += [K.fastq_reads(9 + i, 1200 if backend == "emul" else 20000) for i in range(2)]      # lanes that take the finisher path at the same time
    parsed = [N.parse_archive(a, lib) for a in arcs]
    pipe = N.Pipeline(0, 3, lib)
    from _harness import assert_same_as_oracle
    got = {}

    def consume(bi, i, r):                                   # runs on the lane's thread while its pinned buffers are valid
        got[(bi, i)] = N.ArchiveResult._copy_from(parsed[i].header, r)

    try:
        res = pipe.decode(parsed)
        n = pipe.decode_stream([parsed, parsed[:4], parsed], consume=consume, sub_batch=2)     # no barrier between batches
    finally:
        pipe.close()
    want = [O.decode(a) for a in arcs]
    for i, r in enumerate(res):
        assert_same_as_oracle(r, want[i], f"archive {i}")
    assert n == 20 and len(got) == 20
    for (bi, i), r in got.items():
        assert_same_as_oracle(r, want[i], f"stream batch {bi} archive {i}")


def test_long_frames_are_scanned_by_tiles(backend, monkeypatch):
    """Frames with very many blocks (a FASTQ section flushed per record) get their block offsets and repeat-offset carries from
    a tiled scan (k_fs_reduce / k_fs_prefix / k_fs_apply); small tiles here so that the path runs on a modest archive."""
    monkeypatch.setenv("NAFGPU_FS_TILE", "37")
    lib = library(backend)
    arc = K.fastq_reads(21, 700 if backend == "emul" else 30000, with_mask=True)
    res, d = check_parity(backend, arc, "tiled frame scan")
    assert N.shared_context(0, lib).stats().n_blocks > 1400
    monkeypatch.delenv("NAFGPU_FS_TILE")
    check_parity(backend, arc, "default tiles")


@pytest.mark.parametrize("part", ["1", "5", "64"])
def test_long_frames_are_walked_in_parts(backend, monkeypatch, part):
    """A section of 2 x 10^5+ blocks is walked in parts on several host threads after a pass over the chain of block headers
    alone (frame_walk.h); a part inherits the last Huffman tree (treeless literals) and the last FSE tables (repeat mode) from
    the parts before it.  Parts of 1, 5 and 64 blocks here, so that every block, or nearly, inherits across a cut; the result
    must be what the walk in one piece gives (NAFGPU_WALK_SPLIT=0) and what the oracle gives."""
    lib = library(backend)
    arcs = [K.fastq_reads(31, 200 if backend == "emul" else 20000, with_mask=True),
            K.fastq_reads(32, 120 if backend == "emul" else 5000, level=19),
            K.multi_record_dna(33, 40, 2000, level=3, flush=True)]
    monkeypatch.setenv("NAFGPU_WALK_SPLIT", "0")
    whole = [decode_soa(lib, a) for a in arcs]
    monkeypatch.setenv("NAFGPU_WALK_SPLIT", part)
    for a, w in zip(arcs, whole):
        res, d = check_parity(backend, a, f"parts of {part} blocks")
        assert res.sequence == w.sequence and res.quality == w.quality and res.ids == w.ids
    # the pure-zstd boundary: frames whose blocks repeat tables and trees of earlier blocks
    rng = np.random.default_rng(7)
    payload = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 300_000))
    frame = K.zstd_frame(payload, 3, flush_every=997)
    assert N.shared_context(0, lib).zstd_decompress(frame, len(payload)) == payload
    # a batch: every archive's sections are cut independently
    got = N.shared_context(0, lib).decode([N.parse_archive(a, lib) for a in arcs])
    for r, a in zip(got, arcs):
        assert_same_as_oracle(r, O.decode(a), "batch walked in parts")


def test_corrupt_frames_walked_in_parts_fail_cleanly(backend, monkeypatch):
    """Bit flips in a long section with the walk in parts: same outcome classes as in one piece (decoded identically, or an
    error; never a hang or a crash), and a truncated chain is reported."""
    lib = library(backend)
    arc = K.fastq_reads(34, 150, with_mask=True)
    L = O.parse(arc)
    s = L.sec[4]
    monkeypatch.setenv("NAFGPU_WALK_SPLIT", "3")
    rng = np.random.default_rng(11)
    for pos in rng.integers(s.offset + 2, s.offset + s.compressed_size, size=16):
        bad = bytearray(arc)
        bad[int(pos)] ^= 1 << int(rng.integers(0, 8))
        try:
            r = decode_soa(lib, bytes(bad))
        except (N.NafError, UnicodeError):
            continue
        try:
            want = O.decode(bytes(bad))
        except O.OracleError:
            continue                      # (libzstd rejects what the device decoded: judged by test_corrupt_input_never_hangs' closed list, in one piece)
        assert r.sequence == want.sequence
    with pytest.raises(N.NafError):
        decode_soa(lib, arc[:s.offset + s.compressed_size // 2])


@pytest.mark.parametrize("block_min", ["1", "1000000"])
def test_both_huffman_kernels(backend, monkeypatch, block_min):
    """Big Huffman streams are decoded by one CTA per block (throughput: batches, chromosomes) or by one CTA per stream in
    clusters of four (latency: a single small archive); the job size picks.  Both on the same archives here."""
    monkeypatch.setenv("NAFGPU_HUF_BLOCK_MIN", block_min)
    check_parity(backend, read_golden("NZ_AAEN01000029.naf"), "fixture, forced kernel")
    check_parity(backend, K.genome(77, 600_000 if backend == "emul" else 3_000_000, level=19), "genome, forced kernel")
    ctx = N.shared_context(0, library(backend))
    p = bytes(np.random.default_rng(3).integers(0, 16, size=300_000).astype(np.uint8))        # fixed-length codes
    frame = K.zstd_frame(p, 3)
    assert ctx.zstd_decompress(frame, len(p)) == p


def test_match_rounds_of_small_jobs_in_both_kernels(backend, monkeypatch):
    """A job of at most 2048 matches resolves them in one CTA (k_lz_small: every other test with a small input); the same
    inputs through the general rounds (k_lz_index, k_lz_first, k_lz_resolve) here."""
    monkeypatch.setenv("NAFGPU_LZ_SMALL", "0")
    for name in ("NZ_AAEN01000029.naf", "phix.naf", "LuxC.naf", "masked.naf"):
        check_parity(backend, read_golden(name), name + ", general rounds")
    check_parity(backend, K.genome(5, 400_000, level=19), "genome, general rounds")
    monkeypatch.delenv("NAFGPU_LZ_SMALL")
    ctx = N.shared_context(0, library(backend))
    monkeypatch.setenv("NAFGPU_LZ_SMALL", "8192")
    check_parity(backend, K.genome(6, 1_500_000 if backend != "emul" else 600_000, level=19), "genome, one CTA, several matches per thread")
    monkeypatch.delenv("NAFGPU_LZ_SMALL")
    res, d = check_parity(backend, K.genome(5, 400_000, level=19), "genome, one CTA")
    assert ctx.stats().lz_rounds >= 1 and ctx.stats().n_sequences <= 2048


def test_in_order_kernel_for_chains_of_some_depth(backend, monkeypatch):
    """Chains of a few dozen generations of matches (a diverged repeat family) leave k_lz_resolve's rounds for k_lz_flow: the
    matches in order, 32 per warp by a ticket, each waiting for the ones its source needs.  Forced here on small inputs
    (NAFGPU_LZ_FLOW=2), with the one-CTA stage off; then with a deadline of zero, which must hand what is left to the finisher."""
    monkeypatch.setenv("NAFGPU_LZ_SMALL", "0")
    monkeypatch.setenv("NAFGPU_LZ_FLOW", "2")
    ctx = N.shared_context(0, library(backend))
    rng = np.random.default_rng(11)
    gens = 400 if backend == "emul" else 6000
    p = _generations(rng, gens, 120, far_every=7, long_every=90)
    for level in (3, 19):
        frame = K.zstd_frame(p, level)
        assert ctx.zstd_decompress(frame, len(p)) == p, level
        # (thousands of generations: on the device the deadline passes and the finisher takes the rest -- the bytes above are
        #  the point; the emulator has no clock and finishes in the kernel)
        assert ctx.stats().lz_flow == 1, (level, ctx.stats().lz_rounds)
    res, d = check_parity(backend, K.cfg3_chromosome(300_000 if backend == "emul" else 4_000_000, workers=0), "repeat family, in-order kernel")
    st = N.shared_context(0, library(backend)).stats()
    assert st.lz_flow == 0 or st.lz_handover == 0, "a few dozen generations: finished in the in-order kernel"
    check_parity(backend, read_golden("NZ_AAEN01000029.naf"), "fixture, in-order kernel allowed")
    if backend != "emul":                                  # (the emulator has no clock: its deadline never passes)
        monkeypatch.setenv("NAFGPU_FIN_COST_US", "0")
        frame = K.zstd_frame(p, 3)
        assert ctx.zstd_decompress(frame, len(p)) == p
        assert ctx.stats().lz_handover > 0, "deadline of zero: the finisher takes over"


@pytest.mark.parametrize("early", ["1", "0"])
def test_in_order_kernel_before_the_rounds(backend, monkeypatch, early):
    """Jobs of a few thousand matches (one genome alone) run k_lz_flow BEFORE the rounds, over every match, with a short deadline;
    what it leaves goes through the rounds, which then count from 2.  With and without it on the same inputs: genomes (finished
    in the kernel), a text-like chain (on the device the deadline passes: rounds, then the finisher)."""
    monkeypatch.setenv("NAFGPU_LZ_FLOW_EARLY", early)
    monkeypatch.setenv("NAFGPU_LZ_SMALL", "0")                     # (also the small inputs below take the general path)
    monkeypatch.setenv("NAFGPU_LZ_FLOW_EARLY", early)
    ctx = N.shared_context(0, library(backend))
    for name in ("NZ_AAEN01000029.naf", "phix.naf", "masked.naf"):
        check_parity(backend, read_golden(name), name + ", early in-order kernel " + early)
    res, d = check_parity(backend, K.genome(9, 600_000 if backend == "emul" else 5_000_000, level=19), "genome, early " + early)
    st = ctx.stats()
    if early == "1" and st.n_sequences > 0: assert st.lz_flow & 2, "a genome's few generations finish before the deadline"
    rng = np.random.default_rng(13)
    p = _generations(rng, 500 if backend == "emul" else 20000, 100, far_every=9)
    frame = K.zstd_frame(p, 3)
    assert ctx.zstd_decompress(frame, len(p)) == p
    if backend != "emul" and early == "1": assert ctx.stats().lz_handover > 0 and not (ctx.stats().lz_flow & 2)


@pytest.mark.parametrize("tiny", ["0", "1"])
def test_tiny_blocks_take_the_warp_per_block_kernels(backend, monkeypatch, tiny):
    """Blocks of at most 32 sequences / 2 KiB of literals (a FASTQ archive in the reference encoder's framing: one flush per
    record) are decoded by one warp each (k_decode_sequences_tiny, k_lz_literals_tiny) once a job has thousands of them; both
    paths on the same archives here, mixed with ordinary blocks, raw and RLE blocks."""
    monkeypatch.setenv("NAFGPU_TINY_BLOCKS", tiny)
    rng = np.random.default_rng(23)
    n = 150 if backend == "emul" else 3000
    motifs = [K.random_dna(rng, 40, b"ACGT") for _ in range(6)]
    seqs, quals = [], []
    for i in range(n):
        parts = [motifs[int(rng.integers(0, 6))] if rng.random() < 0.6 else K.random_dna(rng, int(rng.integers(1, 60)), b"ACGTN") for _ in range(int(rng.integers(1, 6)))]
        s = b"".join(parts)
        if i % 17 == 0: s = b"A" * int(rng.integers(1, 400))                 # RLE blocks
        if i % 29 == 0: s = K.random_dna(rng, 5000, b"ACGT") + s + s        # a block with more than 32 sequences among the tiny ones
        seqs.append(s)
        quals.append(bytes(rng.choice(np.frombuffer(b"FFFF:,#", dtype=np.uint8), size=len(s))))
    ids = [b"r%d" % i for i in range(n)]
    arc = O.encode(ids=ids, sequences=seqs, qualities=quals, level=3, flush_per_record=True)
    res, d = check_parity(backend, arc, "tiny blocks " + tiny)
    assert res.n_lengths == n
    check_parity(backend, read_golden("phix.naf"), "phix " + tiny)
    check_parity(backend, read_golden("NZ_AAEN01000029.naf"), "fixture " + tiny)


def test_sections_longer_than_one_scan_slice(backend):
    """ids, comments and lengths sections of more than 32 KiB are scanned by several CTAs (k_naf_agg + sliced k_naf_scan +
    k_naf_lengths); continuation words and empty strings fall on slice boundaries here."""
    rng = np.random.default_rng(17)
    n = 9000
    lens = rng.integers(0, 12, size=n)
    seqs = [K.random_dna(rng, int(l), b"ACGTN") for l in lens]
    ids = [b"read/%d/%s" % (i, b"x" * int(rng.integers(0, 9))) for i in range(n)]
    coms = [b"" if i % 5 == 0 else b"c%d" % (i * 7919) for i in range(n)]
    arc = O.encode(ids=ids, comments=coms, sequences=seqs, mask_runs_=[3, 5, 40000, 9, int(lens.sum())], level=3, flush_per_record=False)
    L = O.parse(arc)
    assert L.sec[0].original_size > 2 * 32768 and L.sec[2].original_size > 32768
    res, d = check_parity(backend, arc, "sliced scans")
    assert res.n_ids == n and res.n_lengths == n
    check_parity(backend, arc, "sliced scans, no ids", id=False)
