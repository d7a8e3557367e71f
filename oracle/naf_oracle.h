/*
 * naf_oracle.h -- CPU ORACLE for the nafcodec hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  Nothing under nafcodec_b200/ links, imports or dlopens it.
 *
 * It restates, in plain C, what the reference Rust crate computes on the decode path
 * (reference files, relative to the upstream tree):
 *     nafcodec/src/decoder/parser.rs   header / title / variable_u64
 *     nafcodec/src/decoder/mod.rs      setup_block! section table, next_record, mask_sequence
 *     nafcodec/src/decoder/reader.rs   CStringReader, LengthReader, SequenceReader, MaskReader
 * and, for generating parity archives, the encode path
 *     nafcodec/src/encoder/mod.rs      write_variable_length, write_length, push, write
 *     nafcodec/src/encoder/writer.rs   SequenceWriter (IUPAC -> 4 bit, odd-length cache)
 *
 * The zstd arithmetic itself lives in a third-party dependency that is NOT in the reference
 * tree: crate zstd ^0.13.1 -> zstd-safe -> zstd-sys -> libzstd (C).  The oracle reaches the
 * same C library through dlopen("libzstd.so.1") (1.5.5 in this image), in the magicless
 * format the reference selects with include_magicbytes(false) (decoder/mod.rs:221-223,
 * encoder/mod.rs:151-153).
 *
 * Parity status: PINNED against every golden vector the reference's own tests hold for this
 * path (tests/test_oracle_golden.py: nafcodec/tests/decoder/{dna,fastq,protein}.rs,
 * decoder/mod.rs:463-516, parser.rs:141-152, encoder/mod.rs:403-412, tests/encoder.rs,
 * nafcodec-py tests) and the source FASTA/FASTQ texts shipped beside the fixtures.
 * The Rust original cannot be compiled here (no cargo/rustc in the image).
 */
#ifndef NAF_ORACLE_H
#define NAF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* data.rs:80-97 */
enum {
    NAF_FLAG_QUALITY = 0x01, NAF_FLAG_SEQUENCE = 0x02, NAF_FLAG_MASK = 0x04, NAF_FLAG_LENGTH = 0x08,
    NAF_FLAG_COMMENT = 0x10, NAF_FLAG_ID = 0x20, NAF_FLAG_TITLE = 0x40, NAF_FLAG_EXTENDED = 0x80
};
/* data.rs:56-62 */
enum { NAF_DNA = 0, NAF_RNA = 1, NAF_PROTEIN = 2, NAF_TEXT = 3 };

/* error.rs:4-11 as integer codes (0 = ok). */
enum {
    NAFO_OK = 0,
    NAFO_ERR_IO_EOF = -1,        /* Error::Io(UnexpectedEof)                      */
    NAFO_ERR_IO_INVALID = -2,    /* Error::Io(InvalidData) incl. zstd corruption  */
    NAFO_ERR_NOM = -3,           /* Error::Nom (bad magic / version / separator)  */
    NAFO_ERR_UTF8 = -4,          /* invalid UTF-8 (reader.rs:108-109 -> Io(InvalidData) upstream) */
    NAFO_ERR_INVALID_SEQUENCE = -5,
    NAFO_ERR_INVALID_LENGTH = -6,
    NAFO_ERR_MISSING_FIELD = -7,
    NAFO_ERR_NOMEM = -8,
    NAFO_ERR_ZSTD_MISSING = -9
};

/* Section order is FIXED: Id, Comment, Length, Mask, Sequence, Quality (decoder/mod.rs:237-242). */
enum { NAF_SEC_ID = 0, NAF_SEC_COMMENT = 1, NAF_SEC_LENGTH = 2, NAF_SEC_MASK = 3, NAF_SEC_SEQUENCE = 4, NAF_SEC_QUALITY = 5 };

typedef struct {
    int32_t  format_version;      /* 1 | 2                      parser.rs:55-62  */
    int32_t  sequence_type;       /* NAF_DNA..NAF_TEXT          parser.rs:64-73  */
    uint32_t flags;               /* raw flag byte              parser.rs:75-85  */
    int32_t  name_separator;      /* printable byte             parser.rs:87-91  */
    uint64_t line_length;         /* varint                     parser.rs:93-95  */
    uint64_t number_of_sequences; /* varint                     parser.rs:97-99  */
    uint64_t header_size;         /* bytes consumed by header (+ title)          */
    struct {
        int32_t  present;
        int32_t  _pad;
        uint64_t original_size;   /* Sequence: residues, others: bytes (mod.rs:212) */
        uint64_t compressed_size; /* mod.rs:213 */
        uint64_t offset;          /* file offset of the compressed bytes         */
    } sec[6];
} nafo_layout;

/* Flat, record-indexed result. offsets arrays have n_records+1 entries; a field absent for
 * record i has present[i]==0.  Strings exclude the NUL terminator. */
typedef struct {
    uint64_t n_records;
    uint8_t* ids;        uint64_t* id_off;   uint8_t* id_present;
    uint8_t* comments;   uint64_t* com_off;  uint8_t* com_present;
    uint8_t* sequence;   uint64_t* seq_off;  uint8_t* seq_present;
    uint8_t* quality;    uint64_t* qual_off; uint8_t* qual_present;
    uint64_t* lengths;   uint8_t* len_present;
} nafo_records;

const char* nafo_zstd_version(void);
const char* nafo_last_error(void);

/* parser.rs:27-48 ; returns bytes consumed (>0) or <0 */
int nafo_variable_u64(const uint8_t* buf, size_t len, uint64_t* out);
/* encoder/mod.rs:22-35 ; returns bytes written */
int nafo_write_variable_length(uint64_t n, uint8_t* out);

int nafo_parse(const uint8_t* buf, size_t len, nafo_layout* out);
int nafo_parse_header(const uint8_t* buf, size_t len, nafo_layout* out);   /* header only */

/* want_* mirror DecoderBuilder::{id,comment,sequence,quality,mask} (mod.rs:117-148). */
int nafo_decode(const uint8_t* buf, size_t len, int want_id, int want_comment, int want_sequence,
                int want_quality, int want_mask, nafo_records* out);
void nafo_free_records(nafo_records* r);

/* Decompress one magicless zstd frame with libzstd (the L0 layer of the reference). */
int nafo_zstd_decompress(const uint8_t* src, size_t src_len, uint8_t** dst, size_t* dst_len);
void nafo_free(void* p);
/* Test generator: one magicless frame, optional content checksum / explicit window log (never written by NAF tools). */
int nafo_zstd_compress(const uint8_t* src, size_t src_len, int level, int checksum, int window_log, uint8_t** dst, size_t* dst_len);

/* MaskReader (reader.rs:196-231): decode run lengths from mask bytes. Returns number of runs
 * written to runs[] (cap entries), stopping once sum >= total like the reference. */
int64_t nafo_mask_runs(const uint8_t* mask, size_t mask_len, uint64_t total, uint64_t* runs, size_t cap);

/* Encoder restatement (encoder/mod.rs:163-384, writer.rs) + a Mask section writer in the
 * format MaskReader decodes (the reference encoder cannot write masks, encoder/mod.rs:240).
 * Records arrive as concatenated blobs + offsets (n+1). A NULL blob means the field is not
 * encoded (flag cleared).  mask_runs (may be NULL): alternating Unmasked,Masked,... run lengths.
 * flush_per_record=1 reproduces the reference's zstd call pattern (flush after every record for
 * com/seq/qual: encoder/mod.rs:271,298,319); 0 writes each stream in one piece (ennaf-like). */
int nafo_encode(int sequence_type, int level, int flush_per_record, uint64_t line_length, int name_separator,
                uint64_t n_records,
                const uint8_t* ids, const uint64_t* id_off,
                const uint8_t* comments, const uint64_t* com_off,
                const uint8_t* sequence, const uint64_t* seq_off,
                const uint8_t* quality, const uint64_t* qual_off,
                const uint64_t* mask_runs, uint64_t n_mask_runs,
                uint8_t** out, size_t* out_len);

/* Deterministic synthetic generators (splitmix64 / xoshiro256**), SURVEY 8(d). */
/* DNA with repeats, IUPAC codes, N stretches; writes n uppercase residues into dst. */
void nafo_synth_dna(uint64_t seed, uint64_t n, double gc, int n_repeat_families, uint64_t repeat_len,
                    int repeat_copies, double iupac_rate, uint64_t n_gap_count, uint64_t n_gap_len,
                    uint64_t telomere_len, uint8_t* dst);
void nafo_synth_chromosome(uint64_t seed, uint64_t n, uint8_t* dst);                      /* cfg3 shape */
void nafo_synth_fastq(uint64_t seed, uint64_t n_reads, uint64_t read_len, uint8_t* ids, uint64_t* ids_off, uint8_t* seq, uint8_t* qual);   /* cfg4 shape */
void nafo_set_encoder_workers(int n);    /* generator-only: ZSTD_c_nbWorkers for the encoders created afterwards (0 = reference pattern) */
/* Alternating U/M run lengths with geometric lengths; returns number of runs (sum == total). */
uint64_t nafo_synth_mask(uint64_t seed, uint64_t total, double mean_unmasked, double mean_masked,
                         int leading_zero_run, uint64_t* runs, uint64_t cap);
/* Apply runs to an uppercase sequence (lowercase masked runs) -- to build expected texts. */
void nafo_apply_mask(uint8_t* seq, uint64_t n, const uint64_t* runs, uint64_t n_runs);

/* Timed CPU baseline: full decode (all wanted fields), returns seconds for `iters` decodes. */
double nafo_time_decode(const uint8_t* buf, size_t len, int want_quality, int want_mask, int iters,
                        uint64_t* ascii_bytes_out);

/* The same for a collection on `threads` native threads, one archive per core at a time; seconds for the whole list. */
double nafo_time_decode_many(const uint8_t* const* bufs, const size_t* lens, size_t n, int threads, int want_quality, int want_mask,
                             uint64_t* ascii_bytes_out);

#ifdef __cplusplus
}
#endif
#endif
