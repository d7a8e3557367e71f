/*
 * naf_oracle.c -- CPU ORACLE (test infrastructure, see naf_oracle.h for the rules).
 *
 * Every function names the reference file:line it follows.  This is a restatement of the
 * reference's *semantics* in C, not a translation of its structure: sections are inflated
 * whole (the reference streams them through BufReader<zstd::Decoder<..>>, decoder/mod.rs:32)
 * and the readers then walk the inflated bytes exactly as reader.rs does.
 */
#define _GNU_SOURCE
#include "naf_oracle.h"

#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------ */
/* libzstd through dlopen (no headers in the image): the L0 layer of the reference.            */

typedef struct { const void* src; size_t size; size_t pos; } zin_t;
typedef struct { void* dst; size_t size; size_t pos; } zout_t;

static struct {
    void* h;
    void* (*createDCtx)(void);
    size_t (*freeDCtx)(void*);
    size_t (*dctxSetParameter)(void*, int, int);
    size_t (*decompressStream)(void*, zout_t*, zin_t*);
    void* (*createCCtx)(void);
    size_t (*freeCCtx)(void*);
    size_t (*cctxSetParameter)(void*, int, int);
    size_t (*compressStream)(void*, zout_t*, zin_t*);
    size_t (*flushStream)(void*, zout_t*);
    size_t (*endStream)(void*, zout_t*);
    unsigned (*isError)(size_t);
    const char* (*getErrorName)(size_t);
    const char* (*versionString)(void);
} Z;

/* experimental parameter ids of libzstd 1.5.x (zstd.h): ZSTD_d_format = ZSTD_d_experimentalParam1,
 * ZSTD_c_format = ZSTD_c_experimentalParam2, ZSTD_f_zstd1_magicless = 1. */
enum { ZSTD_d_format = 1000, ZSTD_c_format = 10, ZSTD_c_compressionLevel = 100, ZSTD_f_magicless = 1 };

static __thread char g_err[256];
static void set_err(const char* fmt, const char* arg) { snprintf(g_err, sizeof g_err, fmt, arg ? arg : ""); }
const char* nafo_last_error(void) { return g_err; }

static int zload(void) {
    if (Z.h) return 0;
    void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) { set_err("dlopen libzstd.so.1 failed: %s", dlerror()); return NAFO_ERR_ZSTD_MISSING; }
#define SYM(field, name) do { *(void**)(&Z.field) = dlsym(h, name); if (!Z.field) { set_err("missing symbol %s", name); return NAFO_ERR_ZSTD_MISSING; } } while (0)
    SYM(createDCtx, "ZSTD_createDCtx"); SYM(freeDCtx, "ZSTD_freeDCtx"); SYM(dctxSetParameter, "ZSTD_DCtx_setParameter");
    SYM(decompressStream, "ZSTD_decompressStream"); SYM(createCCtx, "ZSTD_createCCtx"); SYM(freeCCtx, "ZSTD_freeCCtx");
    SYM(cctxSetParameter, "ZSTD_CCtx_setParameter"); SYM(compressStream, "ZSTD_compressStream");
    SYM(flushStream, "ZSTD_flushStream"); SYM(endStream, "ZSTD_endStream"); SYM(isError, "ZSTD_isError");
    SYM(getErrorName, "ZSTD_getErrorName"); SYM(versionString, "ZSTD_versionString");
#undef SYM
    Z.h = h;
    return 0;
}

const char* nafo_zstd_version(void) { return zload() ? "unavailable" : Z.versionString(); }
void nafo_free(void* p) { free(p); }

/* growable byte vector */
typedef struct { uint8_t* p; size_t n, cap; } vec_t;
static int vec_reserve(vec_t* v, size_t extra) {
    if (v->n + extra <= v->cap) return 0;
    size_t c = v->cap ? v->cap : 4096;
    while (c < v->n + extra) c *= 2;
    uint8_t* q = (uint8_t*)realloc(v->p, c);
    if (!q) return NAFO_ERR_NOMEM;
    v->p = q; v->cap = c;
    return 0;
}
static int vec_push(vec_t* v, const void* d, size_t n) {
    if (vec_reserve(v, n)) return NAFO_ERR_NOMEM;
    if (n) memcpy(v->p + v->n, d, n);
    v->n += n;
    return 0;
}

/* zstd::stream::read::Decoder::new + include_magicbytes(false) (decoder/mod.rs:221-222),
 * drained to the end of the frame. */
static int zstd_decompress_hint(const uint8_t* src, size_t src_len, size_t size_hint, uint8_t** dst, size_t* dst_len);
int nafo_zstd_decompress(const uint8_t* src, size_t src_len, uint8_t** dst, size_t* dst_len) {
    return zstd_decompress_hint(src, src_len, 0, dst, dst_len);
}
/* size_hint: the size the container states for the section, so that the output buffer is allocated once (a streaming reader
 * like the reference's never holds a whole section; an oracle that does must not pay for growing it by doubling) */
static int zstd_decompress_hint(const uint8_t* src, size_t src_len, size_t size_hint, uint8_t** dst, size_t* dst_len) {
    int rc = zload();
    if (rc) return rc;
    void* d = Z.createDCtx();
    if (!d) return NAFO_ERR_NOMEM;
    Z.dctxSetParameter(d, ZSTD_d_format, ZSTD_f_magicless);
    vec_t out = {0, 0, 0};
    if (size_hint && size_hint < ((size_t)1 << 36) && vec_reserve(&out, size_hint + (1u << 17))) { Z.freeDCtx(d); return NAFO_ERR_NOMEM; }
    zin_t in = {src, src_len, 0};
    size_t hint = 1;
    rc = 0;
    while (hint != 0) {
        if (vec_reserve(&out, 1u << 17)) { rc = NAFO_ERR_NOMEM; break; }
        zout_t o = {out.p + out.n, out.cap - out.n, 0};
        size_t before = in.pos;
        hint = Z.decompressStream(d, &o, &in);
        if (Z.isError(hint)) { set_err("zstd: %s", Z.getErrorName(hint)); rc = NAFO_ERR_IO_INVALID; break; }
        out.n += o.pos;
        if (hint != 0 && in.pos == in.size && o.pos == 0 && before == in.pos) {
            set_err("zstd: truncated frame%s", NULL); rc = NAFO_ERR_IO_EOF; break;
        }
    }
    Z.freeDCtx(d);
    if (rc) { free(out.p); return rc; }
    if (!out.p) out.p = (uint8_t*)malloc(1);
    *dst = out.p; *dst_len = out.n;
    return 0;
}

/* One magicless frame of `src` in a single ZSTD_endStream sequence, with optional content checksum
 * (ZSTD_c_checksumFlag = 201) and window log (ZSTD_c_windowLog = 101; 0 = the level's default).
 * Test generator only: NAF writers never set these; third-party zstd frames may. */
int nafo_zstd_compress(const uint8_t* src, size_t src_len, int level, int checksum, int window_log, uint8_t** dst, size_t* dst_len) {
    int rc = zload();
    if (rc) return rc;
    void* c = Z.createCCtx();
    if (!c) return NAFO_ERR_NOMEM;
    Z.cctxSetParameter(c, ZSTD_c_compressionLevel, level);
    Z.cctxSetParameter(c, ZSTD_c_format, ZSTD_f_magicless);
    if (checksum) Z.cctxSetParameter(c, 201, 1);
    if (window_log) Z.cctxSetParameter(c, 101, window_log);
    vec_t out = {0, 0, 0};
    zin_t in = {src, src_len, 0};
    rc = 0;
    while (in.pos < in.size) {
        if (vec_reserve(&out, 1u << 17)) { rc = NAFO_ERR_NOMEM; break; }
        zout_t o = {out.p + out.n, out.cap - out.n, 0};
        size_t r = Z.compressStream(c, &o, &in);
        if (Z.isError(r)) { set_err("zstd: %s", Z.getErrorName(r)); rc = NAFO_ERR_IO_INVALID; break; }
        out.n += o.pos;
    }
    size_t r = 1;
    while (!rc && r != 0) {
        if (vec_reserve(&out, 1u << 17)) { rc = NAFO_ERR_NOMEM; break; }
        zout_t o = {out.p + out.n, out.cap - out.n, 0};
        r = Z.endStream(c, &o);
        if (Z.isError(r)) { set_err("zstd: %s", Z.getErrorName(r)); rc = NAFO_ERR_IO_INVALID; break; }
        out.n += o.pos;
    }
    Z.freeCCtx(c);
    if (rc) { free(out.p); return rc; }
    *dst = out.p; *dst_len = out.n;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* parser.rs                                                                                   */

/* parser.rs:27-48 variable_u64: big-endian base-128, MSB = continuation; overflow -> TooLarge. */
int nafo_variable_u64(const uint8_t* buf, size_t len, uint64_t* out) {
    size_t k = 0;
    while (k < len && (buf[k] & 0x80)) k++;
    if (k >= len) return NAFO_ERR_IO_EOF;           /* nom Incomplete */
    uint64_t num = (uint64_t)(buf[k] & 0x7F), basis = 128;
    for (size_t j = k; j-- > 0;) {
        /* the reference multiplies then checked_add()s; a multiplication that wraps or an add that
         * overflows is a TooLarge failure (release builds wrap the multiply; parity is only defined
         * for values that fit). */
        unsigned __int128 term = (unsigned __int128)(buf[j] & 0x7F) * basis;
        if (term > UINT64_MAX || (uint64_t)term > UINT64_MAX - num) return NAFO_ERR_NOM;
        num += (uint64_t)term;
        basis *= 128;
    }
    *out = num;
    return (int)(k + 1);
}

/* encoder/mod.rs:22-35 */
int nafo_write_variable_length(uint64_t n, uint8_t* out) {
    int k = 0;
    unsigned __int128 basis = 1;
    while (basis * 128 <= n) basis *= 128;
    while (basis > 1) {
        out[k++] = (uint8_t)((n / (uint64_t)basis) | 0x80);
        n %= (uint64_t)basis;
        basis /= 128;
    }
    out[k++] = (uint8_t)n;
    return k;
}

/* parser.rs:101-123 header, 125-139 title; decoder/mod.rs:199-242 setup_block! section table. */
static int parse_impl(const uint8_t* buf, size_t len, nafo_layout* L, int header_only) {
    memset(L, 0, sizeof *L);
    size_t p = 0;
#define NEED(n) do { if (p + (n) > len) { set_err("failed to read header%s", NULL); return NAFO_ERR_IO_EOF; } } while (0)
    NEED(3);
    if (!(buf[0] == 0x01 && buf[1] == 0xF9 && buf[2] == 0xEC)) { set_err("bad format descriptor%s", NULL); return NAFO_ERR_NOM; }
    p = 3;
    NEED(1);
    int ver = buf[p++];
    if (ver != 1 && ver != 2) { set_err("invalid format version%s", NULL); return NAFO_ERR_NOM; }
    L->format_version = ver;
    L->sequence_type = NAF_DNA;                       /* parser.rs:104-107: v1 implies DNA */
    if (ver == 2) {
        NEED(1);
        int t = buf[p++];
        if (t > 3) { set_err("invalid sequence type%s", NULL); return NAFO_ERR_NOM; }
        L->sequence_type = t;
    }
    NEED(1); L->flags = buf[p++];
    NEED(1);
    if (buf[p] < 0x20 || buf[p] > 0x7E) { set_err("name separator not printable%s", NULL); return NAFO_ERR_NOM; }
    L->name_separator = buf[p++];
    int k = nafo_variable_u64(buf + p, len - p, &L->line_length);
    if (k < 0) return k;
    p += k;
    k = nafo_variable_u64(buf + p, len - p, &L->number_of_sequences);
    if (k < 0) return k;
    p += k;
    if (L->flags & NAF_FLAG_TITLE) {                  /* decoder/mod.rs:191-196: parsed and discarded */
        uint64_t tl;
        k = nafo_variable_u64(buf + p, len - p, &tl);
        if (k < 0) return k;
        p += k;
        if (tl > len - p) { set_err("title truncated%s", NULL); return NAFO_ERR_IO_EOF; }
        p += tl;
    }
    L->header_size = p;
    if (header_only) return 0;
    static const unsigned order[6] = {NAF_FLAG_ID, NAF_FLAG_COMMENT, NAF_FLAG_LENGTH, NAF_FLAG_MASK, NAF_FLAG_SEQUENCE, NAF_FLAG_QUALITY};
    for (int s = 0; s < 6; s++) {
        if (!(L->flags & order[s])) continue;
        uint64_t orig, comp;
        k = nafo_variable_u64(buf + p, len - p, &orig);
        if (k < 0) return k;
        p += k;
        k = nafo_variable_u64(buf + p, len - p, &comp);
        if (k < 0) return k;
        p += k;
        if (comp > len - p) { set_err("section truncated%s", NULL); return NAFO_ERR_IO_EOF; }
        L->sec[s].present = 1; L->sec[s].original_size = orig; L->sec[s].compressed_size = comp; L->sec[s].offset = p;
        p += comp;                                    /* mod.rs:228 seek past the block */
    }
#undef NEED
    return 0;
}

int nafo_parse(const uint8_t* buf, size_t len, nafo_layout* L) { return parse_impl(buf, len, L, 0); }
/* parser::header alone (parser.rs:101-123), without the section table */
int nafo_parse_header(const uint8_t* buf, size_t len, nafo_layout* L) { return parse_impl(buf, len, L, 1); }

/* ------------------------------------------------------------------------------------------ */
/* reader.rs                                                                                   */

typedef struct { const uint8_t* p; size_t n, pos; } cur_t;

/* CStringReader::next (reader.rs:20-31): read_until(0). Returns 1 = string, 0 = None (EOF),
 * <0 = the reference would panic ("buffer should contain a single nul byte"). */
static int cstring_next(cur_t* c, const uint8_t** s, size_t* l) {
    if (c->pos >= c->n) return 0;
    const uint8_t* z = (const uint8_t*)memchr(c->p + c->pos, 0, c->n - c->pos);
    if (!z) { set_err("string without NUL terminator%s", NULL); return NAFO_ERR_IO_INVALID; }
    *s = c->p + c->pos; *l = (size_t)(z - *s);
    c->pos += *l + 1;
    return 1;
}

/* LengthReader::next (reader.rs:46-68): sum LE u32 words while word == u32::MAX; EOF -> None. */
static int length_next(cur_t* c, uint64_t* out) {
    uint64_t n = 0; uint32_t x = UINT32_MAX;
    while (x == UINT32_MAX) {
        if (c->n - c->pos < 4) { c->pos = c->n; return 0; }   /* read_exact -> UnexpectedEof -> None */
        const uint8_t* b = c->p + c->pos;
        x = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
        c->pos += 4;
        n += x;
    }
    *out = n;
    return 1;
}

/* SequenceReader::decode (reader.rs:151-172) */
static const char IUPAC_DNA[16] = {'-', 'T', 'G', 'K', 'C', 'Y', 'S', 'B', 'A', 'W', 'R', 'D', 'M', 'H', 'V', 'N'};

typedef struct { cur_t c; int ty; int cache; } seqreader_t;   /* cache: -1 = None, else pending high nibble */

/* String::from_utf8 (reader.rs:108-109): strict UTF-8 (no overlongs, no surrogates, <= U+10FFFF). */
static int utf8_valid(const uint8_t* s, size_t n) {
    size_t i = 0;
    while (i < n) {
        uint8_t b = s[i];
        if (b < 0x80) { i++; continue; }
        size_t need; uint32_t cp;
        if (b >= 0xC2 && b <= 0xDF) { need = 1; cp = b & 0x1F; }
        else if (b >= 0xE0 && b <= 0xEF) { need = 2; cp = b & 0x0F; }
        else if (b >= 0xF0 && b <= 0xF4) { need = 3; cp = b & 0x07; }
        else return 0;
        if (n - i <= need) return 0;
        for (size_t k = 1; k <= need; k++) {
            if ((s[i + k] & 0xC0) != 0x80) return 0;
            cp = (cp << 6) | (s[i + k] & 0x3F);
        }
        if (need == 2 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return 0;
        if (need == 3 && (cp < 0x10000 || cp > 0x10FFFF)) return 0;
        i += need + 1;
    }
    return 1;
}

/* SequenceReader::next (reader.rs:88-111), read_nucleotide (121-149), read_text (113-119).
 * The reference spins forever on a truncated stream (no progress when fill_buf is empty); the
 * oracle reports UnexpectedEof instead. */
static int seq_next(seqreader_t* r, uint64_t l, uint8_t* dst) {
    if (r->ty == NAF_DNA || r->ty == NAF_RNA) {
        char t = r->ty == NAF_DNA ? 'T' : 'U';
        uint64_t w = 0;
        if (r->cache >= 0 && l > 0) { dst[w++] = (uint8_t)(r->cache == 1 ? t : IUPAC_DNA[r->cache]); r->cache = -1; }
        while (w < l) {
            uint64_t rem = l - w;
            size_t avail = r->c.n - r->c.pos;
            if (avail == 0) { set_err("sequence stream truncated%s", NULL); return NAFO_ERR_IO_EOF; }
            size_t n = avail < rem / 2 ? avail : (size_t)(rem / 2);
            const uint8_t* b = r->c.p + r->c.pos;
            for (size_t i = 0; i < n; i++) {                      /* HOT LOOP reader.rs:131-136 */
                uint8_t x = b[i];
                uint8_t lo = x & 0x0F, hi = x >> 4;
                dst[w++] = (uint8_t)(lo == 1 ? t : IUPAC_DNA[lo]);
                dst[w++] = (uint8_t)(hi == 1 ? t : IUPAC_DNA[hi]);
            }
            if (n < avail && w == l - 1) {                        /* odd tail reader.rs:138-143 */
                uint8_t x = b[n];
                uint8_t lo = x & 0x0F;
                dst[w++] = (uint8_t)(lo == 1 ? t : IUPAC_DNA[lo]);
                r->cache = x >> 4;
                r->c.pos += n + 1;
            } else {
                r->c.pos += n;
            }
        }
        return 0;
    }
    if (r->c.n - r->c.pos < l) { set_err("text stream truncated%s", NULL); return NAFO_ERR_IO_EOF; }
    memcpy(dst, r->c.p + r->c.pos, l);
    r->c.pos += l;
    if (!utf8_valid(dst, l)) { set_err("invalid utf-8%s", NULL); return NAFO_ERR_UTF8; }   /* reader.rs:108-109 */
    return 0;
}

/* MaskReader (reader.rs:178-231) */
typedef struct { cur_t c; uint64_t total, current; int mask; } maskreader_t;

/* returns 1 = unit (masked flag + n), 0 = None, -1 = the reference would spin (zero-length units at EOF) */
static int mask_next(maskreader_t* m, int* masked, uint64_t* n_out) {
    if (m->current >= m->total) return 0;
    uint64_t n = 0;
    int at_eof = (m->c.pos >= m->c.n);
    while (m->c.pos < m->c.n) {
        uint8_t b = m->c.p[m->c.pos++];
        n += b;                       /* 0xFF adds 255 and continues; any other byte terminates */
        if (b != 0xFF) break;
    }
    if (at_eof) return -1;
    m->current += n;
    *masked = m->mask;
    m->mask = !m->mask;
    *n_out = n;
    return 1;
}

int64_t nafo_mask_runs(const uint8_t* mask, size_t mask_len, uint64_t total, uint64_t* runs, size_t cap) {
    maskreader_t m = {{mask, mask_len, 0}, total, 0, 0};
    int64_t k = 0; int masked; uint64_t n;
    while (mask_next(&m, &masked, &n) == 1) { if ((size_t)k < cap) runs[k] = n; k++; }
    return k;
}

/* Decoder::mask_sequence (decoder/mod.rs:402-441). The unit state is carried across records
 * (self.unit, initial Unmasked(0), mod.rs:254).
 *
 * QUIRK pinned here because parity is against the reference, not against the NAF spec: when a
 * Masked(n) unit reaches or passes the end of the record (n >= seq.len(), mod.rs:413-416) the
 * reference stores the remainder and breaks WITHOUT lower-casing the tail of the record.  Only
 * units that end strictly inside the record lower-case anything.  In global coordinates: residue
 * i inside masked run [a,b) is lower-cased iff b < end_of_record(i). */
typedef struct { int masked; uint64_t n; } maskunit_t;

static int mask_sequence(maskreader_t* mr, maskunit_t* unit, uint8_t* seq, uint64_t len) {
    maskunit_t mask = *unit;
    uint64_t pos = 0;
    for (;;) {
        uint64_t remaining = len - pos;
        if (mask.n < remaining) {
            if (mask.masked)                                           /* make_ascii_lowercase, mod.rs:411 */
                for (uint64_t i = 0; i < mask.n; i++) { uint8_t c = seq[pos + i]; if (c >= 'A' && c <= 'Z') seq[pos + i] = c | 0x20; }
            pos += mask.n;
        } else {
            unit->masked = mask.masked; unit->n = mask.n - remaining;  /* mod.rs:414 / 422 */
            break;
        }
        int masked; uint64_t n;
        int rc = mask_next(mr, &masked, &n);
        if (rc == 0) { set_err("failed to get mask unit%s", NULL); return NAFO_ERR_IO_EOF; }   /* mod.rs:429-434 */
        if (rc < 0) { set_err("mask stream exhausted (reference would not terminate)%s", NULL); return NAFO_ERR_IO_EOF; }
        mask.masked = masked; mask.n = n;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Decoder::next_record / Iterator::next (decoder/mod.rs:356-399, 444-457)                     */

typedef struct { vec_t blob; uint64_t* off; uint8_t* present; } field_t;

static int field_init(field_t* f, uint64_t n) {
    memset(f, 0, sizeof *f);
    f->off = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
    f->present = (uint8_t*)calloc(n ? n : 1, 1);
    return (f->off && f->present) ? 0 : NAFO_ERR_NOMEM;
}

void nafo_free_records(nafo_records* r) {
    if (!r) return;
    free(r->ids); free(r->id_off); free(r->id_present);
    free(r->comments); free(r->com_off); free(r->com_present);
    free(r->sequence); free(r->seq_off); free(r->seq_present);
    free(r->quality); free(r->qual_off); free(r->qual_present);
    free(r->lengths); free(r->len_present);
    memset(r, 0, sizeof *r);
}

int nafo_decode(const uint8_t* buf, size_t len, int want_id, int want_comment, int want_sequence,
                int want_quality, int want_mask, nafo_records* out) {
    memset(out, 0, sizeof *out);
    nafo_layout L;
    int rc = nafo_parse(buf, len, &L);
    if (rc) return rc;
    /* setup_block! (mod.rs:237-242): a section is inflated only if flagged AND wanted; Length always. */
    const int want[6] = {want_id, want_comment, 1, want_mask, want_sequence, want_quality};
    uint8_t* sec[6] = {0}; size_t sec_len[6] = {0}; int have[6] = {0};
    for (int s = 0; s < 6; s++) {
        if (!L.sec[s].present || !want[s]) continue;
        {
            uint64_t hint = L.sec[s].original_size;
            if (s == NAF_SEC_SEQUENCE && (L.sequence_type == NAF_DNA || L.sequence_type == NAF_RNA)) hint = hint / 2 + 1;
            rc = zstd_decompress_hint(buf + L.sec[s].offset, L.sec[s].compressed_size, (size_t)hint, &sec[s], &sec_len[s]);
        }
        if (rc) goto done;
        have[s] = 1;
    }
    uint64_t n = L.number_of_sequences;
    uint64_t seqlen = L.sec[NAF_SEC_SEQUENCE].present ? L.sec[NAF_SEC_SEQUENCE].original_size : 0;   /* mod.rs:236,241 */
    field_t fid, fcom, fseq, fqual;
    if (field_init(&fid, n) || field_init(&fcom, n) || field_init(&fseq, n) || field_init(&fqual, n)) { rc = NAFO_ERR_NOMEM; goto done; }
    out->lengths = (uint64_t*)calloc(n ? n : 1, sizeof(uint64_t));
    out->len_present = (uint8_t*)calloc(n ? n : 1, 1);
    /* one allocation per output blob, sized from the container (bounded: a lying header must not drive the allocation) */
    if (have[NAF_SEC_ID]) vec_reserve(&fid.blob, sec_len[NAF_SEC_ID] + 1);
    if (have[NAF_SEC_COMMENT]) vec_reserve(&fcom.blob, sec_len[NAF_SEC_COMMENT] + 1);
    if (have[NAF_SEC_SEQUENCE] && seqlen < ((uint64_t)1 << 36)) vec_reserve(&fseq.blob, (size_t)seqlen + 1);
    if (have[NAF_SEC_QUALITY]) vec_reserve(&fqual.blob, sec_len[NAF_SEC_QUALITY] + 1);

    cur_t ids = {sec[NAF_SEC_ID], sec_len[NAF_SEC_ID], 0};
    cur_t com = {sec[NAF_SEC_COMMENT], sec_len[NAF_SEC_COMMENT], 0};
    cur_t lens = {sec[NAF_SEC_LENGTH], sec_len[NAF_SEC_LENGTH], 0};
    seqreader_t seq = {{sec[NAF_SEC_SEQUENCE], sec_len[NAF_SEC_SEQUENCE], 0}, L.sequence_type, -1};
    seqreader_t qual = {{sec[NAF_SEC_QUALITY], sec_len[NAF_SEC_QUALITY], 0}, NAF_TEXT, -1};      /* mod.rs:249 */
    maskreader_t mask = {{sec[NAF_SEC_MASK], sec_len[NAF_SEC_MASK], 0}, seqlen, 0, 0};           /* mod.rs:250 */
    maskunit_t unit = {0, 0};                                                                    /* mod.rs:254 */

    for (uint64_t i = 0; i < n; i++) {
        const uint8_t* s; size_t l;
        fid.off[i] = fid.blob.n; fcom.off[i] = fcom.blob.n; fseq.off[i] = fseq.blob.n; fqual.off[i] = fqual.blob.n;
        if (have[NAF_SEC_ID]) {
            int r = cstring_next(&ids, &s, &l);
            if (r < 0) { rc = r; goto done2; }
            if (r) {
                if (!utf8_valid(s, l)) { set_err("id is not utf-8 (reference panics)%s", NULL); rc = NAFO_ERR_UTF8; goto done2; }
                fid.present[i] = 1; if (vec_push(&fid.blob, s, l)) { rc = NAFO_ERR_NOMEM; goto done2; }
            }
        }
        if (have[NAF_SEC_COMMENT]) {
            int r = cstring_next(&com, &s, &l);
            if (r < 0) { rc = r; goto done2; }
            if (r) {
                if (!utf8_valid(s, l)) { set_err("comment is not utf-8 (reference panics)%s", NULL); rc = NAFO_ERR_UTF8; goto done2; }
                fcom.present[i] = 1; if (vec_push(&fcom.blob, s, l)) { rc = NAFO_ERR_NOMEM; goto done2; }
            }
        }
        uint64_t rl = 0; int have_len = 0;
        if (have[NAF_SEC_LENGTH]) have_len = length_next(&lens, &rl);
        if (have_len) {
            out->lengths[i] = rl; out->len_present[i] = 1;
            if (have[NAF_SEC_SEQUENCE]) {
                if (vec_reserve(&fseq.blob, rl + 1)) { rc = NAFO_ERR_NOMEM; goto done2; }
                rc = seq_next(&seq, rl, fseq.blob.p + fseq.blob.n);
                if (rc) goto done2;
                fseq.present[i] = 1;
            }
            if (have[NAF_SEC_QUALITY]) {
                if (vec_reserve(&fqual.blob, rl + 1)) { rc = NAFO_ERR_NOMEM; goto done2; }
                rc = seq_next(&qual, rl, fqual.blob.p + fqual.blob.n);
                if (rc) goto done2;
                fqual.present[i] = 1; fqual.blob.n += rl;
            }
            if (fseq.present[i]) {
                if (have[NAF_SEC_MASK]) {                                   /* mod.rs:386-388, 406 */
                    rc = mask_sequence(&mask, &unit, fseq.blob.p + fseq.blob.n, rl);
                    if (rc) goto done2;
                }
                fseq.blob.n += rl;
            }
        }
    }
    fid.off[n] = fid.blob.n; fcom.off[n] = fcom.blob.n; fseq.off[n] = fseq.blob.n; fqual.off[n] = fqual.blob.n;
done2:
    out->n_records = n;
    out->ids = fid.blob.p; out->id_off = fid.off; out->id_present = fid.present;
    out->comments = fcom.blob.p; out->com_off = fcom.off; out->com_present = fcom.present;
    out->sequence = fseq.blob.p; out->seq_off = fseq.off; out->seq_present = fseq.present;
    out->quality = fqual.blob.p; out->qual_off = fqual.off; out->qual_present = fqual.present;
    if (rc) nafo_free_records(out);
done:
    for (int s = 0; s < 6; s++) free(sec[s]);
    return rc;
}

double nafo_time_decode(const uint8_t* buf, size_t len, int want_quality, int want_mask, int iters, uint64_t* ascii_bytes_out) {
    struct timespec t0, t1;
    uint64_t bytes = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < iters; i++) {
        nafo_records r;
        if (nafo_decode(buf, len, 1, 1, 1, want_quality, want_mask, &r)) return -1.0;
        bytes = r.seq_off ? r.seq_off[r.n_records] : 0;
        nafo_free_records(&r);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (ascii_bytes_out) *ascii_bytes_out = bytes;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* The reference CPU path for a COLLECTION: one archive per core (what running one reference Decoder per thread does; the arc
 * feature makes Decoder Send, lib.rs:24-28).  A native pool: threads pull archive indices from an atomic counter, no
 * interpreter in between.  Large buffers are kept on the threads' malloc arenas instead of being mmap'ed and unmapped for
 * every archive (glibc: M_MMAP_THRESHOLD), which is what made a Python-driven pool lose a third of its per-core rate at
 * 16 threads.  Returns seconds; *ascii_bytes_out = sequence bytes decoded in total. */
#include <malloc.h>
#include <pthread.h>
typedef struct {
    const uint8_t* const* bufs; const size_t* lens; size_t n; int want_quality, want_mask;
    size_t next; uint64_t bytes; int failed; pthread_mutex_t mu;
} pool_t;
static void* pool_worker(void* arg) {
    pool_t* p = (pool_t*)arg;
    uint64_t mine = 0;
    for (;;) {
        size_t i = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
        if (i >= p->n) break;
        nafo_records r;
        if (nafo_decode(p->bufs[i], p->lens[i], 1, 1, 1, p->want_quality, p->want_mask, &r)) { __atomic_store_n(&p->failed, 1, __ATOMIC_RELAXED); break; }
        mine += r.seq_off ? r.seq_off[r.n_records] : 0;
        nafo_free_records(&r);
    }
    __atomic_fetch_add(&p->bytes, mine, __ATOMIC_RELAXED);
    return NULL;
}
double nafo_time_decode_many(const uint8_t* const* bufs, const size_t* lens, size_t n, int threads, int want_quality, int want_mask,
                             uint64_t* ascii_bytes_out) {
    if (zload()) return -1.0;
    if (threads < 1) threads = 1;
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    pool_t p; memset(&p, 0, sizeof p);
    p.bufs = bufs; p.lens = lens; p.n = n; p.want_quality = want_quality; p.want_mask = want_mask;
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, pool_worker, &p);
    pool_worker(&p);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    if (p.failed) return -1.0;
    if (ascii_bytes_out) *ascii_bytes_out = p.bytes;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------------ */
/* Encoder restatement (generator of parity archives)                                          */

typedef struct { void* c; vec_t out; int err; } zenc_t;

/* Generator-only knob (never used by the reference): ZSTD_c_nbWorkers = 400 lets libzstd compress the jobs of ONE frame
 * on several threads, so that the 250 Mbp level-19 bench archive is generated in seconds instead of minutes.  The frame
 * is an ordinary single zstd frame either way; 0 (default) is the reference's call pattern. */
static int g_encoder_workers = 0;
void nafo_set_encoder_workers(int n) { g_encoder_workers = n < 0 ? 0 : n; }

/* EncoderBuilder::new_buffer (encoder/mod.rs:147-154): zstd::Encoder::new(.., level) + include_magicbytes(false) */
static int zenc_init(zenc_t* e, int level) {
    memset(e, 0, sizeof *e);
    e->c = Z.createCCtx();
    if (!e->c) return NAFO_ERR_NOMEM;
    Z.cctxSetParameter(e->c, ZSTD_c_compressionLevel, level);
    Z.cctxSetParameter(e->c, ZSTD_c_format, ZSTD_f_magicless);
    if (g_encoder_workers > 0) Z.cctxSetParameter(e->c, 400, g_encoder_workers);      /* (an error -- no MT support -- is ignored) */
    return 0;
}
/* Write::write_all -> ZSTD_compressStream until the input is consumed */
static void zenc_write(zenc_t* e, const void* d, size_t n) {
    zin_t in = {d, n, 0};
    while (in.pos < in.size && !e->err) {
        if (vec_reserve(&e->out, 1u << 17)) { e->err = NAFO_ERR_NOMEM; return; }
        zout_t o = {e->out.p + e->out.n, e->out.cap - e->out.n, 0};
        size_t r = Z.compressStream(e->c, &o, &in);
        if (Z.isError(r)) { e->err = NAFO_ERR_IO_INVALID; return; }
        e->out.n += o.pos;
    }
}
/* Write::flush -> ZSTD_flushStream until it returns 0 (closes the current block) */
static void zenc_flush(zenc_t* e) {
    size_t r = 1;
    while (r != 0 && !e->err) {
        if (vec_reserve(&e->out, 1u << 17)) { e->err = NAFO_ERR_NOMEM; return; }
        zout_t o = {e->out.p + e->out.n, e->out.cap - e->out.n, 0};
        r = Z.flushStream(e->c, &o);
        if (Z.isError(r)) { e->err = NAFO_ERR_IO_INVALID; return; }
        e->out.n += o.pos;
    }
}
/* Encoder::finish -> ZSTD_endStream until 0 */
static void zenc_finish(zenc_t* e) {
    size_t r = 1;
    while (r != 0 && !e->err) {
        if (vec_reserve(&e->out, 1u << 17)) { e->err = NAFO_ERR_NOMEM; return; }
        zout_t o = {e->out.p + e->out.n, e->out.cap - e->out.n, 0};
        r = Z.endStream(e->c, &o);
        if (Z.isError(r)) { e->err = NAFO_ERR_IO_INVALID; return; }
        e->out.n += o.pos;
    }
    Z.freeCCtx(e->c); e->c = NULL;
}

/* SequenceWriter::encode (encoder/writer.rs:31-55) */
static int iupac_encode(uint8_t c, int ty) {
    switch (c) {
        case 'A': return 0x08; case 'C': return 0x04; case 'G': return 0x02;
        case 'T': return ty == NAF_DNA ? 0x01 : -1;
        case 'U': return ty == NAF_RNA ? 0x01 : -1;
        case 'R': return 0x0A; case 'Y': return 0x05; case 'S': return 0x06; case 'W': return 0x09;
        case 'K': return 0x03; case 'M': return 0x0C; case 'B': return 0x07; case 'D': return 0x0B;
        case 'H': return 0x0D; case 'V': return 0x0E; case 'N': return 0x0F; case '-': return 0x00;
        default: return -1;
    }
}

/* write_block! (encoder/mod.rs:357-374): varint(orig) varint(comp) bytes */
static int put_section(vec_t* file, uint64_t orig, zenc_t* e) {
    uint8_t tmp[24];
    int k = nafo_write_variable_length(orig, tmp);
    k += nafo_write_variable_length(e->out.n, tmp + k);
    if (vec_push(file, tmp, (size_t)k)) return NAFO_ERR_NOMEM;
    return vec_push(file, e->out.p, e->out.n);
}

int nafo_encode(int sequence_type, int level, int flush_per_record, uint64_t line_length, int name_separator,
                uint64_t n,
                const uint8_t* ids, const uint64_t* id_off,
                const uint8_t* comments, const uint64_t* com_off,
                const uint8_t* sequence, const uint64_t* seq_off,
                const uint8_t* quality, const uint64_t* qual_off,
                const uint64_t* mask_runs, uint64_t n_mask_runs,
                uint8_t** out, size_t* out_len) {
    int rc = zload();
    if (rc) return rc;
    int nucl = (sequence_type == NAF_DNA || sequence_type == NAF_RNA);
    zenc_t e_len, e_id, e_com, e_seq, e_qual, e_mask;
    memset(&e_id, 0, sizeof e_id); memset(&e_com, 0, sizeof e_com); memset(&e_seq, 0, sizeof e_seq);
    memset(&e_qual, 0, sizeof e_qual); memset(&e_mask, 0, sizeof e_mask);
    zenc_init(&e_len, level);                                       /* encoder/mod.rs:188: always created */
    if (ids) zenc_init(&e_id, level);
    if (comments) zenc_init(&e_com, level);
    if (sequence) zenc_init(&e_seq, level);
    if (quality) zenc_init(&e_qual, level);
    uint64_t c_len = 0, c_id = 0, c_com = 0, c_seq = 0, c_qual = 0;  /* WriteCounter (counter.rs:25-34) */
    int cache = -1;                                                  /* SequenceWriter::cache */
    vec_t enc = {0, 0, 0};
    rc = 0;
    for (uint64_t i = 0; i < n && !rc; i++) {
        /* Encoder::push (encoder/mod.rs:250-327). record.length is taken from the sequence/quality. */
        int wrote_len = 0;
        if (ids) { zenc_write(&e_id, ids + id_off[i], id_off[i + 1] - id_off[i]); zenc_write(&e_id, "\0", 1); c_id += id_off[i + 1] - id_off[i] + 1; }
        if (comments) {
            zenc_write(&e_com, comments + com_off[i], com_off[i + 1] - com_off[i]); zenc_write(&e_com, "\0", 1);
            c_com += com_off[i + 1] - com_off[i] + 1;
            if (flush_per_record) zenc_flush(&e_com);
        }
        if (sequence) {
            const uint8_t* s = sequence + seq_off[i]; uint64_t l = seq_off[i + 1] - seq_off[i];
            /* write_length (encoder/mod.rs:37-44) */
            uint64_t ll = l; uint8_t w[4];
            while (ll >= UINT32_MAX) { memset(w, 0xFF, 4); zenc_write(&e_len, w, 4); c_len += 4; ll -= UINT32_MAX; }
            w[0] = (uint8_t)ll; w[1] = (uint8_t)(ll >> 8); w[2] = (uint8_t)(ll >> 16); w[3] = (uint8_t)(ll >> 24);
            zenc_write(&e_len, w, 4); c_len += 4; wrote_len = 1;
            if (l > 0) {
                if (!nucl) { zenc_write(&e_seq, s, l); }
                else {
                    /* SequenceWriter::write (encoder/writer.rs:58-90): first residue in the LOW nibble */
                    enc.n = 0;
                    if (vec_reserve(&enc, l / 2 + 2)) { rc = NAFO_ERR_NOMEM; break; }
                    uint64_t k = 0;
                    if (cache >= 0) {
                        int a = iupac_encode(s[0], sequence_type);
                        if (a < 0) { rc = NAFO_ERR_INVALID_SEQUENCE; break; }
                        enc.p[enc.n++] = (uint8_t)((a << 4) | cache); cache = -1; k = 1;
                    }
                    for (; k + 2 <= l; k += 2) {
                        int a = iupac_encode(s[k], sequence_type), b = iupac_encode(s[k + 1], sequence_type);
                        if (a < 0 || b < 0) { rc = NAFO_ERR_INVALID_SEQUENCE; break; }
                        enc.p[enc.n++] = (uint8_t)((b << 4) | a);
                    }
                    if (rc) break;
                    if (k < l) { int a = iupac_encode(s[k], sequence_type); if (a < 0) { rc = NAFO_ERR_INVALID_SEQUENCE; break; } cache = a; }
                    zenc_write(&e_seq, enc.p, enc.n);
                    if (flush_per_record) zenc_flush(&e_seq);          /* writer.rs:88 */
                }
                c_seq += l;
            }
            if (flush_per_record) zenc_flush(&e_seq);                  /* encoder/mod.rs:298 */
        }
        if (quality) {
            const uint8_t* q = quality + qual_off[i]; uint64_t l = qual_off[i + 1] - qual_off[i];
            if (sequence && l != seq_off[i + 1] - seq_off[i]) { rc = NAFO_ERR_INVALID_LENGTH; break; }
            if (!wrote_len) {
                uint64_t ll = l; uint8_t w[4];
                while (ll >= UINT32_MAX) { memset(w, 0xFF, 4); zenc_write(&e_len, w, 4); c_len += 4; ll -= UINT32_MAX; }
                w[0] = (uint8_t)ll; w[1] = (uint8_t)(ll >> 8); w[2] = (uint8_t)(ll >> 16); w[3] = (uint8_t)(ll >> 24);
                zenc_write(&e_len, w, 4); c_len += 4;
            }
            zenc_write(&e_qual, q, l); c_qual += l;
            if (flush_per_record) zenc_flush(&e_qual);                 /* encoder/mod.rs:319 */
        }
    }
    vec_t file = {0, 0, 0};
    if (!rc) {
        /* Encoder::write (encoder/mod.rs:334-384) */
        uint8_t hdr[32]; int k = 0;
        hdr[k++] = 0x01; hdr[k++] = 0xF9; hdr[k++] = 0xEC;
        unsigned flags = 0;
        if (ids) flags |= NAF_FLAG_ID;
        if (comments) flags |= NAF_FLAG_COMMENT;
        if (sequence) flags |= NAF_FLAG_SEQUENCE | NAF_FLAG_LENGTH;
        if (quality) flags |= NAF_FLAG_QUALITY | NAF_FLAG_LENGTH;
        if (mask_runs && sequence) flags |= NAF_FLAG_MASK;
        if (sequence_type == NAF_DNA) { hdr[k++] = 1; }                /* v1 iff DNA (encoder/mod.rs:167-171) */
        else { hdr[k++] = 2; hdr[k++] = (uint8_t)sequence_type; }
        hdr[k++] = (uint8_t)flags; hdr[k++] = (uint8_t)name_separator;
        k += nafo_write_variable_length(line_length, hdr + k);
        k += nafo_write_variable_length(n, hdr + k);
        vec_push(&file, hdr, (size_t)k);
        if (ids) { zenc_finish(&e_id); put_section(&file, c_id, &e_id); }
        if (comments) { zenc_finish(&e_com); put_section(&file, c_com, &e_com); }
        zenc_finish(&e_len); put_section(&file, c_len, &e_len);        /* ALWAYS written (encoder/mod.rs:378) */
        if (flags & NAF_FLAG_MASK) {
            /* Mask section in the format MaskReader reads (reader.rs:196-231): run = 255*k + b, b < 255 */
            vec_t mb = {0, 0, 0};
            for (uint64_t r = 0; r < n_mask_runs; r++) {
                uint64_t v = mask_runs[r];
                while (v >= 255) { uint8_t ff = 0xFF; vec_push(&mb, &ff, 1); v -= 255; }
                uint8_t b = (uint8_t)v; vec_push(&mb, &b, 1);
            }
            zenc_init(&e_mask, level);
            zenc_write(&e_mask, mb.p, mb.n);
            zenc_finish(&e_mask);
            put_section(&file, mb.n, &e_mask);
            free(mb.p);
        }
        if (sequence) {
            if (cache >= 0) { uint8_t b = (uint8_t)cache; zenc_write(&e_seq, &b, 1); }   /* writer.rs:21-28: pad hi=0 */
            zenc_flush(&e_seq);
            zenc_finish(&e_seq); put_section(&file, c_seq, &e_seq);
        }
        if (quality) { zenc_finish(&e_qual); put_section(&file, c_qual, &e_qual); }
        if (e_len.err || e_id.err || e_com.err || e_seq.err || e_qual.err || e_mask.err) rc = NAFO_ERR_IO_INVALID;
    }
    zenc_t* all[6] = {&e_len, &e_id, &e_com, &e_seq, &e_qual, &e_mask};
    for (int i = 0; i < 6; i++) { if (all[i]->c) Z.freeCCtx(all[i]->c); free(all[i]->out.p); }
    free(enc.p);
    if (rc) { free(file.p); return rc; }
    *out = file.p; *out_len = file.n;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Deterministic synthetic inputs: splitmix64 seeding + xoshiro256**             */

typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix64(uint64_t* x) { uint64_t z = (*x += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static void rng_seed(rng_t* r, uint64_t seed) { for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&seed); }
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t* r) {
    uint64_t* s = r->s; uint64_t res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
}
static inline double rng_unit(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t rng_below(rng_t* r, uint64_t n) { return n ? (uint64_t)(((unsigned __int128)rng_next(r) * n) >> 64) : 0; }

void nafo_synth_dna(uint64_t seed, uint64_t n, double gc, int n_repeat_families, uint64_t repeat_len,
                    int repeat_copies, double iupac_rate, uint64_t n_gap_count, uint64_t n_gap_len,
                    uint64_t telomere_len, uint8_t* dst) {
    rng_t r; rng_seed(&r, seed);
    static const char AMBIG[] = "RYSWKMBDHV";
    /* order-0 bases with a per-100 kb GC drift */
    double g = gc;
    for (uint64_t i = 0; i < n; i++) {
        if (i % 100000 == 0) { g = gc + (rng_unit(&r) - 0.5) * 0.08; if (g < 0.05) g = 0.05; if (g > 0.95) g = 0.95; }
        uint64_t x = rng_next(&r);
        double u = (double)(x >> 11) * (1.0 / 9007199254740992.0);
        uint8_t c = (u < g) ? ((x & 1) ? 'G' : 'C') : ((x & 1) ? 'A' : 'T');
        dst[i] = c;
    }
    /* repeat families: exact / diverged copies => long-range matches inside the zstd window */
    for (int f = 0; f < n_repeat_families && repeat_len > 0 && repeat_len < n; f++) {
        uint8_t* unit = (uint8_t*)malloc(repeat_len);
        for (uint64_t i = 0; i < repeat_len; i++) unit[i] = (uint8_t)"ACGT"[rng_next(&r) & 3];
        double divergence = (f % 2) ? 0.10 : 0.0;
        for (int c = 0; c < repeat_copies; c++) {
            uint64_t at = rng_below(&r, n - repeat_len);
            for (uint64_t i = 0; i < repeat_len; i++) {
                uint8_t b = unit[i];
                if (divergence > 0 && rng_unit(&r) < divergence) b = (uint8_t)"ACGT"[rng_next(&r) & 3];
                dst[at + i] = b;
            }
        }
        free(unit);
    }
    if (iupac_rate > 0) {
        uint64_t k = (uint64_t)((double)n * iupac_rate);
        for (uint64_t i = 0; i < k; i++) dst[rng_below(&r, n)] = (uint8_t)AMBIG[rng_below(&r, 10)];
    }
    for (uint64_t gph = 0; gph < n_gap_count && n_gap_len > 0 && n_gap_len < n; gph++) {
        uint64_t at = rng_below(&r, n - n_gap_len);
        memset(dst + at, 'N', n_gap_len);
    }
    if (telomere_len > 0 && 2 * telomere_len < n) { memset(dst, 'N', telomere_len); memset(dst + n - telomere_len, 'N', telomere_len); }
}

/* cfg3: a human-chromosome-shaped record: GC 0.41, 10 kb N telomeres, one N centromere (1.2 % of the
 * length: 3 Mbp at 250 Mbp), 20 N gaps of 50 kb, a 300 bp repeat family at 10 % divergence with one copy per 2.5 kb. */
void nafo_synth_chromosome(uint64_t seed, uint64_t n, uint8_t* dst) {
    uint64_t gap = n >= 5000000 ? 50000 : n / 100;
    nafo_synth_dna(seed, n, 0.41, 2, n > 3000 ? 300 : 0, (int)(n / 5000), 1e-5, 20, gap, n >= 1000000 ? 10000 : n / 100, dst);
    uint64_t cen = n / 83, at = n * 2 / 5;
    if (cen && at + cen < n) memset(dst + at, 'N', cen);
}

/* cfg4: n reads of exactly read_len bp sampled from a fixed 5386 bp genome with 1 % substitutions and N at
 * 1e-3; quality from a 4-symbol Markov chain over "F:,#"; ids "SRR0000001.<i>".  Fills the three blobs; ids_off gets
 * n + 1 entries (sequence / quality offsets are i * read_len).  ids must hold 24 bytes per read. */
void nafo_synth_fastq(uint64_t seed, uint64_t n_reads, uint64_t read_len, uint8_t* ids, uint64_t* ids_off, uint8_t* seq, uint8_t* qual) {
    rng_t r; rng_seed(&r, seed * 0x9E3779B97F4A7C15ull + 17);
    enum { G = 5386 };
    static const char Q[] = "F:,#";
    uint8_t genome[G];
    for (int i = 0; i < G; i++) genome[i] = (uint8_t)"ACGT"[rng_next(&r) & 3];
    uint64_t io = 0;
    for (uint64_t k = 0; k < n_reads; k++) {
        ids_off[k] = io;
        io += (uint64_t)sprintf((char*)ids + io, "SRR0000001.%llu", (unsigned long long)(k + 1));
        uint64_t p = rng_below(&r, G - read_len);
        uint8_t* s = seq + k * read_len; uint8_t* q = qual + k * read_len;
        int state = (int)(rng_next(&r) & 3);
        for (uint64_t i = 0; i < read_len; i++) {
            uint64_t x = rng_next(&r);
            uint8_t b = genome[p + i];
            if ((x & 0x7F) == 0 && ((x >> 7) & 1)) b = (uint8_t)"ACGT"[(x >> 8) & 3];      /* ~1 % (incl. silent) substitutions */
            if (((x >> 16) & 0x3FF) == 0) b = 'N';                                            /* ~1e-3 */
            s[i] = b;
            if (((x >> 32) & 0xF) == 0) state = (int)((x >> 40) & 3);                         /* quality runs of mean length 16 */
            q[i] = (uint8_t)Q[state];
        }
    }
    ids_off[n_reads] = io;
}

uint64_t nafo_synth_mask(uint64_t seed, uint64_t total, double mean_unmasked, double mean_masked,
                         int leading_zero_run, uint64_t* runs, uint64_t cap) {
    rng_t r; rng_seed(&r, seed ^ 0xA5A5A5A5DEADBEEFull);
    uint64_t k = 0, sum = 0; int masked = 0;
    if (leading_zero_run && cap) { runs[k++] = 0; masked = 1; }      /* sequence starts lower-case: U(0) first */
    /* force the awkward lengths once: exactly 255, > 255, > 65535 (mask bytes FF 00, FF.. xx) */
    const uint64_t forced[4] = {255, 510, 70000, 254};
    int fi = 0;
    while (sum < total && k < cap) {
        double mean = masked ? mean_masked : mean_unmasked;
        uint64_t len = (uint64_t)(-log(1.0 - rng_unit(&r)) * mean) + 1;
        if (fi < 4 && k >= 4 && (k % 7) == 4) len = forced[fi++];
        if (len > total - sum) len = total - sum;
        runs[k++] = len; sum += len; masked = !masked;
    }
    return k;
}

void nafo_apply_mask(uint8_t* seq, uint64_t n, const uint64_t* runs, uint64_t n_runs) {
    uint64_t pos = 0;
    for (uint64_t k = 0; k < n_runs && pos < n; k++) {
        uint64_t len = runs[k]; if (len > n - pos) len = n - pos;
        if (k & 1) for (uint64_t i = 0; i < len; i++) { uint8_t c = seq[pos + i]; if (c >= 'A' && c <= 'Z') seq[pos + i] = c | 0x20; }
        pos += len;
    }
}
