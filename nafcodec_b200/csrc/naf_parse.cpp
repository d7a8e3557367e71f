// naf_parse.cpp -- host side of the NAF container: the nom parsers of nafcodec/src/decoder/parser.rs and the
// section table `setup_block!` builds in DecoderBuilder::with_reader (nafcodec/src/decoder/mod.rs:169-242).
// Pure C++ (no CUDA): what stays on the host per the north star.
#include <string.h>

#include "../../include/nafgpu.h"

extern "C" {

// parser::variable_u64 (parser.rs:27-48): big-endian base-128, MSB set = more limbs follow.
int nafgpu_variable_u64(const uint8_t* b, uint64_t len, uint64_t* value) {
    if (!b || !value) return NAFGPU_ERR_ARGUMENT;
    uint64_t k = 0;
    while (k < len && (b[k] & 0x80)) k++;
    if (k >= len) return NAFGPU_ERR_UNEXPECTED_EOF;          // nom::Err::Incomplete
    unsigned __int128 num = 0;
    for (uint64_t j = 0; j <= k; j++) {
        num = (num << 7) | (b[j] & 0x7F);
        if (num > (unsigned __int128)UINT64_MAX) return NAFGPU_ERR_PARSE;   // ErrorKind::TooLarge
    }
    *value = (uint64_t)num;
    return (int)(k + 1);
}

int nafgpu_parse_archive(const uint8_t* bytes, uint64_t len, nafgpu_archive* out) {
    if (!out || (!bytes && len)) return NAFGPU_ERR_ARGUMENT;
    memset(out, 0, sizeof *out);
    uint64_t p = 0;
    // format_descriptor (parser.rs:50-53), format_version (55-62), sequence_type (64-73), flags (75-85),
    // name_separator (87-91), line_length / number_of_sequences (93-99)
    if (len < 3) return NAFGPU_ERR_UNEXPECTED_EOF;            // "failed to read header" (mod.rs:181-186)
    if (!(bytes[0] == 0x01 && bytes[1] == 0xF9 && bytes[2] == 0xEC)) return NAFGPU_ERR_PARSE;
    p = 3;
    if (p >= len) return NAFGPU_ERR_UNEXPECTED_EOF;
    int ver = bytes[p++];
    if (ver != 1 && ver != 2) return NAFGPU_ERR_PARSE;
    nafgpu_header& h = out->header;
    h.format_version = ver;
    h.sequence_type = 0;                                      // v1 implies DNA (parser.rs:104-107)
    if (ver == 2) {
        if (p >= len) return NAFGPU_ERR_UNEXPECTED_EOF;
        int t = bytes[p++];
        if (t > 3) return NAFGPU_ERR_PARSE;
        h.sequence_type = t;
    }
    if (p >= len) return NAFGPU_ERR_UNEXPECTED_EOF;
    h.flags = bytes[p++];
    if (p >= len) return NAFGPU_ERR_UNEXPECTED_EOF;
    if (bytes[p] < 0x20 || bytes[p] > 0x7E) return NAFGPU_ERR_PARSE;
    h.name_separator = bytes[p++];
    int k = nafgpu_variable_u64(bytes + p, len - p, &h.line_length);
    if (k < 0) return k;
    p += k;
    k = nafgpu_variable_u64(bytes + p, len - p, &h.number_of_sequences);
    if (k < 0) return k;
    p += k;
    if (h.flags & 0x40) {                                     // title: parsed and discarded (mod.rs:191-196)
        uint64_t tl;
        k = nafgpu_variable_u64(bytes + p, len - p, &tl);
        if (k < 0) return k;
        p += k;
        if (tl > len - p) return NAFGPU_ERR_UNEXPECTED_EOF;
        p += tl;
    }
    static const uint32_t order[6] = {0x20, 0x10, 0x08, 0x04, 0x02, 0x01};   // Id, Comment, Length, Mask, Sequence, Quality
    for (int s = 0; s < 6; s++) {
        if (!(h.flags & order[s])) continue;
        uint64_t orig, comp;
        k = nafgpu_variable_u64(bytes + p, len - p, &orig);
        if (k < 0) return k;
        p += k;
        k = nafgpu_variable_u64(bytes + p, len - p, &comp);
        if (k < 0) return k;
        p += k;
        if (comp > len - p) return NAFGPU_ERR_UNEXPECTED_EOF;
        out->sections[s].data = bytes + p;
        out->sections[s].compressed_size = comp;
        out->sections[s].original_size = orig;
        out->sections[s].present = 1;
        p += comp;                                            // mod.rs:228
    }
    return NAFGPU_OK;
}

const char* nafgpu_strerror(int status) {
    switch (status) {
        case NAFGPU_OK: return "ok";
        case NAFGPU_ERR_UNEXPECTED_EOF: return "unexpected end of file";
        case NAFGPU_ERR_INVALID_DATA: return "invalid data";
        case NAFGPU_ERR_PARSE: return "failed to parse the NAF header";
        case NAFGPU_ERR_UTF8: return "invalid utf-8";
        case NAFGPU_ERR_CUDA: return "CUDA runtime error";
        case NAFGPU_ERR_NOMEM: return "out of memory";
        case NAFGPU_ERR_ARGUMENT: return "invalid argument";
        case NAFGPU_ERR_NO_DEVICE: return "no CUDA device (this backend has no CPU fallback)";
        case NAFGPU_ERR_UNSUPPORTED: return "unsupported zstd feature";
        default: return "unknown status";
    }
}

const char* nafgpu_version(void) { return "nafgpu 0.1.0 (sm_100a)"; }

}  // extern "C"
