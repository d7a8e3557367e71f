// The reference's own decoder tests, restated against the C++ mirror (include/nafgpu.hpp):
//   nafcodec/tests/decoder/dna.rs      decode, mask, force_nomask
//   nafcodec/tests/decoder/fastq.rs    decode_header, decode, decode_no_id / no_seq / no_comment / no_quality
//   nafcodec/tests/decoder/protein.rs  decode
//   nafcodec/src/decoder/mod.rs:478-504  error on a truncated archive
//   nafcodec/src/decoder/mod.rs:104-112  buffer_size: same records through windows of any size
// Usage: test_decoder <directory with the .naf fixtures>.  Exit code 0 = all passed.
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "nafgpu.hpp"

using nafgpu::Decoder;
using nafgpu::DecoderBuilder;
using nafgpu::Flag;
using nafgpu::Header;
using nafgpu::Record;
using nafgpu::SequenceType;

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "%s:%d: CHECK failed: %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)
#define CHECK_EQ(a, b) do { if (!((a) == (b))) { std::cerr << __FILE__ << ":" << __LINE__ << ": " #a " == " #b " failed (" << (a) << " vs " << (b) << ")\n"; failures++; } } while (0)

static std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); std::exit(2); }
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static size_t count(const std::string& s, char c) { return (size_t)std::count(s.begin(), s.end(), c); }
static bool all_upper(const std::string& s, size_t b, size_t e) { return std::all_of(s.begin() + b, s.begin() + e, [](unsigned char x) { return std::isupper(x); }); }
static bool all_lower(const std::string& s, size_t b, size_t e) { return std::all_of(s.begin() + b, s.begin() + e, [](unsigned char x) { return std::islower(x); }); }
static bool starts_with(const std::string& s, const char* p) { return s.rfind(p, 0) == 0; }

static std::string DIR;

static void dna_decode() {                                       // dna.rs:8-34
    std::ifstream c(DIR + "/NZ_AAEN01000029.naf", std::ios::binary);
    Decoder decoder = Decoder::from_reader(c);
    CHECK_EQ(decoder.header().name_separator(), ' ');
    CHECK_EQ(decoder.header().number_of_sequences(), 30u);
    CHECK_EQ(decoder.header().line_length(), 80u);
    CHECK(decoder.header().sequence_type() == SequenceType::Dna);
    Record r1 = decoder.next().value();
    CHECK_EQ(r1.id.value(), "NZ_AAEN01000029.1");
    CHECK_EQ(r1.comment.value(), "Bacillus anthracis str. CNEVA-9066 map unlocalized plasmid pXO1 cont2250, whole genome shotgun sequence");
    const std::string seq = r1.sequence.value();
    CHECK_EQ(seq.size(), 182777u);
    CHECK_EQ(count(seq, 'A'), 62115u);
    CHECK_EQ(count(seq, 'C'), 28747u);
    CHECK_EQ(count(seq, 'G'), 30763u);
    CHECK_EQ(count(seq, 'T'), 61152u);
    Record r2 = decoder.next().value();
    CHECK_EQ(r2.id.value(), "NZ_AAEN01000030.3");
    CHECK_EQ(r2.comment.value(), "Bacillus anthracis str. CNEVA-9066 map unlocalized plasmid pXO2 cont2251, whole genome shotgun sequence");
    CHECK_EQ(decoder.len(), 28u);
    CHECK_EQ(decoder.collect().size(), 28u);
    CHECK(!decoder.next().has_value());
}

static void dna_mask() {                                         // dna.rs:36-63
    const std::vector<uint8_t> MASKED = read_file(DIR + "/masked.naf");
    Decoder decoder = Decoder::from_bytes(MASKED.data(), MASKED.size());
    CHECK_EQ(decoder.header().name_separator(), ' ');
    CHECK_EQ(decoder.header().number_of_sequences(), 2u);
    CHECK_EQ(decoder.header().line_length(), 50u);
    CHECK(decoder.header().sequence_type() == SequenceType::Dna);
    Record r1 = decoder.next().value();
    CHECK_EQ(r1.id.value(), "test1");
    std::string seq = r1.sequence.value();
    CHECK(all_upper(seq, 0, 657)); CHECK(all_lower(seq, 657, 676)); CHECK(all_upper(seq, 676, 1311)); CHECK(all_lower(seq, 1311, 1350));
    Record r2 = decoder.next().value();
    CHECK_EQ(r2.id.value(), "test2");
    seq = r2.sequence.value();
    CHECK(all_upper(seq, 0, 525)); CHECK(all_lower(seq, 525, 621)); CHECK(all_upper(seq, 621, 720)); CHECK(all_lower(seq, 720, 733));
    CHECK(!decoder.next().has_value());
}

static void dna_force_nomask() {                                 // dna.rs:65-88
    const std::vector<uint8_t> MASKED = read_file(DIR + "/masked.naf");
    Decoder decoder = DecoderBuilder().mask(false).with_bytes(MASKED);
    CHECK_EQ(decoder.header().number_of_sequences(), 2u);
    CHECK_EQ(decoder.header().line_length(), 50u);
    Record r1 = decoder.next().value();
    CHECK_EQ(r1.id.value(), "test1");
    CHECK(all_upper(r1.sequence.value(), 0, r1.sequence->size()));
    Record r2 = decoder.next().value();
    CHECK_EQ(r2.id.value(), "test2");
    CHECK(all_upper(r2.sequence.value(), 0, r2.sequence->size()));
    CHECK(!decoder.next().has_value());
}

static void check_header_flags(const Header& header) {           // fastq.rs:9-14
    CHECK(header.flags().test(Flag::Quality));
    CHECK(header.flags().test(Flag::Sequence));
    CHECK(header.flags().test(Flag::Id));
    CHECK(header.flags().test(Flag::Comment));
}

static void fastq_decode() {                                     // fastq.rs:16-55
    const std::vector<uint8_t> ARCHIVE = read_file(DIR + "/phix.naf");
    Decoder decoder = Decoder::from_bytes(ARCHIVE.data(), ARCHIVE.size());
    CHECK_EQ(decoder.header().name_separator(), ' ');
    CHECK_EQ(decoder.header().number_of_sequences(), 42u);
    CHECK(decoder.header().sequence_type() == SequenceType::Dna);
    check_header_flags(decoder.header());
    Record r1 = decoder.next().value();
    CHECK_EQ(r1.id.value(), "SRR1377138.1");
    CHECK_EQ(r1.comment.value(), "a comment that should not be included in the SAM output");
    CHECK(starts_with(r1.sequence.value(), "NGCTCTTAAACCTGCTATTGAGGCTTGTGGCATTTC"));
    CHECK(starts_with(r1.quality.value(), "#8CCCGGGGGGGGGGGGGGGGGGGGGGGGGG"));
    Record r2 = decoder.next().value();
    CHECK_EQ(r2.id.value(), "SRR1377138.2");
    CHECK_EQ(r2.comment.value(), "some lowercase nucleotides");
    CHECK_EQ(decoder.collect().size(), 40u);
}

static void fastq_skip(const char* field) {                      // fastq.rs:57-118
    const std::vector<uint8_t> ARCHIVE = read_file(DIR + "/phix.naf");
    DecoderBuilder b;
    const std::string f = field;
    if (f == "id") b.id(false); else if (f == "sequence") b.sequence(false); else if (f == "comment") b.comment(false); else b.quality(false);
    Decoder decoder = b.with_bytes(ARCHIVE);
    check_header_flags(decoder.header());
    for (int k = 0; k < 2; k++) {
        Record r = decoder.next().value();
        CHECK_EQ(r.id.has_value(), f != "id");
        CHECK_EQ(r.sequence.has_value(), f != "sequence");
        CHECK_EQ(r.comment.has_value(), f != "comment");
        CHECK_EQ(r.quality.has_value(), f != "quality");
        CHECK(r.length.has_value());
    }
    size_t n = 0;
    for (const Record& r : decoder) { (void)r; n++; }            // range-for over what is left
    CHECK_EQ(n, 40u);
}

static void protein_decode() {                                   // protein.rs:4-22
    Decoder decoder = Decoder::from_path(DIR + "/LuxC.naf");
    CHECK_EQ(decoder.header().name_separator(), ' ');
    CHECK_EQ(decoder.header().number_of_sequences(), 12u);
    CHECK_EQ(decoder.header().line_length(), 60u);
    CHECK(decoder.header().sequence_type() == SequenceType::Protein);
    Record r1 = decoder.next().value();
    CHECK(r1.id.has_value());
    CHECK(r1.sequence.has_value());
    CHECK_EQ(r1.sequence->size(), 488u);
}

static void errors() {                                           // decoder/mod.rs:478-504 and parser errors
    std::vector<uint8_t> a = read_file(DIR + "/masked.naf");
    bool threw = false;
    try { Decoder::from_bytes(a.data(), 3); } catch (const nafgpu::Error& e) { threw = true; CHECK(e.status() != 0); }
    CHECK(threw);
    std::vector<uint8_t> bad = a;
    bad[0] ^= 0xFF;                                              // format descriptor
    threw = false;
    try { Decoder::from_bytes(bad.data(), bad.size()); } catch (const nafgpu::Error& e) { threw = true; CHECK(e.kind() == nafgpu::Error::Kind::Nom); }
    CHECK(threw);
    std::vector<uint8_t> corrupt = read_file(DIR + "/phix.naf");
    for (size_t i = corrupt.size() - 900; i < corrupt.size() - 700; i++) corrupt[i] ^= 0x5A;   // inside a compressed section
    threw = false;
    try { Decoder d = Decoder::from_bytes(corrupt.data(), corrupt.size()); d.collect(); } catch (const nafgpu::Error& e) { threw = true; CHECK(e.kind() == nafgpu::Error::Kind::Io); }
    CHECK(threw);
}

// DecoderBuilder::buffer_size (mod.rs:104-112): the same records whatever the buffer; here it bounds the host side, the records
// crossing PCIe in windows (nafgpu_job_fetch_window).
static void buffer_size_windows() {
    for (const char* name : {"/phix.naf", "/masked.naf", "/LuxC.naf", "/CP040672.naf"}) {
        const std::vector<uint8_t> ARCHIVE = read_file(DIR + name);
        std::vector<Record> whole = Decoder::from_bytes(ARCHIVE.data(), ARCHIVE.size()).collect();
        for (size_t bs : {size_t(1), size_t(4096), size_t(1) << 20}) {
            Decoder d = DecoderBuilder().buffer_size(bs).with_bytes(ARCHIVE);
            CHECK_EQ(d.len(), whole.size());
            std::vector<Record> got = d.collect();
            CHECK_EQ(got.size(), whole.size());
            for (size_t i = 0; i < got.size() && i < whole.size(); i++) {
                CHECK(got[i].id == whole[i].id);
                CHECK(got[i].comment == whole[i].comment);
                CHECK(got[i].sequence == whole[i].sequence);
                CHECK(got[i].quality == whole[i].quality);
                CHECK(got[i].length == whole[i].length);
            }
        }
    }
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s <fixture directory>\n", argv[0]); return 2; }
    DIR = argv[1];
    try {
        dna_decode(); dna_mask(); dna_force_nomask();
        fastq_decode();
        fastq_skip("id"); fastq_skip("sequence"); fastq_skip("comment"); fastq_skip("quality");
        protein_decode();
        errors();
        buffer_size_windows();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "unexpected exception: %s\n", e.what());
        return 1;
    }
    if (failures) { std::fprintf(stderr, "%d check(s) failed\n", failures); return 1; }
    std::puts("all reference decoder tests passed");
    return 0;
}
