// nafgpu.hpp -- C++17 host-side mirror of the reference decoder surface over the C ABI of nafgpu.h (header-only).
//
// The reference is a Rust crate; this image has no Rust toolchain, so the host side above the C ABI is written in C++
// with the reference's names, argument meaning and error behaviour (the Rust sources of the same layer are kept, unbuilt,
// under rust_shim/).  What mirrors what (reference file:line, relative to the upstream tree):
//
//   nafgpu::SequenceType / FormatVersion / Flag / Flags / Header      nafcodec/src/data.rs:43-236
//   nafgpu::Record                                                    nafcodec/src/data.rs:29-40  (Option<Cow<str>> -> std::optional<std::string>)
//   nafgpu::DecoderBuilder {id,comment,sequence,quality,mask,buffer_size,from_flags,with_bytes,with_path,with_reader}
//                                                                     nafcodec/src/decoder/mod.rs:53-257
//   nafgpu::Decoder {from_path,from_reader(new),header,sequence_type,next,len,into_inner} + range-for iteration
//                                                                     nafcodec/src/decoder/mod.rs:298-461
//   nafgpu::Error (kind Io / Nom / Utf8 / ...)                        nafcodec/src/error.rs:4-11
//
// Rust's `Iterator<Item = Result<Record, Error>>` becomes `std::optional<Record> next()` that throws nafgpu::Error where
// the reference yields `Some(Err(e))`: at the record whose text is not valid UTF-8 (reader.rs:108-109), at the first
// record for corrupt compressed data.  The six per-section streaming zstd readers are replaced by ONE device decode of
// the whole archive on first access (nafgpu_decode); records are sliced out of the structure-of-arrays result.
// With DecoderBuilder::buffer_size(n) the decoded archive stays in device memory and the records cross PCIe in windows of about n
// decoded bytes (nafgpu_job_fetch_window): host memory bounded like the reference's BufReaders (mod.rs:69,104-112).  Contexts
// are pooled per device and outlive the Decoders that used them.
// Link with -lnafgpu (nafcodec_b200/csrc/libnafgpu.so).  There is no CPU fallback.
#ifndef NAFGPU_HPP
#define NAFGPU_HPP

#include <cstdint>
#include <cstring>
#include <fstream>
#include <istream>
#include <mutex>
#include <iterator>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "nafgpu.h"

namespace nafgpu {

enum class SequenceType { Dna = 0, Rna = 1, Protein = 2, Text = 3 };          // data.rs:56-62
enum class FormatVersion { V1 = 1, V2 = 2 };                                  // data.rs:46-50
enum class Flag : uint8_t {                                                   // data.rs:80-97
    Quality = 0x01, Sequence = 0x02, Mask = 0x04, Length = 0x08, Comment = 0x10, Id = 0x20, Title = 0x40, Extended = 0x80
};

class Flags {                                                                 // data.rs:118-196
public:
    Flags() = default;
    explicit Flags(uint8_t v) : v_(v) {}
    bool test(Flag f) const { return (v_ & static_cast<uint8_t>(f)) != 0; }
    void set(Flag f) { v_ |= static_cast<uint8_t>(f); }
    void unset(Flag f) { v_ &= static_cast<uint8_t>(~static_cast<uint8_t>(f)); }
    uint8_t as_byte() const { return v_; }
private:
    uint8_t v_ = 0;
};

class Header {                                                                // data.rs:198-236
public:
    Header() = default;
    explicit Header(const nafgpu_header& h) : h_(h) {}
    FormatVersion format_version() const { return static_cast<FormatVersion>(h_.format_version); }
    SequenceType sequence_type() const { return static_cast<SequenceType>(h_.sequence_type); }
    Flags flags() const { return Flags(static_cast<uint8_t>(h_.flags)); }
    char name_separator() const { return static_cast<char>(h_.name_separator); }
    uint64_t line_length() const { return h_.line_length; }
    uint64_t number_of_sequences() const { return h_.number_of_sequences; }
private:
    nafgpu_header h_{};
};

struct Record {                                                               // data.rs:29-40
    std::optional<std::string> id, comment, sequence, quality;
    std::optional<uint64_t> length;
};

class Error : public std::runtime_error {                                     // error.rs:4-11
public:
    enum class Kind { Io, Nom, Utf8, Device, Argument };
    Error(Kind k, int status, const std::string& what) : std::runtime_error(what), kind_(k), status_(status) {}
    Kind kind() const { return kind_; }
    int status() const { return status_; }                                    // nafgpu_status
private:
    Kind kind_;
    int status_;
};

namespace detail {
inline void check(int rc, const nafgpu_ctx* ctx, const char* what) {
    if (rc == NAFGPU_OK) return;
    std::string msg = nafgpu_strerror(rc);
    if (ctx) { const char* d = nafgpu_last_error(ctx); if (d && *d) { msg += ": "; msg += d; } }
    if (what && *what) { msg += " ["; msg += what; msg += "]"; }
    Error::Kind k = Error::Kind::Device;
    switch (rc) {
        case NAFGPU_ERR_UNEXPECTED_EOF: case NAFGPU_ERR_INVALID_DATA: case NAFGPU_ERR_UNSUPPORTED: case NAFGPU_ERR_NOMEM: k = Error::Kind::Io; break;
        case NAFGPU_ERR_PARSE: k = Error::Kind::Nom; break;
        case NAFGPU_ERR_UTF8: k = Error::Kind::Utf8; break;
        case NAFGPU_ERR_ARGUMENT: k = Error::Kind::Argument; break;
        default: break;
    }
    throw Error(k, rc, msg);
}
// Contexts (streams, events, device arenas, pinned result buffers) outlive the Decoders that used them: a Decoder takes an idle
// one of its device and gives it back when it is dropped, so that a program that opens archive after archive, as callers of
// nafcodec::Decoder do, allocates device and pinned memory once, not per archive.  (Never torn down at exit: the CUDA runtime
// may be gone by then.)
class ContextPool {
public:
    static ContextPool& instance() { static ContextPool* p = new ContextPool; return *p; }
    nafgpu_ctx* take(int device) {
        {
            std::lock_guard<std::mutex> g(m_);
            for (size_t i = 0; i < idle_.size(); i++)
                if (idle_[i].first == device) { nafgpu_ctx* c = idle_[i].second; idle_.erase(idle_.begin() + (std::ptrdiff_t)i); return c; }
        }
        nafgpu_ctx* c = nullptr;
        check(nafgpu_ctx_create(device, &c), nullptr, "nafgpu_ctx_create");
        return c;
    }
    void give(int device, nafgpu_ctx* c) {
        if (!c) return;
        {
            std::lock_guard<std::mutex> g(m_);
            if (idle_.size() < kMaxIdle) { idle_.emplace_back(device, c); return; }
        }
        nafgpu_ctx_destroy(c);
    }
private:
    static constexpr size_t kMaxIdle = 4;
    std::mutex m_;
    std::vector<std::pair<int, nafgpu_ctx*>> idle_;
};
}  // namespace detail

class DecoderBuilder;

class Decoder {
public:
    Decoder(const Decoder&) = delete;
    Decoder& operator=(const Decoder&) = delete;
    Decoder(Decoder&& o) noexcept { *this = std::move(o); }
    Decoder& operator=(Decoder&& o) noexcept {
        if (this != &o) {
            detail::ContextPool::instance().give(device_, ctx_);
            bytes_ = std::move(o.bytes_); arc_ = o.arc_; header_ = o.header_; want_ = o.want_; device_ = o.device_;
            ctx_ = o.ctx_; o.ctx_ = nullptr; res_ = o.res_; decoded_ = o.decoded_; n_ = o.n_; window_bytes_ = o.window_bytes_; window_first_ = o.window_first_; ran_ = o.ran_;   // (a moved vector keeps its buffer: arc_'s section pointers stay valid)
        }
        return *this;
    }
    ~Decoder() { detail::ContextPool::instance().give(device_, ctx_); }

    static Decoder from_path(const std::string& path);                        // mod.rs:304-306
    static Decoder from_reader(std::istream& reader);                         // Decoder::new, mod.rs:315-317
    static Decoder from_bytes(const uint8_t* data, size_t len);

    const Header& header() const { return header_; }                          // mod.rs:326-328
    SequenceType sequence_type() const { return header_.sequence_type(); }    // mod.rs:336-338
    std::vector<uint8_t> into_inner() && { return std::move(bytes_); }        // mod.rs:343-350 (the reader; here: the archive bytes)

    // Iterator::next (mod.rs:444-451): nullopt after header.number_of_sequences records; throws nafgpu::Error where the
    // reference yields Some(Err(_)).
    std::optional<Record> next() {
        if (n_ >= header_.number_of_sequences()) return std::nullopt;
        uint64_t i = n_;
        if (window_bytes_) { fetch_window(n_); i = n_ - window_first_; } else decode_once();
        n_++;
        if (res_.record_status != 0 && res_.first_bad_record == i) detail::check(res_.record_status, nullptr, "record text");
        Record r;
        if (res_.ids && i < res_.n_ids) r.id = slice(res_.ids, res_.id_offsets[i], res_.id_offsets[i + 1] - 1);
        if (res_.comments && i < res_.n_comments) r.comment = slice(res_.comments, res_.comment_offsets[i], res_.comment_offsets[i + 1] - 1);
        if (res_.lengths && i < res_.n_lengths) {
            r.length = res_.lengths[i];
            if (res_.sequence) r.sequence = slice(res_.sequence, res_.record_offsets[i], res_.record_offsets[i + 1]);
            if (res_.quality) r.quality = slice(res_.quality, res_.record_offsets[i], res_.record_offsets[i + 1]);
        }
        return r;
    }
    // ExactSizeIterator::len (mod.rs:453-459): records left
    uint64_t len() const { return header_.number_of_sequences() - n_; }

    // range-for support: `for (const nafgpu::Record& r : decoder)`
    class iterator {
    public:
        using iterator_category = std::input_iterator_tag;
        using value_type = Record;
        using difference_type = std::ptrdiff_t;
        using pointer = const Record*;
        using reference = const Record&;
        iterator() = default;
        explicit iterator(Decoder* d) : d_(d) { ++*this; }
        reference operator*() const { return *cur_; }
        pointer operator->() const { return &*cur_; }
        iterator& operator++() { cur_ = d_->next(); if (!cur_) d_ = nullptr; return *this; }
        bool operator==(const iterator& o) const { return d_ == o.d_; }
        bool operator!=(const iterator& o) const { return d_ != o.d_; }
    private:
        Decoder* d_ = nullptr;
        std::optional<Record> cur_;
    };
    iterator begin() { return iterator(this); }
    iterator end() { return iterator(); }

    // every remaining record (`decoder.collect::<Result<Vec<_>, _>>()`)
    std::vector<Record> collect() { std::vector<Record> v; while (auto r = next()) v.push_back(std::move(*r)); return v; }

private:
    friend class DecoderBuilder;
    Decoder() = default;
    static std::string slice(const uint8_t* blob, uint64_t b, uint64_t e) { return std::string(reinterpret_cast<const char*>(blob) + b, e - b); }
    void decode_once() {
        if (decoded_) return;
        if (!ctx_) ctx_ = detail::ContextPool::instance().take(device_);
        detail::check(nafgpu_decode(ctx_, &arc_, want_, &res_), ctx_, "decode");
        decoded_ = true;
    }
    // buffer_size mode: the archive is decoded into device memory once; records cross PCIe in windows of about window_bytes_
    // decoded bytes (nafgpu_job_fetch_window), so host memory is bounded as with the reference's BufReaders (mod.rs:69,105-112).
    void fetch_window(uint64_t i) {
        if (!ctx_) ctx_ = detail::ContextPool::instance().take(device_);
        if (!ran_) {
            detail::check(nafgpu_job_prepare(ctx_, &arc_, 1, want_), ctx_, "decode");
            detail::check(nafgpu_job_run(ctx_), ctx_, "decode");
            ran_ = true;
        }
        if (decoded_ && i >= window_first_ && i < window_first_ + res_.n_records) return;
        detail::check(nafgpu_job_fetch_window(ctx_, 0, i, header_.number_of_sequences() - i, window_bytes_, &res_), ctx_, "decode");
        window_first_ = i;
        decoded_ = true;
    }
    uint64_t window_bytes_ = 0, window_first_ = 0;
    bool ran_ = false;
    std::vector<uint8_t> bytes_;
    nafgpu_archive arc_{};
    Header header_;
    uint32_t want_ = NAFGPU_WANT_ALL;
    int device_ = 0;
    nafgpu_ctx* ctx_ = nullptr;
    nafgpu_result res_{};
    bool decoded_ = false;
    uint64_t n_ = 0;
};

class DecoderBuilder {                                                        // mod.rs:53-257
public:
    DecoderBuilder() = default;                                               // DecoderBuilder::new (mod.rs:66-75): everything on
    static DecoderBuilder from_flags(Flags flags) {                           // mod.rs:93-101 (quirk kept: never clears `id`)
        DecoderBuilder b;
        b.quality_ = flags.test(Flag::Quality); b.sequence_ = flags.test(Flag::Sequence); b.mask_ = flags.test(Flag::Mask);
        b.comment_ = flags.test(Flag::Comment);
        return b;
    }
    DecoderBuilder& buffer_size(size_t n) { buffer_size_ = n; return *this; }  // mod.rs:104-108: records are fetched in windows of about n decoded bytes (bounded host memory)
    DecoderBuilder& id(bool v) { id_ = v; return *this; }                      // mod.rs:117-121
    DecoderBuilder& comment(bool v) { comment_ = v; return *this; }
    DecoderBuilder& sequence(bool v) { sequence_ = v; return *this; }
    DecoderBuilder& quality(bool v) { quality_ = v; return *this; }
    DecoderBuilder& mask(bool v) { mask_ = v; return *this; }                  // mod.rs:144-148
    DecoderBuilder& device(int index) { device_ = index; return *this; }       // extension: which GPU

    Decoder with_bytes(const uint8_t* data, size_t len) const {                // mod.rs:151-156
        Decoder d;
        d.bytes_.assign(data, data + len);
        d.want_ = (id_ ? NAFGPU_WANT_ID : 0u) | (comment_ ? NAFGPU_WANT_COMMENT : 0u) | (sequence_ ? NAFGPU_WANT_SEQUENCE : 0u) |
                  (quality_ ? NAFGPU_WANT_QUALITY : 0u) | (mask_ ? NAFGPU_WANT_MASK : 0u);
        d.device_ = device_;
        d.window_bytes_ = buffer_size_;
        detail::check(nafgpu_parse_archive(d.bytes_.data(), d.bytes_.size(), &d.arc_), nullptr, "header");
        d.header_ = Header(d.arc_.header);
        return d;
    }
    Decoder with_bytes(const std::vector<uint8_t>& v) const { return with_bytes(v.data(), v.size()); }
    Decoder with_reader(std::istream& reader) const {                          // mod.rs:169-257
        std::vector<uint8_t> v((std::istreambuf_iterator<char>(reader)), std::istreambuf_iterator<char>());
        return with_bytes(v.data(), v.size());
    }
    Decoder with_path(const std::string& path) const {                         // mod.rs:159-166
        std::ifstream f(path, std::ios::binary);
        if (!f) throw Error(Error::Kind::Io, NAFGPU_ERR_UNEXPECTED_EOF, "cannot open " + path);
        return with_reader(f);
    }

private:
    bool id_ = true, comment_ = true, sequence_ = true, quality_ = true, mask_ = true;
    size_t buffer_size_ = 0;
    int device_ = 0;
};

inline Decoder Decoder::from_path(const std::string& path) { return DecoderBuilder().with_path(path); }
inline Decoder Decoder::from_reader(std::istream& reader) { return DecoderBuilder().with_reader(reader); }
inline Decoder Decoder::from_bytes(const uint8_t* data, size_t len) { return DecoderBuilder().with_bytes(data, len); }

}  // namespace nafgpu

#endif  // NAFGPU_HPP
