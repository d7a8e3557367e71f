// naf_text.cuh -- FASTA / FASTQ text straight from the decoded archive on the device.
//
// The reference crate stops at `Record`s and only CARRIES what a formatter needs: Header::line_length and
// Header::name_separator (nafcodec/src/data.rs:198-236, accessors decoder/mod.rs:319-328).  The text itself is what
// upstream `unnaf` prints and what the reference's fixtures hold (data/masked.fna, data/LuxC.faa, data/phix.fastq):
//   FASTA   '>' id [sep comment] '\n'  sequence wrapped at line_length (0 = one line)  '\n'
//   FASTQ   '@' id [sep comment] '\n'  sequence '\n' '+' '\n' quality '\n'
// (the separator and comment are written only when the comment is not empty).
#pragma once
#include "cuda_compat.h"
#include "naf_kernels.cuh"
#include <stdint.h>

namespace nk {

// One per archive of the job.  Offsets are byte offsets into the job's text buffer.
struct TextDev {
    uint64_t text_off;         // where the text of this archive goes (16 B aligned)
    uint64_t cap;              // capacity computed by the host from the section sizes (an upper bound)
    uint64_t offs_off;         // u64[n_records + 1]: text offset of every record (exclusive scan), device only
    uint64_t line_length;      // FASTA wrap; 0 = unwrapped
    uint32_t fastq;            // 1: FASTQ, 0: FASTA
    uint32_t sep;              // name separator byte
};

constexpr uint32_t TEXT_CHUNK = 8192;      // output bytes per CTA of k_text_write
constexpr uint32_t TEXT_THREADS = 128;     // 4 x 16 output bytes per thread: small CTAs, many in flight (the set-up is a chain of dependent loads)

// sizes[a] (u64 at the start of the text buffer) <- text bytes of archive a; then the text.  Returns kernels launched.
int launch_text_stage(uint8_t* arena, const NafDev* archives, uint8_t* text, const TextDev* texts, uint32_t n_archives,
                      uint64_t max_cap, uint32_t* status, cudaStream_t stream);

}  // namespace nk
