// naf_pack.cuh -- encode-side kernels: IUPAC -> 4-bit packing, length words, soft-mask run extraction.
// See naf_pack.cu for the reference semantics (nafcodec/src/encoder/writer.rs:21-90, encoder/mod.rs:37-44,240).
#pragma once
#include "cuda_compat.h"
#include <stdint.h>

namespace nk {

// Enqueues the pack stage on `stream`.  All pointers are device pointers:
//   seq        n_residues ASCII bytes, 16 B aligned, readable up to the next multiple of 32
//   packed     (n_residues + 1) / 2 bytes, writable up to the next multiple of 16
//   lowbits    one flag word per 32 residues (only with extract_mask)
//   lengths    n_records u64;  words: worst case sum(len / (2^32 - 1) + 1) u32
//   mask       worst case n_residues / 255 + runs bytes (only with extract_mask)
//   counters   [0] first invalid residue (preset to ~0), [1] length words written, [2] mask bytes, [3] mask runs
// Returns the number of kernels launched.
int launch_pack_stage(const uint8_t* seq, uint64_t n_residues, uint32_t seq_type, bool extract_mask, uint8_t* packed, uint32_t* lowbits,
                      const uint64_t* lengths, uint64_t n_records, uint32_t* words, uint8_t* mask, unsigned long long* counters,
                      cudaStream_t stream);

}  // namespace nk
