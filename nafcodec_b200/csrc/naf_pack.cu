// naf_pack.cu -- the encode-side counterpart of k_unpack: what the reference computes per record in
//   SequenceWriter::encode / write   nafcodec/src/encoder/writer.rs:31-90   IUPAC -> 4 bit, first residue in the LOW nibble,
//                                                                           odd-length `cache` carried into the next record,
//                                                                           into_inner pads the last nibble (writer.rs:21-28)
//   write_length                     nafcodec/src/encoder/mod.rs:37-44      u32 LE words, 0xFFFFFFFF continues a length
// for all records of an archive at once, plus the soft-mask extraction the reference never wrote (the mask writer is
// commented out, encoder/mod.rs:240, and SequenceWriter rejects lower case): lower-case runs become the Mask section in the
// format MaskReader decodes (decoder/reader.rs:196-231: run = 255 k + b, alternating from an unmasked run).
// zstd COMPRESSION stays on the CPU, as in the reference; these kernels produce the bytes that go into the compressors.
//
//   k_pack          thread = 32 residues: ASCII -> 16 packed bytes (one 16-byte store) + one word of lower-case flags;
//                   records are consecutive slices of the global residue stream, so the odd-length cache is global indexing
//   k_pack_lengths  one CTA: lengths -> words (a length >= 2^32 - 1 takes several), exclusive scan with a running carry
//   k_mask_runs     one CTA streams the flag bitmap: run boundaries are the bit flips; run lengths -> bytes by two CTA-wide
//                   scans per tile (previous boundary: a max-scan; byte offsets: a sum-scan)
// All HBM-bound byte work: algorithmic bytes = residues in + residues / 2 out (+ lengths, mask).
#include "naf_pack.cuh"

namespace nk {

#define FULL 0xFFFFFFFFu

constexpr int PACK_T = 256;
constexpr int SCAN_T = 1024;

struct U2 { uint64_t a, b; };

// Exclusive scan of (sum a, max b) over a SCAN_T-thread CTA; *total gets (sum, max).  All threads must call.
__device__ __forceinline__ U2 block_scan_sum_max(uint64_t a, uint64_t b, U2* total) {
    __shared__ uint64_t wa[33], wb[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint64_t ia = a, ib = b;
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t ta = __shfl_up_sync(FULL, ia, d), tb = __shfl_up_sync(FULL, ib, d);
        if (lane >= d) { ia += ta; ib = tb > ib ? tb : ib; }
    }
    if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
    __syncthreads();
    if (warp == 0) {
        uint64_t xa = lane < nwarps ? wa[lane] : 0, xb = lane < nwarps ? wb[lane] : 0;
        const uint64_t oa = xa;
        uint64_t eb = 0;                                           // exclusive max
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t ta = __shfl_up_sync(FULL, xa, d), tb = __shfl_up_sync(FULL, xb, d);
            if (lane >= d) { xa += ta; xb = tb > xb ? tb : xb; }
        }
        eb = __shfl_up_sync(FULL, xb, 1);
        if (lane == 0) eb = 0;
        wa[lane] = xa - oa; wb[lane] = eb;
        if (lane == 31) { wa[32] = xa; wb[32] = xb; }
    }
    __syncthreads();
    U2 r;
    uint64_t eb = __shfl_up_sync(FULL, ib, 1);                     // exclusive max inside the warp
    if (lane == 0) eb = 0;
    r.a = ia - a + wa[warp];
    r.b = eb > wb[warp] ? eb : wb[warp];
    total->a = wa[32]; total->b = wb[32];
    __syncthreads();
    return r;
}

// SequenceWriter::encode (writer.rs:31-55) as a 256-entry table: low nibble = code; 0x40 = lower-case letter (accepted only
// with extract_mask); 0x80 = "unexpected sequence character".
__device__ __forceinline__ uint8_t pack_lut_entry(uint32_t c, uint32_t seq_type, uint32_t extract_mask) {
    const uint32_t u = (c >= 'a' && c <= 'z') ? c - 32 : c;
    const bool lower = u != c;
    uint32_t code;
    switch (u) {
        case 'A': code = 0x08; break; case 'C': code = 0x04; break; case 'G': code = 0x02; break;
        case 'T': code = seq_type == 0 ? 0x01 : 0x80; break;
        case 'U': code = seq_type == 1 ? 0x01 : 0x80; break;
        case 'R': code = 0x0A; break; case 'Y': code = 0x05; break; case 'S': code = 0x06; break; case 'W': code = 0x09; break;
        case 'K': code = 0x03; break; case 'M': code = 0x0C; break; case 'B': code = 0x07; break; case 'D': code = 0x0B; break;
        case 'H': code = 0x0D; break; case 'V': code = 0x0E; break; case 'N': code = 0x0F; break; case '-': code = 0x00; break;
        default: code = 0x80; break;
    }
    if (lower) code = (extract_mask && !(code & 0x80)) ? (code | 0x40) : 0x80;
    return (uint8_t)code;
}

__global__ void __launch_bounds__(PACK_T) k_pack(const uint8_t* __restrict__ seq, uint64_t n_residues, uint32_t seq_type, uint32_t extract_mask,
                                                 uint8_t* __restrict__ packed, uint32_t* __restrict__ lowbits, unsigned long long* first_invalid) {
    __shared__ uint8_t lut[256];
    for (int i = threadIdx.x; i < 256; i += PACK_T) lut[i] = pack_lut_entry((uint32_t)i, seq_type, extract_mask);
    __syncthreads();
    const uint64_t t = (uint64_t)blockIdx.x * PACK_T + threadIdx.x;
    const uint64_t r0 = t * 32;
    if (r0 >= n_residues) return;                                   // (buffers are padded to whole 32-residue groups)
    const uint4 a = ((const uint4*)(seq + r0))[0], b = ((const uint4*)(seq + r0))[1];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const uint32_t nvalid = n_residues - r0 >= 32 ? 32u : (uint32_t)(n_residues - r0);
    uint32_t out[4] = {0, 0, 0, 0}, low = 0, bad = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        uint32_t e = lut[(w[j >> 2] >> (8 * (j & 3))) & 0xFFu];
        if ((uint32_t)j >= nvalid) e = 0;                           // past the end: the padding nibble is 0 (writer.rs:21-28)
        out[j >> 3] |= (e & 15u) << (4 * (j & 7));                  // residue 2i in the low nibble of byte i
        low |= ((e >> 6) & 1u) << j;
        bad |= (e >> 7) << j;
    }
    ((uint4*)(packed + r0 / 2))[0] = make_uint4(out[0], out[1], out[2], out[3]);
    if (lowbits) lowbits[t] = low;
    if (bad) atomicMin(first_invalid, (unsigned long long)(r0 + (uint64_t)(__ffs((int)bad) - 1)));
}

// write_length (encoder/mod.rs:37-44): while l >= u32::MAX emit FFFFFFFF, l -= u32::MAX; then emit l.
__global__ void __launch_bounds__(SCAN_T) k_pack_lengths(const uint64_t* __restrict__ lengths, uint64_t n_records, uint32_t* __restrict__ words,
                                                         unsigned long long* n_words_out) {
    uint64_t carry = 0;
    for (uint64_t base = 0; base < n_records; base += SCAN_T) {
        const uint64_t r = base + threadIdx.x;
        const uint64_t len = r < n_records ? lengths[r] : 0;
        const uint64_t nw = r < n_records ? len / 0xFFFFFFFFull + 1 : 0;
        U2 tot;
        const U2 ex = block_scan_sum_max(nw, 0, &tot);
        if (r < n_records) {
            uint32_t* o = words + carry + ex.a;
            for (uint64_t k = 0; k + 1 < nw; k++) o[k] = 0xFFFFFFFFu;
            o[nw - 1] = (uint32_t)(len - (nw - 1) * 0xFFFFFFFFull);
        }
        carry += tot.a;
    }
    if (threadIdx.x == 0) *n_words_out = carry;
}

// Run k of the mask ends where the k-th flip of the lower-case flag is (the flag before residue 0 counts as upper case, so a
// sequence that starts lower-case gets the leading zero-length unmasked run the format needs); the last run ends at n.
// A run of length L is L / 255 bytes 0xFF and one byte L % 255 (decoder/reader.rs:196-231).
constexpr int MR_WORDS = 4;                                         // flag words per thread and tile
__device__ __forceinline__ uint64_t put_run(uint8_t* dst, uint64_t len) {
    uint64_t k = 0;
    for (; len >= 255; len -= 255) dst[k++] = 0xFF;
    dst[k++] = (uint8_t)len;
    return k;
}

__global__ void __launch_bounds__(SCAN_T) k_mask_runs(const uint32_t* __restrict__ lowbits, uint64_t n_residues, uint8_t* __restrict__ mask,
                                                      unsigned long long* out_sizes /* [0] bytes, [1] runs */) {
    const uint64_t n_words = (n_residues + 31) / 32;
    uint64_t byte_carry = 0, run_carry = 0, prev_carry = 0;         // bytes / runs written so far; position of the last flip so far
    for (uint64_t base = 0; base < n_words; base += (uint64_t)SCAN_T * MR_WORDS) {
        const uint64_t w0 = base + (uint64_t)threadIdx.x * MR_WORDS;
        uint32_t flips[MR_WORDS];
        uint32_t nflip = 0;
        uint64_t lastpos = 0;                                       // 1 + position of this thread's last flip (0: none)
#pragma unroll
        for (int j = 0; j < MR_WORDS; j++) {
            const uint64_t wi = w0 + j;
            uint32_t cur = 0, prevbit = 0;
            if (wi < n_words) {
                cur = lowbits[wi];
                const uint64_t rem = n_residues - wi * 32;
                if (rem < 32) cur &= (1u << rem) - 1u;              // flags past the end do not exist (they read as "same as the last")
                if (wi > 0) prevbit = lowbits[wi - 1] >> 31;
                if (rem < 32) { const uint32_t lastbit = (cur >> (rem - 1)) & 1u; if (lastbit) cur |= ~((1u << rem) - 1u); }
            }
            flips[j] = wi < n_words ? (cur ^ ((cur << 1) | prevbit)) : 0u;
            nflip += (uint32_t)__popc(flips[j]);
            if (flips[j]) lastpos = wi * 32 + (uint64_t)(31 - __clz((int)flips[j])) + 1;
        }
        U2 tot;
        const U2 ex = block_scan_sum_max(nflip, lastpos, &tot);
        // bytes of this thread's runs: the first one starts at the last flip of everything before this thread
        uint64_t prev = ex.b ? ex.b - 1 : prev_carry;
        uint64_t nbytes = 0;
#pragma unroll
        for (int j = 0; j < MR_WORDS; j++) {
            for (uint32_t m = flips[j]; m; m &= m - 1) {
                const uint64_t p = (w0 + j) * 32 + (uint64_t)(__ffs((int)m) - 1);
                nbytes += (p - prev) / 255 + 1;
                prev = p;
            }
        }
        U2 btot;
        const U2 bex = block_scan_sum_max(nbytes, 0, &btot);
        uint8_t* dst = mask + byte_carry + bex.a;
        prev = ex.b ? ex.b - 1 : prev_carry;
#pragma unroll
        for (int j = 0; j < MR_WORDS; j++) {
            for (uint32_t m = flips[j]; m; m &= m - 1) {
                const uint64_t p = (w0 + j) * 32 + (uint64_t)(__ffs((int)m) - 1);
                dst += put_run(dst, p - prev);
                prev = p;
            }
        }
        byte_carry += btot.a;
        run_carry += tot.a;
        if (tot.b) prev_carry = tot.b - 1;
    }
    if (threadIdx.x == 0) {
        if (n_residues > 0) { byte_carry += put_run(mask + byte_carry, n_residues - prev_carry); run_carry++; }
        out_sizes[0] = byte_carry;
        out_sizes[1] = run_carry;
    }
}

int launch_pack_stage(const uint8_t* seq, uint64_t n_residues, uint32_t seq_type, bool extract_mask, uint8_t* packed, uint32_t* lowbits,
                      const uint64_t* lengths, uint64_t n_records, uint32_t* words, uint8_t* mask, unsigned long long* counters,
                      cudaStream_t st) {
    int launches = 0;
    if (n_residues) {
        const uint64_t groups = (n_residues + 31) / 32;
        NAF_LAUNCH(k_pack, (uint32_t)((groups + PACK_T - 1) / PACK_T), PACK_T, 0, st, seq, n_residues, seq_type, extract_mask ? 1u : 0u, packed,
                   extract_mask ? lowbits : nullptr, counters + 0);
        launches++;
    }
    NAF_LAUNCH(k_pack_lengths, 1, SCAN_T, 0, st, lengths, n_records, words, counters + 1); launches++;
    if (extract_mask) { NAF_LAUNCH(k_mask_runs, 1, SCAN_T, 0, st, lowbits, n_residues, mask, counters + 2); launches++; }
    return launches;
}

}  // namespace nk
