// ps_decode_bits.h — the per-chunk part of the FASTA decoder's write pass, free of CUDA-only constructs so
// that tests/ can compile it for the host (tests/test_decode_bits.py) and compare it with the byte-at-a-time
// transducer of ps_decode.cuh on the CPU.
//
// Replaces the reader half of `glistmaker` (modeling.py:303-315) together with ps_decode.cuh.
//
// A thread owns a 64-byte chunk of text (16 words) and turns it into at most 64 codes. Instead of storing
// one code byte per position and packing afterwards, the codes are kept as two bit strings in registers —
// 2 bits per base, first base most significant (the stream's own order), and 1 break bit per position,
// first position least significant — and the finished strings are OR-ed into the tile's packed output at
// the chunk's bit offset. A word whose four bytes are all bases (15 of 16 words of a 60-column FASTA file)
// is classified, packed and appended with word-wide (SWAR) arithmetic: ~11 instructions per byte instead of
// ~48 for the byte loop.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PS_HD __host__ __device__ __forceinline__
#else
#define PS_HD inline
#endif

#define PSD_CODE_BREAK 4u
#define PSD_CODE_SKIP 5u

// byte -> 0..3 (A C G T/U, either case), 4 = window break, 5 = skipped (1..31)
PS_HD uint32_t psd_classify(uint32_t b) {
    if (b - 1u < 31u) return PSD_CODE_SKIP;
    const uint32_t u = b & 0xDFu;
    const uint32_t idx = u - 'A';
    const bool ok = idx < 32u && ((0x00180045u >> idx) & 1u);
    uint32_t c = (u >> 1) & 3u;
    c ^= c >> 1;
    return ok ? c : PSD_CODE_BREAK;
}

struct PsdBits {
    uint64_t sq_hi, sq_lo;   // 2 bits per code, appended at the low end (first code ends up most significant)
    uint64_t bd;             // break bit of code i at bit i
    uint32_t n;              // codes so far (<= 64)
    uint32_t s;              // transducer state: 0 sequence, 1 header
};

PS_HD void psd_append1(PsdBits &o, uint32_t code) {
    o.sq_hi = (o.sq_hi << 2) | (o.sq_lo >> 62);
    o.sq_lo = (o.sq_lo << 2) | (uint64_t)(code & 3u);
    o.bd |= (uint64_t)(code >> 2) << o.n;
    o.n++;
}

// 0x80 in every byte of v that is zero (exact: no carries between bytes)
PS_HD uint32_t psd_zero_bytes(uint32_t v) { return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v | 0x7F7F7F7Fu); }

// FASTA transducer over the bytes [jlo, jhi) of one 64-byte chunk (16 little-endian words), starting in state
// o.s; appends to o.
PS_HD void psd_fasta_chunk(const uint32_t *w, int jlo, int jhi, PsdBits &o) {
    for (int q = 0; q < 16; q++) {
        const uint32_t x = w[q];
        const int j0 = 4 * q;
        bool fast = o.s == 0u && j0 >= jlo && j0 + 4 <= jhi;
        if (fast) {
            const uint32_t u = x & 0xDFDFDFDFu;                       // fold case
            const uint32_t m = psd_zero_bytes(u ^ 0x41414141u) | psd_zero_bytes(u ^ 0x43434343u) |
                               psd_zero_bytes(u ^ 0x47474747u) | psd_zero_bytes(u ^ 0x54545454u) |
                               psd_zero_bytes(u ^ 0x55555555u);
            fast = m == 0x80808080u;                                  // A C G T U in all four bytes
        }
        if (fast) {
            const uint32_t t = (x >> 1) & 0x03030303u;
            const uint32_t c = t ^ ((t >> 1) & 0x01010101u);          // A0 C1 G2 T3 per byte
            const uint32_t c8 = (c * 0x40100401u) >> 24;              // first base in the top two bits
            o.sq_hi = (o.sq_hi << 8) | (o.sq_lo >> 56);
            o.sq_lo = (o.sq_lo << 8) | (uint64_t)c8;
            o.n += 4;
        } else {
            for (int jj = 0; jj < 4; jj++) {
                const int j = j0 + jj;
                if (j < jlo || j >= jhi) continue;
                const uint32_t b = (x >> (8 * jj)) & 0xFFu;
                if (o.s == 1u) { if (b == '\n') o.s = 0u; continue; }
                if (b == '>') { psd_append1(o, PSD_CODE_BREAK); o.s = 1u; continue; }
                const uint32_t c = psd_classify(b);
                if (c != PSD_CODE_SKIP) psd_append1(o, c);
            }
        }
    }
}

// OR the chunk's strings into packed output arrays whose position 0 is `pos0` positions before this chunk's
// first code: seq words hold 16 positions each, first position in the top bits; bad words hold 32 positions
// each, first position in bit 0. or_fn(word pointer, value) performs the (atomic) OR.
template <typename OrFn>
PS_HD void psd_place(const PsdBits &o, uint32_t pos0, uint32_t *seq_words, uint32_t *bad_words, OrFn or_fn) {
    if (o.n == 0) return;
    // left-align the 2n-bit string in 128 bits
    uint64_t hi = o.sq_hi, lo = o.sq_lo;
    const uint32_t up = 128u - 2u * o.n;                               // 0 .. 126
    if (up >= 64u) { hi = lo << (up - 64u); lo = 0; }
    else if (up) { hi = (hi << up) | (lo >> (64u - up)); lo <<= up; }
    const uint32_t S[6] = {0u, (uint32_t)(hi >> 32), (uint32_t)hi, (uint32_t)(lo >> 32), (uint32_t)lo, 0u};
    const uint32_t w0 = pos0 >> 4, sh = 2u * (pos0 & 15u);
    const uint32_t nw = (sh + 2u * o.n + 31u) >> 5;                     // words touched, <= 5
    for (uint32_t j = 0; j < 5; j++) {
        if (j >= nw) break;
        const uint32_t v = sh ? ((S[j] << (32u - sh)) | (S[j + 1] >> sh)) : S[j + 1];
        if (v) or_fn(seq_words + w0 + j, v);
    }
    if (o.bd) {
        const uint32_t b0 = pos0 >> 5, bs = pos0 & 31u;
        const uint64_t l64 = o.bd << bs;
        const uint32_t v0 = (uint32_t)l64, v1 = (uint32_t)(l64 >> 32), v2 = bs ? (uint32_t)(o.bd >> (64u - bs)) : 0u;
        if (v0) or_fn(bad_words + b0, v0);
        if (v1) or_fn(bad_words + b0 + 1, v1);
        if (v2) or_fn(bad_words + b0 + 2, v2);
    }
}
