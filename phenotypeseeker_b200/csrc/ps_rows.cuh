// ps_rows.cuh — sorted (k-mer, sample) instances -> union k-mer array + k-mer-major bit matrix.
//
// Replaces `glistcompare -u` (modeling.py:351-380: the union, ascending = feature_vector.list
// order), `glistquery -l` + `split` (modeling.py:317-348: per-sample presence of every union
// k-mer) and the presence binarisation of get_kmers_tested (modeling.py:693-695).
//
// Row r of the matrix belongs to union k-mer r; sample s is bit (s & 31) of word (s >> 5);
// rows are `wp` words apart (multiple of 4 => 16-byte aligned rows for 128-bit loads).
// Duplicate instances of one (k-mer, sample) are dropped with a neighbour compare, the bits
// of one 32-instance step are merged per (row, word) with __match_any_sync +
// __reduce_or_sync, and one atomicOr per distinct (row, word) reaches memory.
#pragma once
#include "ps_common.cuh"
#include "ps_sort.cuh"

// PACKED: `keys` holds 64-bit records (key << 16 | tag) and `tags` is unused.
template <typename KeyT, bool PACKED>
__global__ void __launch_bounds__(RUN_THREADS)
k_row_build(const KeyT *__restrict__ keys, const uint16_t *__restrict__ tags, uint64_t n,
            const unsigned long long *__restrict__ chunk_offs, uint64_t *__restrict__ union_out,
            uint32_t *__restrict__ matrix, int wp) {
    const unsigned lane = threadIdx.x & 31;
    const uint64_t chunk = (uint64_t)blockIdx.x * (RUN_THREADS / 32) + (threadIdx.x >> 5);
    const uint64_t base = chunk * RUN_CHUNK;
    if (base >= n) return;
    unsigned long long run = chunk_offs[chunk];  // heads before this chunk
    for (int it = 0; it < RUN_CHUNK / 32; it++) {
        const uint64_t i = base + it * 32 + lane;
        bool valid = i < n;
        KeyT kk = 0;
        uint16_t tag = 0;
        bool head = false, dup = false;
        if (valid) {
            if (PACKED) {
                const KeyT rec = keys[i];
                kk = rec >> 16;
                tag = (uint16_t)(rec & 0xFFFFu);
                if (kk == (KeyT(~KeyT(0)) >> 16)) valid = false;   // sentinel of k_extract_direct
                else if (i == 0) head = true;
                else {
                    const KeyT prev = keys[i - 1];
                    head = kk != (prev >> 16);
                    dup = rec == prev;
                }
            } else {
                kk = keys[i];
                tag = tags[i];
                if (i == 0) head = true;
                else {
                    head = kk != keys[i - 1];
                    dup = !head && tags[i - 1] == tag;
                }
            }
        }
        const unsigned ball = __ballot_sync(0xffffffffu, head);
        const uint32_t delta = __popc(ball & lanemask_le());  // heads up to and including me
        const unsigned long long row = run + delta - 1;       // valid lanes only
        if (head) union_out[row] = (uint64_t)kk;
        const bool active = valid && !dup;
        const uint32_t word = tag >> 5;
        const uint32_t id = active ? ((delta << 11) | word) : 0xFFFFFFFFu;
        const unsigned grp = __match_any_sync(0xffffffffu, id);
        const uint32_t bits = __reduce_or_sync(grp, active ? (1u << (tag & 31)) : 0u);
        if (active && (int)lane == __ffs(grp) - 1) atomicOr(matrix + row * (uint64_t)wp + word, bits);
        run += __popc(ball);
    }
}

// Gather the matrix rows and k-mers of the survivors (slot order) for the D2H copy.
__global__ void k_gather_rows(const uint32_t *__restrict__ matrix, const uint64_t *__restrict__ uni,
                              const unsigned long long *__restrict__ sv_row, uint64_t ns, int wp,
                              uint32_t *__restrict__ out_bits, uint64_t *__restrict__ out_kmer) {
    const uint64_t total = ns * (uint64_t)wp;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = i / wp;
        const uint32_t w = (uint32_t)(i % wp);
        const unsigned long long r = sv_row[s];
        out_bits[i] = matrix[r * (uint64_t)wp + w];
        if (w == 0) out_kmer[s] = uni[r];
    }
}
