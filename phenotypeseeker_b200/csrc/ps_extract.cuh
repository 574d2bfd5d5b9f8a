// ps_extract.cuh — canonical k-mer extraction from the 2-bit packed stream pool.
//
// Replaces the k-mer enumeration half of `glistmaker` (modeling.py:303-315): every stream
// position q whose window [q, q+k) holds no invalid position yields
// canonical = min(word, revcomp(word)), 2 bits/base, first base most significant.
// k <= 16 packs into u32 keys, k <= 32 into u64 keys.
//
// One block = 4096 consecutive positions (never straddles samples: streams are padded to
// 4096). Warp w walks 512 positions in 16 coalesced steps; valid & in-range k-mers are
// compacted in stream order (ballot + popc), so the output order is deterministic and a
// stable sort keeps every sample's instances of one k-mer adjacent.
#pragma once
#include "ps_common.cuh"

#define EXT_THREADS 256
#define EXT_ITERS 16
#define EXT_BLOCK_POS (EXT_THREADS * EXT_ITERS)  // 4096 == POS_ALIGN

template <typename KeyT> struct KeyTraits;
template <> struct KeyTraits<uint32_t> { static constexpr int BITS = 32; };
template <> struct KeyTraits<uint64_t> { static constexpr int BITS = 64; };

__device__ __forceinline__ uint32_t rev2_32(uint32_t x) {
    x = __brev(x);
    return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}
__device__ __forceinline__ uint64_t rev2_64(uint64_t x) {
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// Canonical k-mer starting at pool position q; returns false if the window is invalid.
template <typename KeyT>
__device__ __forceinline__ bool kmer_at(const uint32_t *__restrict__ seq,
                                        const uint32_t *__restrict__ bad, uint64_t q, int k,
                                        KeyT &out) {
    const uint64_t mw = q >> 5;
    const uint32_t m0 = __ldg(bad + mw), m1 = __ldg(bad + mw + 1);
    const uint32_t win = __funnelshift_r(m0, m1, (uint32_t)(q & 31));
    const uint32_t kmask = k == 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
    if (win & kmask) return false;
    const uint64_t sw = q >> 4;
    const uint32_t sh = 2u * (uint32_t)(q & 15);
    const uint32_t w0 = __ldg(seq + sw), w1 = __ldg(seq + sw + 1);
    if (sizeof(KeyT) == 4) {
        const uint32_t f32 = __funnelshift_l(w1, w0, sh);
        const uint32_t fw = f32 >> (32 - 2 * k);
        const uint32_t rc = rev2_32(~fw) >> (32 - 2 * k);
        out = (KeyT)(fw < rc ? fw : rc);
    } else {
        const uint32_t w2 = __ldg(seq + sw + 2);
        const uint64_t f64 = ((uint64_t)__funnelshift_l(w1, w0, sh) << 32) | __funnelshift_l(w2, w1, sh);
        const uint64_t fw = f64 >> (64 - 2 * k);
        const uint64_t rc = rev2_64(~fw) >> (64 - 2 * k);
        out = (KeyT)(fw < rc ? fw : rc);
    }
    return true;
}

// COUNT pass: per-block number of valid in-range k-mers.
// WRITE pass: compacted keys at blk_offs[block]. TAGS 0: keys only; 1: keys + u16 sample tags
// (two arrays); 2: packed 64-bit records (key << 16 | tag) written through keys_out.
template <typename KeyT, bool WRITE, int TAGS>
__global__ void __launch_bounds__(EXT_THREADS)
k_extract(const uint32_t *__restrict__ seq, const uint32_t *__restrict__ bad, uint64_t pos_begin,
          int k, uint64_t lo, uint64_t hi, int range_all, const uint16_t *__restrict__ blk_sample,
          uint32_t *__restrict__ blk_counts, const uint64_t *__restrict__ blk_offs,
          KeyT *__restrict__ keys_out, uint16_t *__restrict__ tags_out) {
    __shared__ uint32_t wsum[EXT_THREADS / 32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t base = pos_begin + (uint64_t)blockIdx.x * EXT_BLOCK_POS + warp * (32 * EXT_ITERS);
    KeyT keys[EXT_ITERS];
    uint32_t validbits = 0, wcount = 0;
    uint16_t off[EXT_ITERS];
#pragma unroll
    for (int it = 0; it < EXT_ITERS; it++) {
        KeyT key = 0;
        bool ok = kmer_at<KeyT>(seq, bad, base + it * 32 + lane, k, key);
        if (ok && !range_all) ok = (uint64_t)key >= lo && (uint64_t)key < hi;
        const unsigned ball = __ballot_sync(0xffffffffu, ok);
        if (WRITE) {
            keys[it] = key;
            off[it] = (uint16_t)(wcount + __popc(ball & lanemask_lt()));
            validbits |= (ok ? 1u : 0u) << it;
        }
        wcount += __popc(ball);
    }
    if (lane == 0) wsum[warp] = wcount;
    __syncthreads();
    if (!WRITE) {
        if (threadIdx.x == 0) {
            uint32_t s = 0;
#pragma unroll
            for (int w2 = 0; w2 < EXT_THREADS / 32; w2++) s += wsum[w2];
            blk_counts[blockIdx.x] = s;
        }
        return;
    }
    uint64_t o = blk_offs[blockIdx.x];
    for (unsigned w2 = 0; w2 < warp; w2++) o += wsum[w2];
    const uint16_t tag = TAGS ? blk_sample[(pos_begin >> 12) + blockIdx.x] : 0;
#pragma unroll
    for (int it = 0; it < EXT_ITERS; it++) {
        if ((validbits >> it) & 1u) {
            if (TAGS == 2) {
                reinterpret_cast<uint64_t *>(keys_out)[o + off[it]] = ((uint64_t)keys[it] << 16) | tag;
            } else {
                keys_out[o + off[it]] = keys[it];
                if (TAGS == 1) tags_out[o + off[it]] = tag;
            }
        }
    }
}

// Direct variant for whole-range runs over assemblies (almost every window is valid): one
// packed record per POSITION, no count pass and no compaction — an invalid window becomes the
// all-ones sentinel record, which sorts behind every real k-mer (an all-T word is never canonical)
// and is skipped by the row builder. The digit histograms of all radix passes are accumulated on
// the way (shared-memory counters, flushed once per block), so the sort needs no histogram read.
template <typename KeyT>
__global__ void __launch_bounds__(EXT_THREADS)
k_extract_direct(const uint32_t *__restrict__ seq, const uint32_t *__restrict__ bad, uint64_t pos_begin,
                 int k, const uint16_t *__restrict__ blk_sample, uint64_t out_base,
                 uint64_t *__restrict__ recs_out, int npass, int rb, int shift0, unsigned long long *__restrict__ hist) {
    __shared__ uint32_t sh[8][512];
    const uint32_t dmask = (1u << rb) - 1u;
    for (int i = threadIdx.x; i < 8 * 512; i += EXT_THREADS) (&sh[0][0])[i] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t local = (uint64_t)blockIdx.x * EXT_BLOCK_POS + warp * (32 * EXT_ITERS);
    const uint64_t base = pos_begin + local;
    const uint64_t tag = blk_sample[(pos_begin >> 12) + blockIdx.x];
    if (sizeof(KeyT) == 4 && EXT_ITERS == 16) {
        // k <= 16: the warp's 512 positions are 32 sequence words + 17 mask words. Each lane loads one
        // of each, and the two words a window needs come by shuffle (word index it*2 + lane/16 and the
        // next one; mask words it and it+1) instead of four loads with 64-bit address arithmetic per position.
        const uint64_t wbase = base >> 4;                     // base is a multiple of 512 positions
        const uint32_t myw = __ldg(seq + wbase + lane);
        const uint32_t wext = __ldg(seq + wbase + 32);
        const uint32_t myb = __ldg(bad + (base >> 5) + min(lane, 16u));
        const uint32_t shl = 2u * (lane & 15u), half = lane >> 4;
        const uint32_t kmask = k == 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
        uint64_t *dst = recs_out + out_base + local + lane;
#pragma unroll
        for (int it = 0; it < EXT_ITERS; it++) {
            const uint32_t w0 = __shfl_sync(0xffffffffu, myw, it * 2 + half);
            uint32_t w1 = __shfl_sync(0xffffffffu, myw, (it * 2 + half + 1) & 31);
            if (it == EXT_ITERS - 1 && half) w1 = wext;
            const uint32_t m0 = __shfl_sync(0xffffffffu, myb, it), m1 = __shfl_sync(0xffffffffu, myb, it + 1);
            const bool ok = (__funnelshift_r(m0, m1, lane) & kmask) == 0;
            const uint32_t fw = __funnelshift_l(w1, w0, shl) >> (32 - 2 * k);
            const uint32_t rc = rev2_32(~fw) >> (32 - 2 * k);
            const uint64_t rec = ok ? (((uint64_t)(fw < rc ? fw : rc) << 16) | tag) : ~0ull;
            dst[it * 32] = rec;
            if (npass == 2) {        // bucketed build: the two partition digits
                atomicAdd(&sh[0][(uint32_t)(rec >> shift0) & dmask], 1u);
                atomicAdd(&sh[1][(uint32_t)(rec >> (shift0 + rb)) & dmask], 1u);
            } else {
                for (int p = 0; p < npass; p++) atomicAdd(&sh[p][(uint32_t)(rec >> (shift0 + rb * p)) & dmask], 1u);
            }
        }
    } else {
#pragma unroll 4
        for (int it = 0; it < EXT_ITERS; it++) {
            KeyT key = 0;
            const bool ok = kmer_at<KeyT>(seq, bad, base + it * 32 + lane, k, key);
            const uint64_t rec = ok ? (((uint64_t)key << 16) | tag) : ~0ull;
            recs_out[out_base + local + it * 32 + lane] = rec;
            for (int p = 0; p < npass; p++) atomicAdd(&sh[p][(uint32_t)(rec >> (shift0 + rb * p)) & dmask], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 512; i += EXT_THREADS) {
        const uint32_t v = (&sh[0][0])[i];
        if (v) atomicAdd(&hist[i], (unsigned long long)v);
    }
}

// Multi-GPU routing: packed records of this rank's own samples, split by destination k-mer range
// (nparts - 1 ascending splitters; destination d owns [spl[d-1], spl[d])). Output layout is
// destination-major and, inside one destination, stream order — so after the all-to-all the
// records of one k-mer are still in sample order. COUNT: blk_counts[d * nblk_total + block];
// WRITE: records at blk_offs[d * nblk_total + block] (exclusive scan of the counts).
#define PART_MAX 8
// Destination table of the WRITE pass: destination d's records go to ptr[d] + adj[d] + (scan offset).
// ptr[d] may be a peer GPU's receive buffer mapped through CUDA IPC: the routing then happens inside
// this kernel as plain 8-byte stores over NVLink, overlapped with the extraction of the next k-mers.
struct PartDst {
    uint64_t *ptr[PART_MAX];
    long long adj[PART_MAX];
};
template <typename KeyT, bool WRITE>
__global__ void __launch_bounds__(EXT_THREADS)
k_extract_part(const uint32_t *__restrict__ seq, const uint32_t *__restrict__ bad, uint64_t pos_begin, int k,
               const uint16_t *__restrict__ blk_sample, int nparts, const uint64_t *__restrict__ splitters,
               uint32_t *__restrict__ blk_counts, const uint64_t *__restrict__ blk_offs,
               uint64_t nblk_total, uint64_t blk0, PartDst dst) {
    __shared__ uint32_t wsum[EXT_THREADS / 32][PART_MAX];
    __shared__ uint64_t spl[PART_MAX];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < PART_MAX) spl[threadIdx.x] = (int)threadIdx.x < nparts - 1 ? splitters[threadIdx.x] : ~0ull;
    __syncthreads();
    const uint64_t base = pos_begin + (uint64_t)blockIdx.x * EXT_BLOCK_POS + warp * (32 * EXT_ITERS);
    KeyT keys[EXT_ITERS];
    uint16_t off[EXT_ITERS];
    uint32_t dests = 0;                     // 4 bits per item: destination, 0xF = invalid
    uint32_t dests_hi = 0;
    uint32_t wcount[PART_MAX];
#pragma unroll
    for (int p = 0; p < PART_MAX; p++) wcount[p] = 0;
#pragma unroll
    for (int it = 0; it < EXT_ITERS; it++) {
        KeyT key = 0;
        const bool ok = kmer_at<KeyT>(seq, bad, base + it * 32 + lane, k, key);
        uint32_t d = 0;
#pragma unroll
        for (int p = 0; p < PART_MAX - 1; p++) d += ((uint64_t)key >= spl[p]) ? 1u : 0u;
        if (!ok) d = 0xF;
        uint32_t myoff = 0;
#pragma unroll
        for (int p = 0; p < PART_MAX; p++) {
            if (p < nparts) {
                const unsigned ball = __ballot_sync(0xffffffffu, d == (uint32_t)p);
                if (d == (uint32_t)p) myoff = wcount[p] + __popc(ball & lanemask_lt());
                wcount[p] += __popc(ball);
            }
        }
        if (WRITE) {
            keys[it] = key;
            off[it] = (uint16_t)myoff;
            if (it < 8) dests |= d << (4 * it); else dests_hi |= d << (4 * (it - 8));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int p = 0; p < PART_MAX; p++) wsum[warp][p] = wcount[p];
    }
    __syncthreads();
    const uint64_t blk = blk0 + blockIdx.x;
    if (!WRITE) {
        if ((int)threadIdx.x < nparts) {
            uint32_t sum = 0;
            for (int w2 = 0; w2 < EXT_THREADS / 32; w2++) sum += wsum[w2][threadIdx.x];
            blk_counts[(uint64_t)threadIdx.x * nblk_total + blk] = sum;
        }
        return;
    }
    // Stage the block's records in shared memory grouped by destination, then write every group
    // with consecutive lanes on consecutive addresses: stores to a peer GPU leave as full
    // 128-byte NVLink packets instead of a few 8-byte pieces per warp.
    extern __shared__ __align__(16) uint64_t srec[];          // EXT_BLOCK_POS records
    __shared__ uint32_t dstart[PART_MAX + 1];
    __shared__ uint64_t gbase[PART_MAX];
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int p = 0; p < PART_MAX; p++) {
            dstart[p] = run;
            if (p < nparts) for (int w2 = 0; w2 < EXT_THREADS / 32; w2++) run += wsum[w2][p];
        }
        dstart[PART_MAX] = run;
    }
    if ((int)threadIdx.x < nparts)
        gbase[threadIdx.x] = blk_offs[(uint64_t)threadIdx.x * nblk_total + blk] + (uint64_t)dst.adj[threadIdx.x];
    __syncthreads();
    uint32_t wloc[PART_MAX];
#pragma unroll
    for (int p = 0; p < PART_MAX; p++) {
        uint32_t o = dstart[p];
        if (p < nparts) for (unsigned w2 = 0; w2 < warp; w2++) o += wsum[w2][p];
        wloc[p] = o;
    }
    const uint64_t tag = blk_sample[(pos_begin >> 12) + blockIdx.x];
#pragma unroll
    for (int it = 0; it < EXT_ITERS; it++) {
        const uint32_t d = ((it < 8 ? dests >> (4 * it) : dests_hi >> (4 * (it - 8)))) & 0xFu;
        if (d != 0xFu) {
            uint32_t o = 0;
#pragma unroll
            for (int p = 0; p < PART_MAX; p++) if (d == (uint32_t)p) o = wloc[p];
            srec[o + off[it]] = ((uint64_t)keys[it] << 16) | tag;
        }
    }
    __syncthreads();
    const uint32_t total = dstart[PART_MAX];
    for (uint32_t j = threadIdx.x; j < total; j += EXT_THREADS) {
        int d = 0;
#pragma unroll
        for (int p = 1; p < PART_MAX; p++) d += (j >= dstart[p] && p < nparts) ? 1 : 0;
        uint64_t *out = nullptr;
#pragma unroll
        for (int p = 0; p < PART_MAX; p++) if (d == p) out = dst.ptr[p];
        out[gbase[d] + (j - dstart[d])] = srec[j];
    }
}

// Occurrence counts of K sorted query k-mers within [pos_begin, pos_begin + nblocks*4096).
template <typename KeyT>
__global__ void __launch_bounds__(EXT_THREADS)
k_lookup(const uint32_t *__restrict__ seq, const uint32_t *__restrict__ bad, uint64_t pos_begin,
         int k, const uint64_t *__restrict__ queries, int K, uint32_t *__restrict__ counts) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t base = pos_begin + (uint64_t)blockIdx.x * EXT_BLOCK_POS + warp * (32 * EXT_ITERS);
    for (int it = 0; it < EXT_ITERS; it++) {
        KeyT key = 0;
        if (!kmer_at<KeyT>(seq, bad, base + it * 32 + lane, k, key)) continue;
        int a = 0, b = K;  // lower bound
        while (a < b) {
            int m = (a + b) >> 1;
            if (__ldg(queries + m) < (uint64_t)key) a = m + 1; else b = m;
        }
        if (a < K && __ldg(queries + a) == (uint64_t)key) atomicAdd(counts + a, 1u);
    }
}

// Gather per-sample counted lists (list mode) into the instance arrays, range-filtered, in
// list order. One block handles 4096 list entries of one sample (lists padded likewise).
template <typename KeyT, bool WRITE, int TAGS>
__global__ void __launch_bounds__(EXT_THREADS)
k_list_gather(const KeyT *__restrict__ list_keys, uint64_t ent_begin, uint64_t lo, uint64_t hi,
              int range_all, const uint16_t *__restrict__ blk_sample,
              const uint32_t *__restrict__ blk_valid, uint32_t *__restrict__ blk_counts,
              const uint64_t *__restrict__ blk_offs, KeyT *__restrict__ keys_out,
              uint16_t *__restrict__ tags_out) {
    __shared__ uint32_t wsum[EXT_THREADS / 32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t base = ent_begin + (uint64_t)blockIdx.x * EXT_BLOCK_POS + warp * (32 * EXT_ITERS);
    const uint32_t nvalid = blk_valid[blockIdx.x];  // valid entries in this block
    KeyT keys[EXT_ITERS];
    uint32_t validbits = 0, wcount = 0;
    uint16_t off[EXT_ITERS];
#pragma unroll
    for (int it = 0; it < EXT_ITERS; it++) {
        const uint32_t li = warp * (32 * EXT_ITERS) + it * 32 + lane;
        KeyT key = 0;
        bool ok = li < nvalid;
        if (ok) {
            key = list_keys[base + it * 32 + lane];
            if (!range_all) ok = (uint64_t)key >= lo && (uint64_t)key < hi;
        }
        const unsigned ball = __ballot_sync(0xffffffffu, ok);
        if (WRITE) {
            keys[it] = key;
            off[it] = (uint16_t)(wcount + __popc(ball & lanemask_lt()));
            validbits |= (ok ? 1u : 0u) << it;
        }
        wcount += __popc(ball);
    }
    if (lane == 0) wsum[warp] = wcount;
    __syncthreads();
    if (!WRITE) {
        if (threadIdx.x == 0) {
            uint32_t s = 0;
            for (int w2 = 0; w2 < EXT_THREADS / 32; w2++) s += wsum[w2];
            blk_counts[blockIdx.x] = s;
        }
        return;
    }
    uint64_t o = blk_offs[blockIdx.x];
    for (unsigned w2 = 0; w2 < warp; w2++) o += wsum[w2];
    const uint16_t tag = blk_sample[blockIdx.x];
#pragma unroll
    for (int it = 0; it < EXT_ITERS; it++) {
        if ((validbits >> it) & 1u) {
            if (TAGS == 2) {
                reinterpret_cast<uint64_t *>(keys_out)[o + off[it]] = ((uint64_t)keys[it] << 16) | tag;
            } else {
                keys_out[o + off[it]] = keys[it];
                tags_out[o + off[it]] = tag;
            }
        }
    }
}
