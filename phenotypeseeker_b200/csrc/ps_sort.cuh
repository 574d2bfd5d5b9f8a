// ps_sort.cuh — hand-written stable LSD radix sort (one-sweep, decoupled look-back) for
// (k-mer key [, u16 sample tag]) pairs, plus the small scan / run-length kernels around it.
//
// Replaces the counting core of `glistmaker` and the merge passes of `glistcompare -u`
// (modeling.py:303-315, 351-380): k-mer instances of ALL samples are sorted once; equal
// k-mers end up adjacent, in sample order because the sort is stable.
//
// One pass = one read + one write of every pair: a tile of 4096 pairs is ranked in shared
// memory (warp-private digit histograms; lanes with equal digits meet through one shared
// atomicOr on a per-warp mask word), tiles chain their per-digit prefixes through a 64-bit status|value look-back
// word, and the tile is written out digit-run by digit-run (coalesced).
#pragma once
#include "ps_common.cuh"

#ifndef RS_THREADS
#define RS_THREADS 512
#endif
#ifndef RS_ITEMS
#define RS_ITEMS 16
#endif
#define RS_WARPS (RS_THREADS / 32)
#define RS_TILE (RS_THREADS * RS_ITEMS)  // 4096
#define RS_RADIX 256          // 8-bit digits (default)
#define RS_MAX_RADIX 512      // 9-bit digits: used when they save a whole pass (e.g. k = 13: 26 bits)
#define RS_MAX_PASSES 8
#ifndef RS_MIN_BLOCKS
#define RS_MIN_BLOCKS (1024 / RS_THREADS)
#endif
#ifndef RS_LB_BATCH
#define RS_LB_BATCH 4
#endif
#ifndef RS_RANK_ATOMIC
#define RS_RANK_ATOMIC 0
#endif
#ifndef RS_MATCH_BALLOT
#define RS_MATCH_BALLOT 1
#endif

template <typename KeyT, bool HAS_VAL> constexpr size_t rs_dyn_smem() {
    return RS_TILE * sizeof(KeyT) + (HAS_VAL ? RS_TILE * 2 : 0);
}

#define LB_LOCAL (1ull << 62)
#define LB_INCL (2ull << 62)
#define LB_MASK ((1ull << 62) - 1)

// Digit histograms of every pass in one read of the keys.
template <typename KeyT>
__global__ void __launch_bounds__(512)
k_rs_hist(const KeyT *__restrict__ keys, uint64_t n, int npass, int shift0, int rb,
          unsigned long long *__restrict__ hist) {
    __shared__ uint32_t sh[RS_MAX_PASSES][RS_MAX_RADIX];
    const uint32_t dmask = (1u << rb) - 1u;
    for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_MAX_RADIX; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        KeyT key = keys[i];
        for (int p = 0; p < npass; p++) atomicAdd(&sh[p][(uint32_t)(key >> (shift0 + rb * p)) & dmask], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * RS_MAX_RADIX; i += blockDim.x) {
        uint32_t v = (&sh[0][0])[i];
        if (v) atomicAdd(&hist[i], (unsigned long long)v);
    }
}

// Exclusive scan of each pass's 256 bins: one block per pass.
__global__ void k_rs_scan(unsigned long long *__restrict__ hist) {
    __shared__ unsigned long long s[RS_MAX_RADIX];
    unsigned long long *h = hist + (size_t)blockIdx.x * RS_MAX_RADIX;
    s[threadIdx.x] = h[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < RS_MAX_RADIX; i++) { unsigned long long c = s[i]; s[i] = run; run += c; }
    }
    __syncthreads();
    h[threadIdx.x] = s[threadIdx.x];
}

template <typename KeyT, bool HAS_VAL, int RB>
__global__ void __launch_bounds__(RS_THREADS, ((sizeof(KeyT) == 8 && HAS_VAL && RS_MIN_BLOCKS > 1) ? RS_MIN_BLOCKS - 1 : RS_MIN_BLOCKS))
k_rs_pass(const KeyT *__restrict__ kin, KeyT *__restrict__ kout, const uint16_t *__restrict__ vin,
          uint16_t *__restrict__ vout, uint64_t n, int shift,
          const unsigned long long *__restrict__ gbase, unsigned long long *lookback,
          uint32_t *tile_counter) {
    constexpr int RADIX = 1 << RB;
    static_assert(RADIX <= RS_THREADS, "every digit needs an owner thread");
    __shared__ uint32_t whist[RS_WARPS][RADIX];  // per-warp digit counts, later scatter bases
    __shared__ uint32_t wmask[RS_WARPS][RADIX];  // per-warp lane mask per digit (self-clearing)
    extern __shared__ __align__(16) uint8_t rs_dyn[];  // RS_TILE keys, then RS_TILE u16 tags
    KeyT *skeys = reinterpret_cast<KeyT *>(rs_dyn);
    uint16_t *svals = reinterpret_cast<uint16_t *>(rs_dyn + RS_TILE * sizeof(KeyT));
    __shared__ unsigned long long goff[RADIX];
    __shared__ uint32_t wsum[RS_WARPS];
    __shared__ uint32_t s_tile;

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < RS_WARPS * RADIX; i += RS_THREADS) {
        (&whist[0][0])[i] = 0;
        (&wmask[0][0])[i] = 0;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_start = (uint64_t)tile * RS_TILE;
    const uint32_t nvalid = (uint32_t)min((uint64_t)RS_TILE, n - tile_start);

    // warp-striped load: stable order inside the tile is (warp, item, lane)
    KeyT key[RS_ITEMS];
    uint16_t rank[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const uint32_t idx = warp * (32 * RS_ITEMS) + i * 32 + lane;
        key[i] = idx < nvalid ? kin[tile_start + idx] : ~KeyT(0);
    }
    // Stable ranking inside the warp. Lanes holding the same digit find each other through
    // one shared-memory atomicOr on a per-warp mask word (MATCH.ANY costs ~one step per
    // distinct value and was the top stall of the first version of this kernel); the lowest
    // lane of each group bumps the warp-private digit counter and clears the mask.
    uint32_t *wh = whist[warp], *wm = wmask[warp];
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & (uint32_t)(RADIX - 1);
#if RS_RANK_ATOMIC
        rank[i] = (uint16_t)atomicAdd(wh + d, 1u);
#elif RS_MATCH_BALLOT
        // peers = lanes with the same digit: one ballot per digit bit, no shared memory
        unsigned peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < RB; b++) {
            const bool bit = (d >> b) & 1u;
            const unsigned bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
        }
        const uint32_t old = wh[d];
        __syncwarp();
        const uint32_t r = __popc(peers & lt);
        if (r == 0) wh[d] = old + __popc(peers);
        rank[i] = (uint16_t)(old + r);
        __syncwarp();
#else
        atomicOr(wm + d, 1u << lane);
        __syncwarp();
        const unsigned peers = wm[d];
        const uint32_t old = wh[d];
        __syncwarp();
        const uint32_t r = __popc(peers & lt);
        if (r == 0) {
            wh[d] = old + __popc(peers);
            wm[d] = 0;
        }
        rank[i] = (uint16_t)(old + r);
        __syncwarp();
#endif
    }
    // tags: issue the loads now, their latency hides behind the scan + look-back
    uint16_t val[HAS_VAL ? RS_ITEMS : 1];
    if (HAS_VAL) {
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) {
            const uint32_t idx = warp * (32 * RS_ITEMS) + i * 32 + lane;
            val[i] = idx < nvalid ? vin[tile_start + idx] : (uint16_t)0;
        }
    }

    __syncthreads();   // every warp finished ranking, whist is final

    // digit `tid`: tile count, published at once as this tile's LOCAL look-back entry
    const bool dig = tid < RADIX;   // threads that own a digit
    uint32_t cnt = 0;
    if (dig) {
#pragma unroll
        for (int w2 = 0; w2 < RS_WARPS; w2++) cnt += whist[w2][tid];
    }
    const uint32_t real = cnt - (((int)tid == RADIX - 1) ? (RS_TILE - nvalid) : 0u);  // padding keys are all-ones
    volatile unsigned long long *lb = lookback + (size_t)tile * RADIX + (tid & (RADIX - 1));
    if (dig) *lb = (tile == 0 ? LB_INCL : LB_LOCAL) | real;

    // block exclusive scan of cnt -> base of digit `tid` inside the tile
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31 && dig) wsum[warp] = inc;
    __syncthreads();
    uint32_t wp = 0;
#pragma unroll
    for (int w2 = 0; w2 < RADIX / 32; w2++) if (w2 < (int)warp) wp += wsum[w2];
    const uint32_t tb = wp + inc - cnt;
    // per-warp scatter base = tile base of the digit + keys of lower warps
    if (dig) {
        uint32_t run = tb;
#pragma unroll
        for (int w2 = 0; w2 < RS_WARPS; w2++) {
            const uint32_t c = whist[w2][tid];
            whist[w2][tid] = run;
            run += c;
        }
    }

    __syncthreads();

    // scatter into shared memory in tile-sorted order (needs only tile-local bases) ...
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & (uint32_t)(RADIX - 1);
        const uint32_t pos = wh[d] + rank[i];
        skeys[pos] = key[i];
        rank[i] = (uint16_t)pos;
    }
    if (HAS_VAL) {
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) svals[rank[i]] = val[i];
    }

    // ... and only then the decoupled look-back for digit `tid`: the later it runs, the more
    // predecessors already hold INCLUSIVE prefixes. Four predecessor entries are fetched per
    // step so that the walk is not one dependent L2 round trip per tile.
    unsigned long long excl = 0;
    if (tile > 0 && dig) {
        int64_t t = (int64_t)tile - 1;
        bool done = false;
        while (!done) {
            unsigned long long v[RS_LB_BATCH];
#pragma unroll
            for (int j = 0; j < RS_LB_BATCH; j++) {
                const int64_t tj = t - j;
                v[j] = tj >= 0 ? *(volatile unsigned long long *)(lookback + (size_t)tj * RADIX + tid) : LB_INCL;
            }
#pragma unroll
            for (int j = 0; j < RS_LB_BATCH; j++) {
                if (done) break;
                if ((v[j] >> 62) == 0) break;          // not published yet: re-read from here
                excl += v[j] & LB_MASK;
                t--;
                if ((v[j] >> 62) == 2) done = true;
            }
        }
        *lb = LB_INCL | (excl + real);
    }
    if (dig) goff[tid] = gbase[tid] + excl - tb;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const uint32_t p = i * RS_THREADS + tid;
        if (p < nvalid) {
            const KeyT kk = skeys[p];
            const uint32_t d = (uint32_t)(kk >> shift) & (uint32_t)(RADIX - 1);
            const unsigned long long dst = goff[d] + p;
            kout[dst] = kk;
            if (HAS_VAL) vout[dst] = svals[p];
        }
    }
}

// ---------------------------------------------------------------------------------------
// Partition pass for the bucketed build (ps_rows.cuh): same one-sweep structure as k_rs_pass
// (tile staged in shared memory, decoupled look-back per digit, digit-run write-out), but the
// order of equal digits INSIDE a tile is not kept — the bucket kernels OR presence bits and do not
// care — so a record is ranked with a single shared-memory atomicAdd on a block-wide digit
// counter instead of eight ballots + per-warp counters (k_rs_pass issues ~118 instructions per
// record slot and is issue-bound; this pass needs about a quarter of that).
//
// Two passes order the records by 16 bits: pass 1 by the low digit, pass 2 (SEGMENTED) by the
// high digit. Pass 2 must keep the low-digit order among equal high digits; since only tile
// order is kept, its tiles never straddle a low-digit segment of its input (seg_start = pass 1's
// digit offsets; every segment owns at least one, possibly empty, tile). That also makes the
// bucket table free: the first tile of segment L knows, per high digit d, how many records of d
// lie in lower segments — its exclusive look-back prefix — so bstart[d * 256 + L] = gbase[d] + excl.
// OUT32: pass 2 drops the 16 sorted k-mer bits and stores (low k-mer bits << 16 | sample) in 4 bytes.
#ifndef PP_THREADS
#define PP_THREADS 512
#endif
#ifndef PP_ITEMS
#define PP_ITEMS 16
#endif
#ifndef PP_MIN_BLOCKS
#define PP_MIN_BLOCKS 2
#endif
#ifndef PP_MIN_BLOCKS32
#define PP_MIN_BLOCKS32 3
#endif
#ifndef PP_LB_BATCH
#define PP_LB_BATCH 4
#endif
#ifndef PP_PREFETCH_TILES
#define PP_PREFETCH_TILES 150      // about half of the 2 x 148 resident tiles (measured: 0 -> 10.6, 150 -> 9.2, 296 -> 9.3, 600 -> 10.7 ms)
#endif
#define PP_TILE (PP_THREADS * PP_ITEMS)

// cumulative tile counts of the 256 input segments of pass 2; seg_tile0[256] = total
__global__ void k_part_segments(const unsigned long long *__restrict__ seg_start, uint64_t n,
                                uint32_t *__restrict__ seg_tile0) {
    __shared__ uint32_t s[256];
    const unsigned t = threadIdx.x;
    const unsigned long long a = seg_start[t], b = t == 255 ? n : seg_start[t + 1];
    const uint32_t tiles = (uint32_t)max(1ull, (b - a + PP_TILE - 1) / PP_TILE);
    s[t] = tiles;
    __syncthreads();
    if (t == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 256; i++) { const uint32_t c = s[i]; s[i] = run; run += c; }
        seg_tile0[256] = run;
    }
    __syncthreads();
    seg_tile0[t] = s[t];
}

// Record formats. In: PP_REC64 = k-mer << 16 | sample (extraction), PP_NARROW = pass-2 digit << 24 |
// low k-mer bits << 8 | sample (only when every sample id fits 8 bits: n_samples <= 255 — pass 1 then
// already halves the record, and pass 2 moves 4-byte records through shared memory with three
// resident tiles per SM). Out: PP_REC64, PP_NARROW (pass 1) or PP_BUCKET = low k-mer bits << 16 |
// sample (pass 2, what the bucket kernels read). All-ones = invalid window in every format.
enum { PP_REC64 = 0, PP_BUCKET = 1, PP_NARROW = 2 };
template <typename InT, bool SEGMENTED, int OUTF, int EXP = 0>
__global__ void __launch_bounds__(PP_THREADS, sizeof(InT) == 4 ? PP_MIN_BLOCKS32 : PP_MIN_BLOCKS)
k_part_pass(const InT *__restrict__ in, void *__restrict__ out, uint64_t n, int shift,
            const unsigned long long *__restrict__ gbase, const unsigned long long *__restrict__ seg_start,
            const uint32_t *__restrict__ seg_tile0, unsigned long long *lookback, uint32_t *tile_counter,
            unsigned long long *__restrict__ bstart, int lbits) {
    extern __shared__ __align__(16) uint8_t pp_dyn[];
    InT *skeys = reinterpret_cast<InT *>(pp_dyn);      // PP_TILE records
    constexpr bool IN32 = sizeof(InT) == 4;
#define PP_DIGIT(key) (IN32 ? (uint32_t)((key) >> 24) : ((uint32_t)((uint64_t)(key) >> shift) & 255u))
    __shared__ uint32_t hist[257];
    __shared__ unsigned long long goff[256];
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t s_tile;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < 257; i += PP_THREADS) hist[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    // The tile a resident block will take PP_PREFETCH_TILES tickets from now: pull it into L2 so that
    // its loads are L2 hits instead of DRAM round trips (the pass is latency-bound: load, rank, scatter,
    // look-back and write-out of one tile are serialised by barriers). One 128-byte line per thread.
    // Pass 2 tiles lag the linear position by at most 256 padding tiles; close enough for a hint.
    if (PP_PREFETCH_TILES) {
        const uint64_t pf = ((uint64_t)tile + PP_PREFETCH_TILES) * PP_TILE + (uint64_t)tid * (128 / sizeof(InT));
        if (tid * (128 / sizeof(InT)) < PP_TILE && pf < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(in + pf));
    }
    uint64_t tile_start;
    uint32_t nvalid, seg = 0;
    bool first_of_seg = false;
    if (SEGMENTED) {
        if (tile >= seg_tile0[256]) return;
        uint32_t lo = 0, hi = 255;                  // largest seg with seg_tile0[seg] <= tile
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (seg_tile0[mid] <= tile) lo = mid; else hi = mid - 1;
        }
        seg = lo;
        const uint32_t tin = tile - seg_tile0[seg];
        first_of_seg = tin == 0;
        const unsigned long long a = seg_start[seg], b = seg == 255 ? n : seg_start[seg + 1];
        tile_start = a + (uint64_t)tin * PP_TILE;
        nvalid = tile_start < b ? (uint32_t)min((unsigned long long)PP_TILE, b - tile_start) : 0u;
    } else {
        tile_start = (uint64_t)tile * PP_TILE;
        nvalid = (uint32_t)min((uint64_t)PP_TILE, n - tile_start);
    }

    // slots past the end of the tile count in bin 256, whose base after the scan is nvalid: they land
    // behind the valid records in shared memory and are never written out — no branches per slot
    InT key[PP_ITEMS];
    uint32_t rank2[PP_ITEMS / 2];            // two 16-bit ranks per register
    const uint32_t idx0 = warp * (32 * PP_ITEMS) + lane;
    const InT *src = in + tile_start + idx0;
#pragma unroll
    for (int i = 0; i < PP_ITEMS; i++) key[i] = idx0 + i * 32 < nvalid ? src[i * 32] : InT(0);
#pragma unroll
    for (int i = 0; i < PP_ITEMS; i++) {
        const uint32_t d = idx0 + i * 32 < nvalid ? PP_DIGIT(key[i]) : 256u;
        const uint32_t r = atomicAdd(&hist[d], 1u);
        if (i & 1) rank2[i >> 1] |= r << 16; else rank2[i >> 1] = r;
    }
    __syncthreads();

    // digit `tid`: tile count -> LOCAL look-back entry, block scan -> base inside the tile
    const bool dig = tid < 256;
    const uint32_t cnt = dig ? hist[tid] : 0u;
    volatile unsigned long long *lb = lookback + (size_t)tile * 256 + (tid & 255);
    if (dig) *lb = (tile == 0 ? LB_INCL : LB_LOCAL) | cnt;
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31 && dig) wsum[warp] = inc;
    __syncthreads();
    uint32_t wp = 0;
#pragma unroll
    for (int w2 = 0; w2 < 8; w2++) if (w2 < (int)warp) wp += wsum[w2];
    const uint32_t tb = wp + inc - cnt;
    if (dig) hist[tid] = tb;
    if (tid == PP_THREADS - 1) hist[256] = nvalid;
    __syncthreads();

#pragma unroll
    for (int i = 0; i < PP_ITEMS; i++) {
        const uint32_t d = idx0 + i * 32 < nvalid ? PP_DIGIT(key[i]) : 256u;
        skeys[hist[d] + ((i & 1) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xFFFFu))] = key[i];
    }

    unsigned long long excl = 0;
    if (dig) {
        if (EXP & 1) excl = (unsigned long long)tile * 24;      // timing experiment: no look-back
        else if (tile > 0) {
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                unsigned long long v[PP_LB_BATCH];
#pragma unroll
                for (int j = 0; j < PP_LB_BATCH; j++) {
                    const int64_t tj = t - j;
                    v[j] = tj >= 0 ? *(volatile unsigned long long *)(lookback + (size_t)tj * 256 + tid) : LB_INCL;
                }
#pragma unroll
                for (int j = 0; j < PP_LB_BATCH; j++) {
                    if (done) break;
                    if ((v[j] >> 62) == 0) break;          // not published yet: re-read from here
                    excl += v[j] & LB_MASK;
                    t--;
                    if ((v[j] >> 62) == 2) done = true;
                }
            }
            *lb = LB_INCL | (excl + cnt);
        }
        const unsigned long long g = gbase[tid] + excl;
        goff[tid] = g - tb;
        if (SEGMENTED && first_of_seg) bstart[tid * 256 + seg] = g;
    }
    if (SEGMENTED && tile == 0 && tid == 0) bstart[65536] = n;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PP_ITEMS; i++) {
        const uint32_t p = i * PP_THREADS + tid;
        if (p < nvalid) {
            const InT kk = skeys[p];
            const uint32_t d = PP_DIGIT(kk);
            unsigned long long dst = goff[d] + p;
            if (EXP & 2) dst = tile_start + p;                      // timing experiment: sequential write-out
            if (EXP) dst %= n;
            if (OUTF == PP_REC64) {
                reinterpret_cast<uint64_t *>(out)[dst] = (uint64_t)kk;
            } else if (OUTF == PP_NARROW) {      // pass 1, from 64-bit records
                const uint64_t k64 = (uint64_t)kk;
                const uint32_t low = (uint32_t)(k64 >> 16) & ((1u << lbits) - 1u);
                const uint32_t hi8 = (uint32_t)(k64 >> (shift + 8)) & 255u;
                reinterpret_cast<uint32_t *>(out)[dst] = k64 == ~0ull ? ~0u : ((hi8 << 24) | (low << 8) | ((uint32_t)k64 & 0xFFu));
            } else if (IN32) {                   // pass 2, from narrow records
                const uint32_t k32 = (uint32_t)kk;
                reinterpret_cast<uint32_t *>(out)[dst] = k32 == ~0u ? ~0u : ((((k32 >> 8) & 0xFFFFu) << 16) | (k32 & 0xFFu));
            } else {                             // pass 2, from 64-bit records
                const uint64_t k64 = (uint64_t)kk;
                const uint32_t low = (uint32_t)(k64 >> 16) & ((1u << lbits) - 1u);
                reinterpret_cast<uint32_t *>(out)[dst] = k64 == ~0ull ? ~0u : ((low << 16) | ((uint32_t)k64 & 0xFFFFu));
            }
        }
    }
#undef PP_DIGIT
}

// ---------------------------------------------------------------------------------------
// Exclusive scan u32 counts -> u64 offsets (offs[n] = total). Single block, 1024 threads.
__global__ void __launch_bounds__(1024)
k_scan_counts(const uint32_t *__restrict__ counts, uint64_t n, unsigned long long *__restrict__ offs) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    constexpr int PER = 8;
    for (uint64_t base = 0; base < n; base += 1024ull * PER) {
        uint32_t v[PER];
        unsigned long long tsum = 0;
        const uint64_t i0 = base + (uint64_t)tid * PER;
#pragma unroll
        for (int j = 0; j < PER; j++) { v[j] = (i0 + j < n) ? counts[i0 + j] : 0u; tsum += v[j]; }
        unsigned long long inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = wsum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= (unsigned)o) wi += t;
            }
            wsum[lane] = wi - w;  // exclusive
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        unsigned long long run = carry + wsum[warp] + inc - tsum;
#pragma unroll
        for (int j = 0; j < PER; j++) { if (i0 + j < n) offs[i0 + j] = run; run += v[j]; }
        __syncthreads();
        if (tid == 1023) carry_s = run;
        __syncthreads();
    }
    if (tid == 0) offs[n] = carry_s;
}

// ---------------------------------------------------------------------------------------
// Run heads of a sorted key array. One warp = one chunk of RUN_CHUNK consecutive elements.
#define RUN_CHUNK 4096
#define RUN_THREADS 256

// kshift: bits to drop before comparing (16 for packed key<<16|tag records, else 0)
template <typename KeyT>
__global__ void __launch_bounds__(RUN_THREADS)
k_run_count(const KeyT *__restrict__ keys, uint64_t n, int kshift, uint32_t *__restrict__ chunk_counts) {
    const unsigned lane = threadIdx.x & 31;
    const uint64_t chunk = (uint64_t)blockIdx.x * (RUN_THREADS / 32) + (threadIdx.x >> 5);
    const uint64_t base = chunk * RUN_CHUNK;
    if (base >= n) return;
    uint32_t c = 0;
    for (int it = 0; it < RUN_CHUNK / 32; it++) {
        const uint64_t i = base + it * 32 + lane;
        bool head = false;
        if (i < n) {
            const KeyT cur = keys[i] >> kshift;
            head = ((i == 0) || cur != (keys[i - 1] >> kshift)) && cur != (KeyT(~KeyT(0)) >> kshift);  // sentinel
        }
        c += __popc(__ballot_sync(0xffffffffu, head));
    }
    if (lane == 0) chunk_counts[chunk] = c;
}

// Per-sample run-length encode: distinct keys + head positions (counts by difference).
template <typename KeyT>
__global__ void __launch_bounds__(RUN_THREADS)
k_rle_write(const KeyT *__restrict__ keys, uint64_t n, const unsigned long long *__restrict__ chunk_offs,
            KeyT *__restrict__ out_keys, unsigned long long *__restrict__ head_pos) {
    const unsigned lane = threadIdx.x & 31;
    const uint64_t chunk = (uint64_t)blockIdx.x * (RUN_THREADS / 32) + (threadIdx.x >> 5);
    const uint64_t base = chunk * RUN_CHUNK;
    if (base >= n) return;
    unsigned long long run = chunk_offs[chunk];
    for (int it = 0; it < RUN_CHUNK / 32; it++) {
        const uint64_t i = base + it * 32 + lane;
        bool head = false;
        KeyT kk = 0;
        if (i < n) { kk = keys[i]; head = (i == 0) || kk != keys[i - 1]; }
        const unsigned ball = __ballot_sync(0xffffffffu, head);
        if (head) {
            const unsigned long long r = run + __popc(ball & lanemask_lt());
            out_keys[r] = kk;
            head_pos[r] = i;
        }
        run += __popc(ball);
    }
}

// counts[r] = head_pos[r+1] - head_pos[r]; keep flag = count >= cutoff; per-chunk kept counts
__global__ void k_rle_counts(const unsigned long long *__restrict__ head_pos, uint64_t nu, uint64_t n,
                             uint32_t cutoff, uint32_t *__restrict__ counts,
                             uint32_t *__restrict__ chunk_counts) {
    const unsigned lane = threadIdx.x & 31;
    const uint64_t chunk = (uint64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const uint64_t base = chunk * RUN_CHUNK;
    if (base >= nu) return;
    uint32_t c = 0;
    for (int it = 0; it < RUN_CHUNK / 32; it++) {
        const uint64_t r = base + it * 32 + lane;
        bool keep = false;
        if (r < nu) {
            const unsigned long long e = (r + 1 < nu) ? head_pos[r + 1] : n;
            const uint32_t cnt = (uint32_t)(e - head_pos[r]);
            counts[r] = cnt;
            keep = cnt >= cutoff;
        }
        c += __popc(__ballot_sync(0xffffffffu, keep));
    }
    if (lane == 0) chunk_counts[chunk] = c;
}

// compact (key, count) pairs with count >= cutoff, order preserved
template <typename KeyT>
__global__ void k_rle_filter(const KeyT *__restrict__ keys, const uint32_t *__restrict__ counts, uint64_t nu,
                             uint32_t cutoff, const unsigned long long *__restrict__ chunk_offs,
                             KeyT *__restrict__ out_keys, uint32_t *__restrict__ out_counts) {
    const unsigned lane = threadIdx.x & 31;
    const uint64_t chunk = (uint64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const uint64_t base = chunk * RUN_CHUNK;
    if (base >= nu) return;
    unsigned long long run = chunk_offs[chunk];
    for (int it = 0; it < RUN_CHUNK / 32; it++) {
        const uint64_t r = base + it * 32 + lane;
        bool keep = false;
        uint32_t cnt = 0;
        if (r < nu) { cnt = counts[r]; keep = cnt >= cutoff; }
        const unsigned ball = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const unsigned long long o = run + __popc(ball & lanemask_lt());
            out_keys[o] = keys[r];
            out_counts[o] = cnt;
        }
        run += __popc(ball);
    }
}
