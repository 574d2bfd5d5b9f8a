/*
 * oracle/kmer_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the integer half of PhenotypeSeeker's hot path:
 * what `glistmaker` (GenomeTester4 4.2.3, shipped prebuilt as
 * /root/reference/bin/glistmaker and called from modeling.py:303-315) computes
 * for one sample: the sorted list of canonical k-mers with occurrence counts.
 *
 * GenomeTester4's source is NOT vendored in the reference; this file restates
 * its observed behaviour (SURVEY.md Appendix A + the probes recorded in
 * DESIGN.md "Reader semantics"), and is pinned against the shipped binaries by
 * oracle/make_golden.py -> tests/golden/kmer_lists.json.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this. The product path never does.
 *
 * Reader semantics restated here (all observed from the 4.2.3 binary):
 *   - bytes before the first '>' (FASTA) or '@' (FASTQ) are ignored;
 *   - FASTA: '>' ANYWHERE starts a header that runs to the next '\n';
 *     in sequence state bytes 1..31 are skipped (so LF, CR, TAB never break a
 *     window), A/C/G/T/U in either case are bases (U == T), a NUL byte ends
 *     the input, every other byte ends the current window;
 *   - k-mers never span records;
 *   - FASTQ: strict 4-line records (line 1 of every 4 is sequence); well-formed
 *     input only — the binary's behaviour on multi-line FASTQ is a parser
 *     artefact (it drops the first base of continuation lines) and is not
 *     reproduced;
 *   - canonical form = numeric min(word, reverse complement), 2 bits per
 *     base, A=0 C=1 G=2 T=3, first base most significant (Appendix A1/A2);
 *   - every occurrence counts, palindromes once per occurrence.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_BREAK 4
#define ORC_SKIP 5

/* byte -> 0..3 base, ORC_BREAK, ORC_SKIP (sequence-state classification) */
static uint8_t orc_class[256];
static int orc_class_ready = 0;

static void orc_init_class(void) {
    if (orc_class_ready) return;
    for (int b = 0; b < 256; b++) orc_class[b] = ORC_BREAK;
    for (int b = 1; b < 32; b++) orc_class[b] = ORC_SKIP;
    orc_class['A'] = orc_class['a'] = 0;
    orc_class['C'] = orc_class['c'] = 1;
    orc_class['G'] = orc_class['g'] = 2;
    orc_class['T'] = orc_class['t'] = 3;
    orc_class['U'] = orc_class['u'] = 3;
    orc_class_ready = 1;
}

/*
 * Decode a FASTA/FASTQ byte buffer into a code stream: one byte per retained
 * position, 0..3 = base, 4 = window break. Returns the number of codes.
 * `codes` must have room for n + 1 bytes. fmt_out: 0 none, 1 FASTA, 2 FASTQ.
 */
size_t orc_decode(const uint8_t *buf, size_t n, uint8_t *codes, int *fmt_out) {
    orc_init_class();
    size_t i = 0, m = 0;
    /* a NUL byte ends the input */
    const uint8_t *z = memchr(buf, 0, n);
    if (z) n = (size_t)(z - buf);
    while (i < n && buf[i] != '>' && buf[i] != '@') i++;
    if (i == n) { if (fmt_out) *fmt_out = 0; return 0; }
    if (buf[i] == '>') {
        if (fmt_out) *fmt_out = 1;
        int in_header = 0;
        for (; i < n; i++) {
            uint8_t b = buf[i];
            if (in_header) { if (b == '\n') in_header = 0; continue; }
            if (b == '>') { in_header = 1; codes[m++] = ORC_BREAK; continue; }
            uint8_t c = orc_class[b];
            if (c == ORC_SKIP) continue;
            codes[m++] = c;
        }
    } else {
        if (fmt_out) *fmt_out = 2;
        unsigned line = 0; /* line index relative to the first '@' line */
        for (; i < n; i++) {
            uint8_t b = buf[i];
            if (b == '\n') {
                if ((line & 3) == 1) codes[m++] = ORC_BREAK;
                line++;
                continue;
            }
            if ((line & 3) != 1) continue;
            uint8_t c = orc_class[b];
            if (c == ORC_SKIP) continue;
            codes[m++] = c;
        }
    }
    return m;
}

static inline uint64_t orc_revcomp(uint64_t w, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) { r = (r << 2) | (3 - (w & 3)); w >>= 2; }
    return r;
}

/* LSD radix sort of u64 keys, 16 bits per pass, only the passes that matter */
static void orc_sort_u64(uint64_t *a, size_t n, int bits) {
    if (n < 2) return;
    uint64_t *tmp = (uint64_t *)malloc(n * sizeof(uint64_t));
    size_t *cnt = (size_t *)malloc(65536 * sizeof(size_t));
    uint64_t *src = a, *dst = tmp;
    for (int shift = 0; shift < bits; shift += 16) {
        memset(cnt, 0, 65536 * sizeof(size_t));
        for (size_t i = 0; i < n; i++) cnt[(src[i] >> shift) & 0xFFFF]++;
        size_t s = 0;
        for (int d = 0; d < 65536; d++) { size_t c = cnt[d]; cnt[d] = s; s += c; }
        for (size_t i = 0; i < n; i++) dst[cnt[(src[i] >> shift) & 0xFFFF]++] = src[i];
        uint64_t *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(uint64_t));
    free(tmp); free(cnt);
}

/*
 * All canonical k-mers (with multiplicity, unsorted) of a code stream.
 * Returns the number written; out must hold >= m entries.
 */
size_t orc_kmers_from_codes(const uint8_t *codes, size_t m, int k, uint64_t *out) {
    uint64_t mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    uint64_t fw = 0, rc = 0;
    int run = 0;
    size_t o = 0;
    for (size_t i = 0; i < m; i++) {
        uint8_t c = codes[i];
        if (c > 3) { run = 0; fw = rc = 0; continue; }
        fw = ((fw << 2) | c) & mask;
        rc = (rc >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
        if (++run >= k) out[o++] = fw < rc ? fw : rc;
    }
    return o;
}

/*
 * glistmaker restated: sorted distinct canonical k-mers + counts.
 * On return *kmers / *counts are malloc'ed arrays of *nuniq entries
 * (release with orc_free), *ntotal = total k-mer occurrences.
 * Returns 0 on success.
 */
int orc_count(const uint8_t *buf, size_t n, int k,
              uint64_t **kmers, uint32_t **counts, size_t *nuniq, size_t *ntotal) {
    if (k < 1 || k > 32) return -1;
    uint8_t *codes = (uint8_t *)malloc(n + 1);
    int fmt;
    size_t m = orc_decode(buf, n, codes, &fmt);
    uint64_t *all = (uint64_t *)malloc((m ? m : 1) * sizeof(uint64_t));
    size_t t = orc_kmers_from_codes(codes, m, k, all);
    free(codes);
    orc_sort_u64(all, t, 2 * k);
    size_t u = 0;
    for (size_t i = 0; i < t; i++) if (i == 0 || all[i] != all[i - 1]) u++;
    uint64_t *ks = (uint64_t *)malloc((u ? u : 1) * sizeof(uint64_t));
    uint32_t *cs = (uint32_t *)malloc((u ? u : 1) * sizeof(uint32_t));
    size_t j = 0;
    for (size_t i = 0; i < t;) {
        size_t e = i + 1;
        while (e < t && all[e] == all[i]) e++;
        ks[j] = all[i]; cs[j] = (uint32_t)(e - i); j++;
        i = e;
    }
    free(all);
    *kmers = ks; *counts = cs; *nuniq = u; *ntotal = t;
    return 0;
}

void orc_free(void *p) { free(p); }

/*
 * glistcompare -u restated (modeling.py:374-380): sorted set union of two
 * sorted distinct lists. out must hold na + nb entries. Returns union size.
 */
size_t orc_union(const uint64_t *a, size_t na, const uint64_t *b, size_t nb, uint64_t *out) {
    size_t i = 0, j = 0, o = 0;
    while (i < na && j < nb) {
        if (a[i] < b[j]) out[o++] = a[i++];
        else if (b[j] < a[i]) out[o++] = b[j++];
        else { out[o++] = a[i]; i++; j++; }
    }
    while (i < na) out[o++] = a[i++];
    while (j < nb) out[o++] = b[j++];
    return o;
}

/*
 * glistquery A.list -l U.list restated (modeling.py:317-329): for every k-mer
 * of the union (in order) the count in the sample's list, 0 if absent.
 */
void orc_map(const uint64_t *u, size_t nu, const uint64_t *a, const uint32_t *ac, size_t na,
             uint32_t *out) {
    size_t j = 0;
    for (size_t i = 0; i < nu; i++) {
        while (j < na && a[j] < u[i]) j++;
        out[i] = (j < na && a[j] == u[i]) ? ac[j] : 0;
    }
}
