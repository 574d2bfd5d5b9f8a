// Host check of phenotypeseeker_b200/csrc/ps_decode_bits.h (the word-wide FASTA emit of k_decode_write_fasta)
// against a byte-at-a-time transducer + straightforward packing. Built and run by tests/test_decode_bits.py.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../phenotypeseeker_b200/csrc/ps_decode_bits.h"

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

static std::vector<uint8_t> make_text(size_t len, int flavour) {
    std::vector<uint8_t> t;
    const char *bases = "ACGT";
    size_t col = 0;
    t.push_back('>'); t.push_back('h'); t.push_back('\n');
    while (t.size() < len) {
        uint32_t r = rnd() % 1000;
        if (flavour >= 1 && r < 3) { const char *h = "\n>contig_12 len=5\n"; for (const char *p = h; *p; p++) t.push_back(*p); col = 0; continue; }
        if (flavour >= 2 && r < 30) { const char odd[] = {'N', 'n', 'a', 'c', 'g', 't', 'u', 'U', 'R', '-', '*', ' ', '\r', '\t', (char)0x80, (char)0xC1, '>', 'x'}; t.push_back((uint8_t)odd[rnd() % sizeof(odd)]); col++; continue; }
        t.push_back(bases[rnd() & 3]); col++;
        if (col >= 60) { t.push_back('\n'); col = 0; }
    }
    t.resize(len);
    return t;
}

int main() {
    int bad = 0;
    for (int trial = 0; trial < 300; trial++) {
        const size_t len = 1 + rnd() % 5000;
        const int flavour = trial % 3;
        std::vector<uint8_t> text = make_text(len, flavour);
        const size_t start = (trial % 5 == 0) ? rnd() % (len < 40 ? len : 40) : 0;     // bytes before `start` are ignored
        const uint32_t lead = rnd() % 32;                                               // output begins `lead` positions in
        // reference: byte transducer -> codes
        std::vector<uint8_t> codes;
        uint32_t s = 0;
        for (size_t i = start; i < len; i++) {
            const uint32_t b = text[i];
            if (s == 1) { if (b == '\n') s = 0; continue; }
            if (b == '>') { codes.push_back(PSD_CODE_BREAK); s = 1; continue; }
            const uint32_t c = psd_classify(b);
            if (c != PSD_CODE_SKIP) codes.push_back((uint8_t)c);
        }
        const size_t npos = lead + codes.size();
        std::vector<uint32_t> rseq(npos / 16 + 8, 0), rbad(npos / 32 + 8, 0);
        for (size_t i = 0; i < codes.size(); i++) {
            const size_t p = lead + i;
            rseq[p >> 4] |= (uint32_t)(codes[i] & 3) << (30 - 2 * (p & 15));
            rbad[p >> 5] |= (uint32_t)(codes[i] >> 2) << (p & 31);
        }
        // new: chunk by chunk
        std::vector<uint32_t> nseq(rseq.size(), 0), nbad(rbad.size(), 0);
        uint32_t state = 0;
        size_t pos = lead;
        for (size_t base = 0; base < len; base += 64) {
            uint32_t w[16];
            uint8_t buf[64];
            memset(buf, 0, 64);
            memcpy(buf, text.data() + base, len - base < 64 ? len - base : 64);
            memcpy(w, buf, 64);
            const int jlo = start > base ? (int)(start - base < 64 ? start - base : 64) : 0;
            const int jhi = (int)(len - base < 64 ? len - base : 64);
            PsdBits o = {0, 0, 0, 0, state};
            psd_fasta_chunk(w, jlo, jhi, o);
            psd_place(o, (uint32_t)pos, nseq.data(), nbad.data(), [](uint32_t *p, uint32_t v) { *p |= v; });
            pos += o.n;
            state = o.s;
        }
        if (pos != npos || nseq != rseq || nbad != rbad) {
            bad++;
            fprintf(stderr, "trial %d: len %zu start %zu lead %u: positions %zu vs %zu, seq %s, bad %s\n", trial, len, start, lead,
                    pos, npos, nseq == rseq ? "ok" : "DIFFER", nbad == rbad ? "ok" : "DIFFER");
        }
    }
    printf("%s: %d failing trials of 300\n", bad ? "FAIL" : "OK", bad);
    return bad ? 1 : 0;
}
