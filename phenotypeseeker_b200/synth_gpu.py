"""Synthetic inputs of the BASELINE.json shapes rendered ON THE GPU (torch), for bench.py and the
full-size tests: 5,000 x 5 Mbp assemblies are 25 GB of FASTA text, which the numpy generator in
synth.py would take ~15 minutes to produce; here every sample is a handful of device kernels and the
text never exists as Python bytes (it is written straight into one device buffer, and copied to
pinned host memory for the end-to-end measurement).

Same population model as synth.make_dataset: 29 % GC ancestor, clades (SNP rate
3e-3), private SNPs (1e-3), accessory cassettes tied to phenotypes, 50-100 contigs, 60-column FASTA;
raw reads: 150 bp, both strands, 1 % substitutions, 0.1 % N, 4-line FASTQ with fixed-width names.
Every sample is generated from its own seed, so a rank that renders only its own block of samples
produces exactly the bytes a single process would. torch is used for data generation only.
"""
from dataclasses import dataclass, field

import numpy as np

_BASE_P = np.array([0.355, 0.145, 0.145, 0.355])


@dataclass
class Plan:
    """Everything about a population except the per-sample text (host side, cheap)."""
    n_samples: int
    genome_len: int
    seed: int
    binary: bool
    names: list
    pheno: np.ndarray              # N x P float64, NaN = NA
    weights: np.ndarray            # N float64
    clade_of: np.ndarray
    has_cas: np.ndarray            # N x n_cassettes bool
    n_clades: int
    n_cassettes: int
    cassette_len: tuple
    reads: bool = False
    coverage: float = 30.0
    clade_snp: float = 3e-3
    private_snp: float = 1e-3
    contigs: tuple = (50, 100)
    meta: dict = field(default_factory=dict)


def make_plan(n_samples, genome_len, seed, binary=True, n_pheno=1, n_clades=8, n_cassettes=6,
              cassette_len=(2000, 10000), na_rate=0.0, pos_rate=0.5, weighted=False, reads=False, coverage=30.0):
    rng = np.random.default_rng(seed)
    n_clades = max(1, min(n_clades, n_samples))
    clade_of = rng.integers(0, n_clades, size=n_samples)
    pheno = np.zeros((n_samples, n_pheno), dtype=np.float64)
    has_cas = np.zeros((n_samples, n_cassettes), dtype=bool)
    for j in range(n_pheno):
        kind = j % 3  # cassette-linked, clade-linked, random
        if kind == 0:
            cas = j % n_cassettes
            driver = rng.random(n_samples) < pos_rate
            has_cas[:, cas] = driver
        elif kind == 1:
            driver = np.isin(clade_of, rng.choice(n_clades, size=max(1, n_clades // 3), replace=False))
        else:
            driver = rng.random(n_samples) < pos_rate
        if binary:
            flip = rng.random(n_samples) < 0.10
            pheno[:, j] = np.where(flip, ~driver, driver).astype(np.float64)
        else:
            pheno[:, j] = np.round(-1.0 + 4.0 * driver + rng.normal(0, 1, n_samples), 3)
        if na_rate > 0:
            pheno[rng.random(n_samples) < na_rate, j] = np.nan
    for cas in range(n_cassettes):
        if not has_cas[:, cas].any():
            has_cas[:, cas] = rng.random(n_samples) < 0.3
    if weighted:
        w = rng.gamma(2.0, 1.0, size=n_samples)
        w = w / w.mean()
    else:
        w = np.ones(n_samples)
    if genome_len < 200_000:
        cassette_len = (max(300, genome_len // 40), max(300, genome_len // 40))
    return Plan(n_samples=n_samples, genome_len=genome_len, seed=seed, binary=binary,
                names=[f"s{s:04d}" for s in range(n_samples)], pheno=pheno, weights=w, clade_of=clade_of,
                has_cas=has_cas, n_clades=n_clades, n_cassettes=n_cassettes, cassette_len=cassette_len,
                reads=reads, coverage=coverage)


# the five BASELINE.json configs (index 1..5 = configs[0..4])
def config_plan(idx, n_samples=None, genome_len=None):
    full = {
        1: dict(n_samples=20, genome_len=4_300_000, seed=20260101, binary=True),
        2: dict(n_samples=250, genome_len=4_300_000, seed=20260102, binary=True, weighted=True, pos_rate=0.35, n_clades=16),
        3: dict(n_samples=1000, genome_len=5_000_000, seed=20260103, binary=False, na_rate=0.02, n_clades=32),
        4: dict(n_samples=200, genome_len=4_300_000, seed=20260104, binary=True, reads=True),
        5: dict(n_samples=5000, genome_len=5_000_000, seed=20260105, binary=True, n_pheno=10, n_clades=64, n_cassettes=12),
    }[idx]
    if n_samples is not None:
        full["n_samples"] = n_samples
    if genome_len is not None:
        full["genome_len"] = genome_len
    return make_plan(**full)


class Renderer:
    """Renders sample text on one GPU from a Plan. Ancestor, clades and cassettes are built once
    (deterministically from plan.seed), each sample from seed (plan.seed, sample index)."""

    def __init__(self, plan: Plan, device):
        import torch
        self.torch = torch
        self.plan = plan
        self.device = device
        g = torch.Generator(device=device)
        g.manual_seed(plan.seed)
        probs = torch.tensor(_BASE_P, dtype=torch.float32, device=device)
        L = plan.genome_len
        self.ancestor = torch.multinomial(probs, L, replacement=True, generator=g).to(torch.uint8)
        self.clades = [self._mutate(self.ancestor, plan.clade_snp, g) for _ in range(plan.n_clades)]
        lens = np.random.default_rng(plan.seed + 1).integers(plan.cassette_len[0], plan.cassette_len[1] + 1,
                                                             size=plan.n_cassettes)
        self.cassettes = [torch.multinomial(probs, int(n), replacement=True, generator=g).to(torch.uint8) for n in lens]
        self.ascii = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)

    def _mutate(self, genome, rate, g):
        torch = self.torch
        n = int(np.random.default_rng(int(g.initial_seed()) & 0x7FFFFFFF).binomial(len(genome), rate)) if rate > 0 else 0
        out = genome.clone()
        if n:
            # the substitution at a position is a function of the position, so duplicate draws agree
            # (index_put with duplicates is otherwise order-dependent) and every rank renders the same bytes
            pos = torch.randint(0, len(genome), (n,), device=self.device, generator=g)
            delta = ((pos * 2654435761) >> 7) % 3 + 1
            out[pos] = (genome[pos] + delta.to(torch.uint8)) & 3
        return out

    def _sample_genome(self, s):
        torch = self.torch
        plan = self.plan
        g = torch.Generator(device=self.device)
        g.manual_seed((plan.seed * 1_000_003 + s) & 0x7FFFFFFFFFFF)
        genome = self._mutate(self.clades[int(plan.clade_of[s])], plan.private_snp, g)
        extra = [self.cassettes[c] for c in range(plan.n_cassettes) if plan.has_cas[s, c]]
        return genome, extra, g

    def max_text_bytes(self, s):
        """Upper bound of the text size of sample s (to size the destination buffer)."""
        plan = self.plan
        extra = int(sum(len(self.cassettes[c]) for c in range(plan.n_cassettes) if plan.has_cas[s, c]))
        L = plan.genome_len + extra
        if plan.reads:
            n_reads = max(1, int(L * plan.coverage / 150))
            return n_reads * 320
        n_contigs = plan.contigs[1] + plan.n_cassettes
        return L + L // 60 + n_contigs * 24 + 64

    def render_fasta(self, s, out, off):
        """Writes the FASTA text of sample s into out[off:] (uint8 device tensor); returns its length."""
        torch = self.torch
        plan = self.plan
        genome, extra, g = self._sample_genome(s)
        rs = np.random.default_rng((plan.seed << 20) + s)
        nc = int(rs.integers(plan.contigs[0], plan.contigs[1] + 1))
        nc = max(1, min(nc, plan.genome_len // 5000))
        L = plan.genome_len
        cuts = np.sort(rs.choice(np.arange(1000, L - 1000), size=nc - 1, replace=False)) if (nc > 1 and L > 2000 + nc) else np.array([], dtype=np.int64)
        bounds = np.concatenate([[0], cuts, [L]]).astype(np.int64)
        seq = torch.cat([genome] + extra) if extra else genome
        clen = np.diff(bounds).tolist() + [len(e) for e in extra]         # contig lengths
        cstart = np.concatenate([[0], np.cumsum(clen)]).astype(np.int64)   # in sequence coordinates
        name = plan.names[s].encode()
        headers = [b">" + name + b"_c%d\n" % (i + 1) for i in range(len(clen))]
        hlen = np.array([len(h) for h in headers], dtype=np.int64)
        body = np.array([n + (n + 59) // 60 for n in clen], dtype=np.int64)     # bases + one newline per line
        tstart = np.concatenate([[0], np.cumsum(hlen + body)]).astype(np.int64)  # text offset of each contig's header
        total = int(tstart[-1])
        view = out[off:off + total]
        view.fill_(10)                                                            # '\n' everywhere, then headers and bases
        cstart_t = torch.from_numpy(cstart).to(self.device)
        base0_t = torch.from_numpy(tstart[:-1] + hlen).to(self.device)           # text offset of each contig's first base
        idx = torch.arange(len(seq), device=self.device)
        cid = torch.bucketize(idx, cstart_t[1:], right=True)
        o = idx - cstart_t[cid]
        view[base0_t[cid] + o + o // 60] = self.ascii[seq.long()]
        hb = np.frombuffer(b"".join(headers), dtype=np.uint8)
        hpos = np.concatenate([np.arange(int(tstart[i]), int(tstart[i]) + int(hlen[i])) for i in range(len(clen))])
        view[torch.from_numpy(hpos).to(self.device)] = torch.from_numpy(hb.copy()).to(self.device)
        return total

    def render_fastq(self, s, out, off):
        """4-line FASTQ, fixed-width record: '@sNNNN_rNNNNNNN\\n' (16) + 150 bases + '\\n+\\n' + 150 x 'I' + '\\n' = 320 bytes."""
        torch = self.torch
        plan = self.plan
        genome, extra, g = self._sample_genome(s)
        full = torch.cat([genome] + extra) if extra else genome
        L = len(full)
        RL = 150
        n_reads = max(1, int(L * plan.coverage / RL))
        total = n_reads * 320
        view = out[off:off + total].view(n_reads, 320)
        starts = torch.randint(0, L - RL + 1, (n_reads,), device=self.device, generator=g)
        idx = starts[:, None] + torch.arange(RL, device=self.device)[None, :]
        reads = full[idx]
        rev = torch.rand(n_reads, device=self.device, generator=g) < 0.5
        reads = torch.where(rev[:, None], (3 - reads).flip(1), reads)
        err = torch.rand(reads.shape, device=self.device, generator=g) < 0.01
        reads = torch.where(err, (reads + torch.randint(1, 4, reads.shape, device=self.device, generator=g).to(torch.uint8)) & 3, reads)
        asc = self.ascii[reads.long()]
        asc[torch.rand(reads.shape, device=self.device, generator=g) < 0.001] = ord("N")
        hdr = np.frombuffer(b"@" + plan.names[s].encode() + b"_r", dtype=np.uint8)          # 8 bytes: @s0000_r
        view[:, :len(hdr)] = torch.from_numpy(hdr.copy()).to(self.device)[None, :]
        num = torch.arange(n_reads, device=self.device)
        for d in range(7):
            view[:, len(hdr) + d] = ((num // (10 ** (6 - d))) % 10 + 48).to(torch.uint8)
        view[:, 15] = 10
        view[:, 16:16 + RL] = asc
        view[:, 166] = 10
        view[:, 167] = ord("+")
        view[:, 168] = 10
        view[:, 169:169 + RL] = ord("I")
        view[:, 319] = 10
        return total

    def render(self, samples):
        """Text of the given samples back to back (64-byte aligned) in ONE device tensor.
        -> (uint8 device tensor, {sample: (offset, length)})."""
        torch = self.torch
        cap = sum((self.max_text_bytes(s) + 127) // 64 * 64 for s in samples) + 64
        out = torch.empty(cap, dtype=torch.uint8, device=self.device)
        spans, off = {}, 0
        for s in samples:
            n = self.render_fastq(s, out, off) if self.plan.reads else self.render_fasta(s, out, off)
            spans[s] = (off, n)
            off += (n + 63) // 64 * 64
        return out[:max(off, 64)], spans
