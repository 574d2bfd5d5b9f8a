"""Synthetic inputs of the shapes BASELINE.json names.

Fixed-seed generator for C. difficile-shaped assemblies (29 % GC, 4.3 / 5 Mbp,
50-100 contigs, 60-column FASTA), a clade + private-SNP population with
phenotype-linked accessory cassettes, binary / continuous phenotypes with NA,
synthetic Gamma(2) sample weights (mean 1), and 150 bp raw-read FASTQ.

There is no network in this project, so every benchmark and parity test runs
on these; `data: "synthetic"` in bench.py's JSON line refers to this module.
"""
from dataclasses import dataclass, field
import os

import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_BASE_P = np.array([0.355, 0.145, 0.145, 0.355])  # A C G T  -> 29 % GC


@dataclass
class Dataset:
    names: list
    files: list                    # bytes per sample (FASTA or FASTQ text)
    pheno_names: list
    pheno: np.ndarray              # N x P float64; NaN = NA; binary columns hold 0/1
    binary: bool
    weights: np.ndarray            # N float64 (all 1.0 when unweighted)
    k: int = 16
    meta: dict = field(default_factory=dict)

    @property
    def n_samples(self):
        return len(self.names)

    def total_bytes(self):
        return sum(len(f) for f in self.files)

    def write(self, outdir, suffix=None):
        """Write sample files + data.pheno (reference input format, modeling.py:74-97)."""
        os.makedirs(outdir, exist_ok=True)
        suffix = suffix or (".fq" if self.meta.get("reads") else ".fa")
        paths = []
        for name, data in zip(self.names, self.files):
            p = os.path.join(outdir, name + suffix)
            with open(p, "wb") as f:
                f.write(data)
            paths.append(p)
        ph = os.path.join(outdir, "data.pheno")
        with open(ph, "w") as f:
            f.write("ID\tAddress\t" + "\t".join(self.pheno_names) + "\n")
            for i, (name, p) in enumerate(zip(self.names, paths)):
                vals = []
                for j in range(self.pheno.shape[1]):
                    v = self.pheno[i, j]
                    if np.isnan(v):
                        vals.append("NA")
                    elif self.binary:
                        vals.append(str(int(v)))
                    else:
                        vals.append(repr(float(v)))
                f.write(f"{name}\t{p}\t" + "\t".join(vals) + "\n")
        return ph, paths


def _mutate(rng, genome, rate):
    n = rng.binomial(len(genome), rate)
    if n == 0:
        return genome.copy()
    pos = rng.integers(0, len(genome), size=n)
    g = genome.copy()
    g[pos] = (g[pos] + rng.integers(1, 4, size=n).astype(np.uint8)) & 3
    return g


def _fasta_bytes(rng, name, genome, extra_contigs, n_contigs, width=60):
    cuts = np.sort(rng.choice(np.arange(1000, len(genome) - 1000), size=n_contigs - 1, replace=False)) \
        if len(genome) > 2000 + n_contigs else np.array([], dtype=np.int64)
    bounds = np.concatenate([[0], cuts, [len(genome)]])
    parts = []
    contigs = [genome[bounds[i]:bounds[i + 1]] for i in range(len(bounds) - 1)] + list(extra_contigs)
    for ci, c in enumerate(contigs):
        parts.append(f">{name}_c{ci + 1}\n".encode())
        n = len(c)
        rows = (n + width - 1) // width
        buf = np.full((rows, width + 1), ord("\n"), dtype=np.uint8)
        flat = np.zeros(rows * width, dtype=np.uint8)
        flat[:n] = _ASCII[c]
        buf[:, :width] = flat.reshape(rows, width)
        out = buf.reshape(-1)
        # drop the padding of the last row but keep its newline
        tail_pad = rows * width - n
        if tail_pad:
            out = np.concatenate([out[:len(out) - 1 - tail_pad], out[-1:]])
        parts.append(out.tobytes())
    return b"".join(parts)


def _fastq_bytes(rng, name, genome, coverage, read_len=150, err=0.01, n_rate=0.001):
    L = len(genome)
    n_reads = max(1, int(L * coverage / read_len))
    starts = rng.integers(0, L - read_len + 1, size=n_reads)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    reads = genome[idx]
    rev = rng.random(n_reads) < 0.5
    reads[rev] = (3 - reads[rev])[:, ::-1]
    e = rng.random(reads.shape) < err
    reads[e] = (reads[e] + rng.integers(1, 4, size=int(e.sum())).astype(np.uint8)) & 3
    asc = _ASCII[reads]
    asc[rng.random(reads.shape) < n_rate] = ord("N")
    qual = b"I" * read_len
    nm = name.encode()
    # 4-line records; every 7th read gets a '+name' line like real files do
    out = bytearray()
    for i in range(n_reads):
        out += b"@" + nm + b"_r%d\n" % i
        out += asc[i].tobytes() + b"\n"
        out += (b"+" + nm + b"_r%d\n" % i) if i % 7 == 3 else b"+\n"
        out += qual + b"\n"
    return bytes(out)


def make_dataset(n_samples, genome_len=4_300_000, seed=20260101, k=16, binary=True,
                 n_pheno=1, n_clades=8, clade_snp=3e-3, private_snp=1e-3, n_cassettes=6,
                 cassette_len=(2000, 10000), contigs=(50, 100), na_rate=0.0,
                 pos_rate=0.5, weighted=False, reads=False, coverage=30.0,
                 prefix="s"):
    """Generate one population."""
    rng = np.random.default_rng(seed)
    ancestor = rng.choice(4, size=genome_len, p=_BASE_P).astype(np.uint8)
    n_clades = max(1, min(n_clades, n_samples))
    clades = [_mutate(rng, ancestor, clade_snp) for _ in range(n_clades)]
    cassettes = [rng.choice(4, size=int(rng.integers(cassette_len[0], cassette_len[1] + 1)),
                            p=_BASE_P).astype(np.uint8) for _ in range(n_cassettes)]
    if genome_len < 200_000:  # tiny CI scale: keep cassettes proportionate
        cassettes = [c[:max(300, genome_len // 40)] for c in cassettes]

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
            flip = rng.random(n_samples) < 0.10      # 90 % concordance
            pheno[:, j] = np.where(flip, ~driver, driver).astype(np.float64)
        else:
            pheno[:, j] = np.round(-1.0 + 4.0 * driver + rng.normal(0, 1, n_samples), 3)
        if na_rate > 0:
            pheno[rng.random(n_samples) < na_rate, j] = np.nan
    # cassettes not tied to a phenotype are present at random
    for cas in range(n_cassettes):
        if not has_cas[:, cas].any():
            has_cas[:, cas] = rng.random(n_samples) < 0.3

    names, files = [], []
    for s in range(n_samples):
        name = f"{prefix}{s:04d}"
        g = _mutate(rng, clades[clade_of[s]], private_snp)
        extra = [cassettes[c] for c in range(n_cassettes) if has_cas[s, c]]
        if reads:
            full = np.concatenate([g] + extra) if extra else g
            data = _fastq_bytes(rng, name, full, coverage)
        else:
            nc = int(rng.integers(contigs[0], contigs[1] + 1))
            nc = max(1, min(nc, genome_len // 5000))
            data = _fasta_bytes(rng, name, g, extra, nc)
        names.append(name)
        files.append(data)
    if weighted:
        w = rng.gamma(2.0, 1.0, size=n_samples)
        w = w / w.mean()
    else:
        w = np.ones(n_samples)
    return Dataset(names=names, files=files,
                   pheno_names=[f"pheno{j + 1}" for j in range(n_pheno)],
                   pheno=pheno, binary=binary, weights=w, k=k,
                   meta={"genome_len": genome_len, "seed": seed, "reads": reads,
                         "n_clades": n_clades, "weighted": weighted})


# The five BASELINE.json configs, at full scale or a CI scale (L = 60 kbp, N <= 12).
def config(idx, tiny=False, n_samples=None, genome_len=None):
    full = {
        0: dict(n_samples=20, genome_len=4_300_000, seed=20260101, binary=True),
        1: dict(n_samples=250, genome_len=4_300_000, seed=20260102, binary=True, weighted=True,
                pos_rate=0.35, n_clades=16),
        2: dict(n_samples=1000, genome_len=5_000_000, seed=20260103, binary=False, na_rate=0.02,
                n_clades=32),
        3: dict(n_samples=200, genome_len=4_300_000, seed=20260104, binary=True, reads=True),
        4: dict(n_samples=5000, genome_len=5_000_000, seed=20260105, binary=True, n_pheno=10,
                n_clades=64, n_cassettes=12),
    }[idx]
    if tiny:
        full["n_samples"] = min(12, full["n_samples"])
        full["genome_len"] = 60_000
        full["n_clades"] = min(full.get("n_clades", 8), 4)
        if full.get("reads"):
            full["n_samples"] = 4
            full["genome_len"] = 20_000
    if n_samples is not None:
        full["n_samples"] = n_samples
    if genome_len is not None:
        full["genome_len"] = genome_len
    return make_dataset(**full)
