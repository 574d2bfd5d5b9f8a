"""k-mer lookup of `phenotypeseeker prediction` on the GPU.

The reference writes the model's k-mers as a `gmer_counter` text database
(prediction.py:145-148), runs `gmer_counter -db` per sample (:72-80), thresholds the counts at
`-c` (:82-100) and zips the per-sample text files into an N x K matrix (:150-163). Here the
samples are decoded once and `ps_lookup` (k_lookup: extraction + binary search in the sorted
query list) returns the occurrence counts of the K k-mers; single-k-mer gmer_counter nodes are
strand-agnostic and equal glistmaker's canonical counts, so a query
k-mer is canonicalised first.

`install(prediction_module)` re-wires Samples.map_samples / Samples.kmer_counts /
Phenotypes.set_kmer_db / Phenotypes.get_inp_matrix so that the unmodified `prediction()` driver,
model loading and `predict()` keep working.
"""
import os

import numpy as np

from .pipeline import KmerAssociation, read_sample_file

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def canonical_code(kmer: str) -> int:
    """k-mer string -> canonical 2-bit code (A0 C1 G2 T3, first base most significant)."""
    s = kmer.upper().replace("U", "T")
    rc = "".join(_COMP[c] for c in reversed(s))
    v = lambda t: int("".join(str("ACGT".index(c)) for c in t), 4)
    return min(v(s), v(rc))


def presence_matrix(buffers, kmers, cutoff=1, device=0, ka=None):
    """N x K float matrix of (count >= cutoff) for N samples (raw FASTA/FASTQ bytes) and K
    k-mer strings — the matrix Phenotypes.get_inp_matrix builds (prediction.py:150-159).
    Returns (matrix, counts)."""
    kmers = list(kmers)
    if not kmers:
        return np.zeros((len(buffers), 0)), np.zeros((len(buffers), 0), dtype=np.uint32)
    k = len(kmers[0])
    codes = np.array([canonical_code(x) for x in kmers], dtype=np.uint64)
    ka = ka or KmerAssociation(device=device)
    ka.ctx.begin(k, len(buffers), 1)
    ka.ctx.add_samples(0, list(buffers))
    counts = np.stack([ka.ctx.lookup(s, codes) for s in range(len(buffers))], axis=0)
    return (counts >= int(cutoff)).astype(np.float64), counts


def install(p, device=None):
    """Re-wire the reference module `p` (PhenotypeSeeker.prediction)."""
    state = {}

    def map_samples(self, pheno):          # Pool worker: nothing to do
        return None

    def kmer_counts(self, pheno):          # Pool worker: nothing to do
        return None

    def set_kmer_db(self):
        return None

    def get_inp_matrix(self):
        dev = int(os.environ.get("PS_DEVICE", "0")) if device is None else device
        if "bufs" not in state:
            state["bufs"] = [read_sample_file(s.address) for s in p.Input.samples.values()]
        m, _ = presence_matrix(state["bufs"], [str(x) for x in self.kmers], p.Phenotypes.cutoff, dev)
        self.matrix = m
        if self.pca:
            self.scaled_matrix = self.scaler.transform(self.matrix)
            self.matrix = self.pca_model.transform(self.scaled_matrix)
            self.matrix = self.matrix[:, self.PCs_to_keep]

    p.Samples.map_samples = map_samples
    p.Samples.kmer_counts = kmer_counts
    p.Phenotypes.set_kmer_db = set_kmer_db
    p.Phenotypes.get_inp_matrix = get_inp_matrix
    return p
