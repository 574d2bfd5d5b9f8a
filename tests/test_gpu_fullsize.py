"""Parity at BASELINE.json's full sizes (on the GPU box). Config 1 (20 x 4.3 Mbp) is small enough
for a bit-exact comparison with the oracle end to end; config 2 (250 x 4.3 Mbp) is checked through
size-independent properties: strict sortedness of the union, per-sample column popcounts equal to
the oracle's distinct-k-mer counts, every union k-mer canonical, a full re-run being bit-identical,
and range shards tiling the union."""
import numpy as np
import pytest

from oracle import kmers as ok
from oracle import stats as ostats
from phenotypeseeker_b200 import synth
from phenotypeseeker_b200.pipeline import KmerAssociation, unpack_rows

pytestmark = pytest.mark.gpu


def test_config1_full_size_bit_exact(ctx):
    ds = synth.config(0)                                    # 20 x 4.3 Mbp, seed 20260101
    ka = KmerAssociation(ctx=ctx)
    N = ds.n_samples
    res = ka.run(ds.files, 16, ds.pheno, True, None, min_samples=2, max_samples=N - 2, pvalue_cutoff=0.05,
                 omit_b=True)[0]
    lists = ok.count_many(ds.files, 16)
    u = ok.union([l[0] for l in lists])
    assert ka.U == len(u) > 5_000_000
    got_u = ctx.get_union()
    assert np.array_equal(got_u, u)
    # matrix, in slices to bound host memory
    step = 1 << 20
    for a in range(0, len(u), step):
        rows = unpack_rows(ctx.get_rows(a, min(step, len(u) - a)), N)
        exp = np.stack([ok.map_counts(u[a:a + step], *l) > 0 for l in lists], axis=1)
        assert np.array_equal(rows, exp.astype(np.uint8)), a
    pres = np.stack([ok.map_counts(u, *l) > 0 for l in lists], axis=1)
    o = ostats.chi2_rows(pres, ds.pheno[:, 0].astype(np.int8), np.ones(N), 2, N - 2)
    keep = o["tested"] & (o["p"] < 0.05)
    assert keep.sum() > 1000
    assert np.array_equal(res.row, np.nonzero(keep)[0])      # identical filtered k-mer set
    assert np.array_equal(res.stat, o["stat"][keep])         # unweighted chi2: bit-identical
    np.testing.assert_allclose(res.p, o["p"][keep], rtol=1e-13)


def _revcomp16(x):
    x = ~x & np.uint64(0xFFFFFFFF)
    out = np.zeros_like(x)
    for i in range(16):
        out |= ((x >> np.uint64(2 * i)) & np.uint64(3)) << np.uint64(2 * (15 - i))
    return out


def test_config2_full_size_properties(ctx):
    ds = synth.config(1)                                    # 250 x 4.3 Mbp, weighted
    N = ds.n_samples
    ka = KmerAssociation(ctx=ctx)
    ka.count(ds.files, 16)
    U = ka.build()
    u = ctx.get_union()
    assert U == len(u) > 20_000_000
    assert (u[1:] > u[:-1]).all()                            # strictly ascending = sorted + deduplicated
    assert (u <= _revcomp16(u)).all()                        # every union k-mer is canonical
    # per-sample column popcount == number of distinct k-mers of that sample (oracle), on a spread
    col = np.zeros(N, dtype=np.int64)
    xor = np.uint64(0)
    step = 1 << 21
    for a in range(0, U, step):
        rows = ctx.get_rows(a, min(step, U - a))
        bits = unpack_rows(rows, N)
        col += bits.sum(axis=0, dtype=np.int64)
        assert bits.any(axis=1).all()                        # no empty row: every union k-mer is in some sample
    for s in (0, 1, 77, 124, 249):
        km, _ = ok.count_kmers(ds.files[s], 16)
        assert col[s] == len(km), s
        assert np.isin(km[::997], u).all()                   # the sample's k-mers are in the union
    # stage 3 on the full matrix: thresholds nest, and the tested set obeys the min/max rule
    r_all = ka.test(ds.pheno[:, :1], True, ds.weights, max_samples=N - 2, pvalue_cutoff=0.05, omit_b=True)[0]
    r_bonf = ka.test(ds.pheno[:, :1], True, ds.weights, max_samples=N - 2, pvalue_cutoff=0.05, omit_b=False)[0]
    assert 0 < len(r_bonf.kmer) < len(r_all.kmer)
    assert np.isin(r_bonf.row, r_all.row).all()
    assert (r_bonf.p < 0.05 / U).all() and (r_all.p < 0.05).all()
    assert (r_all.n_with >= 2).all() and (r_all.n_with <= N - 2).all()
    assert np.array_equal(r_all.presence.sum(axis=1), r_all.n_with)       # no NA in this config
    np.testing.assert_allclose(r_all.p, np.exp(-r_all.stat / 2), rtol=1e-12)   # df = 2
    # a second full run is bit-identical (deterministic sort / compaction)
    ka.count(ds.files, 16)
    assert ka.build() == U
    assert np.array_equal(ctx.get_union(), u)
    r2 = ka.test(ds.pheno[:, :1], True, ds.weights, max_samples=N - 2, pvalue_cutoff=0.05, omit_b=False)[0]
    assert np.array_equal(r2.row, r_bonf.row) and np.array_equal(r2.stat, r_bonf.stat)
    # k-mer range shards tile the union
    q = ctx.sample_quantiles(0, 4)
    sizes = []
    for i in range(4):
        ctx.set_range(0 if i == 0 else q[i - 1], 0 if i == 3 else q[i])
        sizes.append(ctx.build_union())
        part = ctx.get_union(0, min(sizes[-1], 1000))
        if i > 0:
            assert part[0] >= q[i - 1]
    assert sum(sizes) == U
    assert max(sizes) < 1.6 * (U / 4)                        # quantile splitters balance the ranges


# ---------------------------------------------------------------------------------------
# Oracle comparisons at scale. Inputs come from the GPU-side generator (the bench's); the text is
# copied to the host once for the C oracle.
def _oracle_union_and_bits(texts, k, cutoff=1):
    """C oracle on host threads: per-sample lists -> union -> bit-packed presence rows like the GPU's."""
    lists = [l[0] for l in ok.count_many(texts, k, cutoff)]
    u, rows = ok.union_and_rows(lists)
    return lists, u, rows


def _oracle_stage3(rows, N, pheno_col, binary, weights, min_s, max_s, thr, slice_rows):
    """Oracle statistics of every row that can pass the sample filter (modeling.py:770-772 / :729-731),
    in row slices on host threads. A row whose sample count (over the non-NA samples) is outside
    [min_s, max_s], or that leaves fewer than 2 samples without the k-mer, is never tested, so only the
    others are unpacked and handed to oracle/stats.py. -> (surviving row indices, statistic, p)."""
    from concurrent.futures import ThreadPoolExecutor
    import os
    nonna = ~np.isnan(np.asarray(pheno_col, dtype=np.float64)) if not binary else (np.asarray(pheno_col) >= 0)
    W = rows.shape[1]
    mask = np.zeros(W, dtype=np.uint32)
    for s in np.nonzero(nonna)[0]:
        mask[s >> 5] |= np.uint32(1 << (s & 31))
    n_with = np.bitwise_count(rows & mask[None, :]).sum(axis=1, dtype=np.int64)
    n_tot = int(nonna.sum())
    cand = np.nonzero((n_with >= min_s) & (n_with <= max_s) & (n_tot - n_with >= 2))[0]

    def one(a):
        idx = cand[a:a + slice_rows]
        pres = unpack_rows(rows[idx], N)
        o = (ostats.chi2_rows(pres, pheno_col, weights, min_s, max_s) if binary
             else ostats.welch_rows(pres, pheno_col, weights, min_s, max_s))
        assert o["tested"].all()
        k = np.nonzero(o["p"] < thr)[0]
        return idx[k], o["stat"][k], o["p"][k]

    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        out = list(ex.map(one, range(0, len(cand), slice_rows)))
    if not out:
        return np.empty(0, np.int64), np.empty(0), np.empty(0)
    return np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out]), np.concatenate([o[2] for o in out])


def _render(plan_kwargs):
    import torch
    from phenotypeseeker_b200 import synth_gpu
    plan = synth_gpu.make_plan(**plan_kwargs)
    r = synth_gpu.Renderer(plan, torch.device("cuda", 0))
    dev, spans = r.render(range(plan.n_samples))
    host = dev.cpu().numpy()
    return plan, [host[o:o + n].tobytes() for s, (o, n) in sorted(spans.items())]


def test_config2_full_size_vs_oracle(ctx):
    """The bench workload of config 2 (250 x 4.3 Mbp, weighted chi2, Bonferroni) against the oracle:
    identical union, identical matrix (all 21 M rows), identical survivor set with statistics to 1e-6."""
    from phenotypeseeker_b200 import synth_gpu
    plan, texts = _render(dict(n_samples=250, genome_len=4_300_000, seed=20260102, binary=True, weighted=True,
                               pos_rate=0.35, n_clades=16))
    N = plan.n_samples
    ka = KmerAssociation(ctx=ctx)
    res = ka.run(texts, 16, plan.pheno, True, plan.weights, min_samples=2, max_samples=N - 2, pvalue_cutoff=0.05, omit_b=False)[0]
    lists, u, rows = _oracle_union_and_bits(texts, 16)
    assert ka.U == len(u) > 20_000_000
    assert np.array_equal(ctx.get_union(), u)
    step = 1 << 22
    for a in range(0, len(u), step):
        assert np.array_equal(ctx.get_rows(a, min(step, len(u) - a)), rows[a:a + step]), a
    # stage 3 on the oracle's matrix
    code = plan.pheno[:, 0].astype(np.int8)
    keep_rows, stats, ps = _oracle_stage3(rows, N, code, True, plan.weights, 2, N - 2, 0.05 / len(u), 1 << 17)
    assert len(keep_rows) > 1000
    assert np.array_equal(res.row, keep_rows)                         # identical filtered k-mer set
    np.testing.assert_allclose(res.stat, stats, rtol=1e-6)
    np.testing.assert_allclose(res.p, ps, rtol=1e-6)


def test_thousand_samples_wide_rows_vs_oracle(ctx):
    """N = 1000 (four sample groups, 128-byte rows: config 3's shape at 200 kbp genomes), continuous phenotype
    with NA: union, matrix and the Welch survivors against the oracle."""
    plan, texts = _render(dict(n_samples=1000, genome_len=200_000, seed=20260103, binary=False, na_rate=0.02, n_clades=32,
                               weighted=True))
    N = plan.n_samples
    ka = KmerAssociation(ctx=ctx)
    res = ka.run(texts, 16, plan.pheno, False, plan.weights, min_samples=2, max_samples=N - 2, pvalue_cutoff=0.05)[0]
    lists, u, rows = _oracle_union_and_bits(texts, 16)
    assert ka.U == len(u)
    assert np.array_equal(ctx.get_union(), u)
    assert np.array_equal(ctx.get_rows(), rows)
    keep_rows, stats, _ = _oracle_stage3(rows, N, plan.pheno[:, 0], False, plan.weights, 2, N - 2, 0.05 / len(u), 1 << 14)
    assert len(keep_rows) > 100
    assert np.array_equal(res.row, keep_rows)
    np.testing.assert_allclose(res.stat, stats, rtol=1e-6)


def test_deep_fastq_with_cutoff_vs_oracle(ctx):
    """Raw reads at depth (24 samples, 150 bp reads, 30x of 1.2 Mbp genomes, 1 % errors, cutoff 3): per-sample
    counted lists, union and matrix against the oracle (glistquery dump filtered by count >= cutoff)."""
    plan, texts = _render(dict(n_samples=24, genome_len=1_200_000, seed=20260104, binary=True, reads=True, coverage=30.0))
    ka = KmerAssociation(ctx=ctx)
    ka.count(texts, 16, cutoff=3)
    U = ka.build()
    lists, u, rows = _oracle_union_and_bits(texts, 16, cutoff=3)
    for s in (0, 11, 23):
        km, ct = ctx.sample_kmers(s, 3)
        okm, oct_ = ok.count_kmers(texts[s], 16, 3)
        assert np.array_equal(km, okm) and np.array_equal(ct, oct_)
    assert U == len(u) > 1_000_000
    assert np.array_equal(ctx.get_union(), u)
    assert np.array_equal(ctx.get_rows(), rows)
