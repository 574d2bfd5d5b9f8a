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
    lists = [ok.count_kmers(f, 16) for f in ds.files]
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
