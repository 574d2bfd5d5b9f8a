"""Rows of SURVEY.md 8f ("next"): GenomeTester4 .list files, the prediction lookup
(gmer_counter replacement) and --real_counts."""
import base64
import os

import numpy as np
import pytest

from conftest import load_json
from oracle import kmers as ok
from phenotypeseeker_b200 import glist, synth
from phenotypeseeker_b200 import prediction_gpu as pg


def test_list_writer_is_byte_identical_to_glistmaker(tmp_path):
    g = load_json("union_map.json")
    gold = base64.b64decode(g["a_5_list_b64"])
    km, ct = ok.count_kmers(base64.b64decode(g["a_fa_b64"]), 5)
    p = tmp_path / glist.sample_list_name("a", 5)
    glist.write_list(p, km, ct, 5)
    assert p.read_bytes() == gold
    km2, ct2, k2 = glist.read_list(p)
    assert k2 == 5 and np.array_equal(km2, km) and np.array_equal(ct2, ct)
    with pytest.raises(ValueError):
        glist.write_list(p, km[::-1], ct, 5)


def test_canonical_code():
    assert pg.canonical_code("TTTTT") == 0 and pg.canonical_code("AAAAA") == 0
    assert pg.canonical_code("ACGT") == ok.str_to_kmer("ACGT")
    assert pg.canonical_code("ttgca") == ok.str_to_kmer("TGCAA")


def test_gmer_counter_golden_equals_canonical_counts():
    # the shipped gmer_counter == glistmaker's canonical count for single-k-mer nodes (Appendix A7)
    for case in load_json("gmer_counter.json"):
        codes = np.array([pg.canonical_code(s) for s in case["kmers"]], dtype=np.uint64)
        for fb64, exp in zip(case["files_b64"], case["counts"]):
            km, ct = ok.count_kmers(base64.b64decode(fb64), case["k"])
            assert list(ok.map_counts(codes, km, ct)) == exp


@pytest.mark.gpu
def test_prediction_lookup_matches_gmer_counter(ctx):
    from phenotypeseeker_b200.pipeline import KmerAssociation
    ka = KmerAssociation(ctx=ctx)
    for case in load_json("gmer_counter.json"):
        files = [base64.b64decode(x) for x in case["files_b64"]]
        m, counts = pg.presence_matrix(files, case["kmers"], cutoff=1, ka=ka)
        assert np.array_equal(counts, np.array(case["counts"], dtype=np.uint32))
        assert np.array_equal(m, (np.array(case["counts"]) >= 1).astype(np.float64))
    # a repetitive sample: counts above 1 and a cutoff that matters
    rep = b">r\n" + b"ACGTTGCAAGGCTTAACCGGTTAGC" * 40 + b"\n"
    m, counts = pg.presence_matrix([rep], ["ACGTTGCAAGGCTTAAC", "TTTTTTTTTTTTTTTTT"], cutoff=5, ka=ka)
    km, ct = ok.count_kmers(rep, 17)
    assert counts[0, 0] == ok.map_counts(np.array([pg.canonical_code("ACGTTGCAAGGCTTAAC")], np.uint64), km, ct)[0] >= 5
    assert m.tolist() == [[1.0, 0.0]]


@pytest.mark.gpu
def test_real_counts_columns(ctx):
    # --real_counts (modeling.py:693-695): ML_df carries raw counts, the test still uses presence
    from phenotypeseeker_b200 import modeling_gpu as mg
    mg._STATE.update(ka=None)
    rep = b"ACGTTGCAAGGCTTAACCGGTTAGCATCGA"
    ds = synth.config(0, tiny=True, n_samples=8, genome_len=6000)
    files = [f + b">extra\n" + rep * (s % 3 + 1) + b"\n" for s, f in enumerate(ds.files)]
    U, dfs = mg.run_hot_path(files, ds.names, 16, 1, ds.pheno[:, :1], ["pheno1"], True, None, 2, 6, 1.1, True, 4,
                             real_counts=True)
    df = dfs["pheno1"]
    lists = [ok.count_kmers(f, 16) for f in files]
    for kmer in list(df.columns)[:200]:
        code = np.array([ok.str_to_kmer(kmer)], dtype=np.uint64)
        exp = [int(ok.map_counts(code, *l)[0]) for l in lists]
        assert list(df[kmer].iloc[4:]) == exp
    assert max(int(v) for v in df.iloc[4:].to_numpy().ravel()) >= 2


def test_plan_ranges_boundaries_and_ingest_range():
    """Host logic of runs in k-mer ranges (pipeline.KmerAssociation.plan_ranges): boundaries moved to top-k-mer-byte
    multiples and kept strictly ascending, ranges grouped into super-ranges, pool capacities, and the ingest
    range (= what test_in_ranges hands to ps_scatter_range for the first super-range)."""
    from phenotypeseeker_b200.pipeline import KmerAssociation

    class NoCtx:
        pass

    ka = KmerAssociation(ctx=NoCtx())
    ka.k = 16
    unit = 1 << 24
    q = [unit * 60 + 5, unit * 60 + 9, unit * 200 - 1]            # two quantiles inside one top byte
    p = ka.plan_ranges(4, q, n_instances=1_000_000, n_super=None)
    assert p["grouped"] and p["slack"] == 1.15
    assert p["splitters"] == [unit * 60, unit * 61, unit * 200]     # rounded, then pushed apart
    assert p["first_of"] == {0: 2, 2: 4}                           # 2 super-ranges of 2 ranges
    lo, hi, cap, share = p["ingest"]
    assert (lo, hi, share) == (0, unit * 61, 0.5) and cap == int(1_000_000 * 1.15 * 2 / 4) + (1 << 20)
    assert ka.plan_ranges(4, p["splitters"], 1_000_000)["splitters"] == p["splitters"]      # idempotent
    # one super-range covering everything: the ingest range is the whole space
    assert ka.plan_ranges(3, q[:2], 10, n_super=1)["ingest"][:2] == (0, 0)
    # grouping off, or a k without the paged partition: plain ranges, no ingest range, more slack
    p0 = ka.plan_ranges(4, q, 1_000_000, n_super=0)
    assert not p0["grouped"] and p0["splitters"] == q and p0["ingest"] is None and p0["slack"] == 1.3
    ka.k = 21
    assert not ka.plan_ranges(4, q, 1_000_000)["grouped"]
    ka.k = 16
    assert ka.plan_ranges(1, [], 5)["first_of"] == {} and ka.plan_ranges(1, [], 5)["ingest"] is None
