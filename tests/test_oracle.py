"""CPU tests: the oracle (oracle/) against the golden vectors generated from the REAL reference
(shipped GenomeTester4 binaries + unmodified modeling.py methods), and against the known-answer
rows printed in the reference's own docs."""
import base64
import math

import numpy as np
import pytest

from conftest import load_json, stage3_inputs
from oracle import kmers as ok
from oracle import stats as ostats


def test_oracle_kmer_lists_match_glistmaker(golden_kmer_lists):
    assert len(golden_kmer_lists) >= 30
    for c in golden_kmer_lists:
        km, ct = ok.count_kmers(c["data"], c["k"])
        assert np.array_equal(km, c["kmers"]), (c["name"], c["k"])
        assert np.array_equal(ct, c["counts"]), (c["name"], c["k"])


def test_appendix_a_known_answers():
    # SURVEY.md Appendix A3 (observed from the shipped binary): a.fa k=5 -> 32 distinct / 70 total
    a = b">c1 test\nACGTTGCAAGGCTTAACCGGTTNACGTACGTAGGCTAGCTAGGATCC\nacgttgcaaggcttaa\n>c2\nTTTTTTTTTTTTTTTTTTTT\n"
    km, ct = ok.count_kmers(a, 5)
    assert len(km) == 32 and ct.sum() == 70
    d = {ok.kmer_to_str(x, 5): int(c) for x, c in zip(km, ct)}
    assert d["AAAAA"] == 16 and d["TGCAA"] == 4 and d["ACGTG"] == 1 and d["GTGGA"] == 1
    b = b">x\nACGTACGTRACGTTTGGCCAA-ACGT*ACGTAC\n"
    km, ct = ok.count_kmers(b, 4)
    assert {ok.kmer_to_str(x, 4): int(c) for x, c in zip(km, ct)} == {
        "AAAC": 1, "AACG": 1, "ACGT": 5, "CAAA": 1, "CCAA": 2, "CGTA": 3, "GCCA": 2, "GGCC": 1, "GTAC": 2}


def test_union_and_map_match_glistcompare_glistquery():
    g = load_json("union_map.json")
    a = ok.count_kmers(base64.b64decode(g["a_fa_b64"]), 5)
    r = ok.count_kmers(base64.b64decode(g["r_fq_b64"]), 5)
    u = ok.union([a[0], r[0]])
    exp_u = [ln.split("\t")[0] for ln in g["union"].splitlines()]
    assert [ok.kmer_to_str(x, 5) for x in u] == exp_u
    for key, lst in (("map_a", a), ("map_r", r)):
        exp = [int(ln.split("\t")[1]) for ln in g[key].splitlines()]
        assert list(ok.map_counts(u, *lst)) == exp
    # .list binary layout (Appendix A1): 40-byte header + 12-byte records, ascending
    raw = base64.b64decode(g["a_5_list_b64"])
    assert raw[:4] == b"C4TG" and len(raw) == 40 + 12 * len(a[0])
    rec = np.frombuffer(raw[40:], dtype=np.dtype([("w", "<u8"), ("c", "<u4")]))
    assert np.array_equal(rec["w"], a[0]) and np.array_equal(rec["c"], a[1])


def test_cutoff_semantics():
    r = base64.b64decode(load_json("union_map.json")["r_fq_b64"])
    km1, ct1 = ok.count_kmers(r, 5, cutoff=1)
    km3, ct3 = ok.count_kmers(r, 5, cutoff=3)
    assert np.array_equal(km3, km1[ct1 >= 3]) and (ct3 >= 3).all() and len(km3) < len(km1)


def test_doc_rows_pin_df2():
    # README.md:100-103 and user_manual.md:62-69: every (chi2, p) pair satisfies p = exp(-chi2/2)
    rows = [(6.05, "4.86E-02"), (19.46, "5.96E-05"), (10.83, "4.44E-03"), (22.43, "1.35E-05"),
            (0.18, "9.13E-01"), (2.42, "2.98E-01"), (4.73, "9.39E-02"), (3.96, "1.38E-01"),
            (0.02, "9.92E-01"), (0.07, "9.67E-01")]
    from scipy import stats
    for chi2, ptxt in rows:
        lo, hi = math.exp(-(chi2 + 0.005) / 2), math.exp(-(chi2 - 0.005) / 2)
        assert lo * 0.995 <= float(ptxt) <= hi * 1.005
        assert abs(stats.chi2.sf(chi2, 2) - math.exp(-chi2 / 2)) < 1e-15


def _names(case, pres_row, ph):
    out = []
    for i, v in enumerate(pres_row):
        na = (ph[i] == -1) if case["kind"] == "chi2" else math.isnan(ph[i])
        if v and not na:
            out.append(f"s{i}")
    return " ".join(["|"] + out)


def test_literal_restatement_matches_real_reference_methods(golden_stage3):
    n_rows = n_kept = 0
    for case in golden_stage3:
        pres, ph, w = stage3_inputs(case)
        weights = list(w) if not np.all(w == 1.0) else [1] * len(w)
        for pv, exp in zip(pres, case["rows"]):
            n_rows += 1
            if case["kind"] == "chi2":
                r = ostats.chi2_row(pv, ph, weights, case["min"], case["max"])
                keep = r is not None and ostats.passes(r[1], case["cutoff"], case["U"], case["omit_B"], True)
                if not keep:
                    assert exp is None
                    continue
                assert exp is not None
                assert exp[1] == round(r[0], 2) and exp[2] == "%.2E" % r[1] and exp[3] == r[2]
                assert exp[4] == _names(case, pv, ph) and exp[5:] == [int(x) for x in pv]
            else:
                phl = [None if math.isnan(x) else float(x) for x in ph]
                r = ostats.welch_row(pv, phl, weights, case["min"], case["max"])
                keep = r is not None and ostats.passes(r[1], case["cutoff"], case["U"], False, False)
                if not keep:
                    assert exp is None
                    continue
                assert exp is not None
                assert exp[1] == round(r[0], 2) and exp[2] == "%.2E" % r[1]
                assert exp[3] == round(r[2], 2) and exp[4] == round(r[3], 2) and exp[5] == r[4]
                assert exp[6] == _names(case, pv, ph)
            n_kept += 1
    assert n_rows > 1000 and n_kept > 100


def test_appendix_kats(golden_stage3):
    # Appendix A9/A10/A11 rows, produced by the real conduct_* methods
    a9, a10, a11w, a11u = golden_stage3[0], golden_stage3[1], golden_stage3[2], golden_stage3[3]
    assert a9["rows"][0][:5] == ["ACGT", 9.0, "1.11E-02", 4, "| s0 s1 s2 s3"]
    assert a10["rows"][0][:5] == ["ACGT", 1.89, "3.89E-01", 4, "| s0 s1 s2 s4"]
    assert a11w["rows"][0][:7] == ["ACGT", 2.41, "6.17E-02", 3.18, 0.63, 4, "| s0 s1 s2 s4"]
    assert a11u["rows"][0][:7] == ["ACGT", 2.63, "4.46E-02", 3.25, 0.6, 4, "| s0 s1 s2 s4"]
    # full-precision values quoted in SURVEY.md A10/A11
    pres, ph, w = stage3_inputs(a10)
    chi2, p, _ = ostats.chi2_row(pres[0], ph, list(w), 2, 8)
    assert chi2 == pytest.approx(1.8858552631578949, rel=1e-14) and p == pytest.approx(0.3894858933804423, rel=1e-14)
    t, p, dof = ostats.ttest_ind_weighted([3.0, 4.0, 5.0, 1.0], [2.0, 0.0, -1.0, 0.5, 1.5],
                                          [0.5, 1.5, 0.7, 1.1], [1.3, 0.9, 1.2, 0.8, 1.0])
    assert t == pytest.approx(2.4146663552427943, rel=1e-13)
    assert p == pytest.approx(0.06171479201819995, rel=1e-12)
    assert dof == pytest.approx(4.883228538221529, rel=1e-13)


def test_ttest_restatement_equals_scipy_for_unit_and_integer_weights():
    from scipy import stats
    rng = np.random.default_rng(3)
    x, y = rng.normal(1, 2, 17), rng.normal(0, 1, 25)
    t, p, _ = ostats.ttest_ind_weighted(x, y, np.ones(17), np.ones(25))
    ref = stats.ttest_ind(x, y, equal_var=False)
    assert t == pytest.approx(ref.statistic, rel=1e-13) and p == pytest.approx(ref.pvalue, rel=1e-12)
    wx, wy = rng.integers(1, 4, 17), rng.integers(1, 4, 25)
    t, p, _ = ostats.ttest_ind_weighted(x, y, wx, wy)
    ref = stats.ttest_ind(np.repeat(x, wx), np.repeat(y, wy), equal_var=False)
    assert t == pytest.approx(ref.statistic, rel=1e-12) and p == pytest.approx(ref.pvalue, rel=1e-11)


def test_vectorised_oracle_equals_literal(golden_stage3):
    for case in golden_stage3:
        pres, ph, w = stage3_inputs(case)
        if case["kind"] == "chi2":
            v = ostats.chi2_rows(pres, ph, w, case["min"], case["max"])
            for i, pv in enumerate(pres):
                r = ostats.chi2_row(pv, ph, list(w), case["min"], case["max"])
                assert (r is not None) == bool(v["tested"][i])
                if r is not None:
                    assert v["n_with"][i] == r[2]
                    if math.isnan(r[0]):
                        assert math.isnan(v["stat"][i])
                    else:
                        assert v["stat"][i] == r[0] and v["p"][i] == r[1]   # bit-identical
        else:
            v = ostats.welch_rows(pres, ph, w, case["min"], case["max"])
            phl = [None if math.isnan(x) else float(x) for x in ph]
            for i, pv in enumerate(pres):
                r = ostats.welch_row(pv, phl, list(w), case["min"], case["max"])
                assert (r is not None) == bool(v["tested"][i])
                if r is not None and not math.isnan(r[1]):
                    assert v["stat"][i] == pytest.approx(r[0], rel=1e-10)
                    assert v["p"][i] == pytest.approx(r[1], rel=1e-9)
                    assert v["mean_x"][i] == pytest.approx(r[2], rel=1e-12)


def test_pack_rows_roundtrip():
    rng = np.random.default_rng(5)
    from phenotypeseeker_b200.pipeline import unpack_rows
    for n in (1, 31, 32, 33, 250, 1000):
        pres = (rng.random((40, n)) < 0.3).astype(np.uint8)
        packed = ok.pack_rows(pres)
        assert packed.shape == (40, (n + 31) // 32)
        assert np.array_equal(unpack_rows(packed, n), pres)


def test_threaded_union_and_rows_equal_the_plain_helpers():
    """oracle.kmers.union_and_rows / count_many (what the full-size GPU tests use) against union() +
    presence_matrix() + pack_rows() on a set small enough for both, cut into many k-mer ranges."""
    from phenotypeseeker_b200 import synth
    ds = synth.make_dataset(37, genome_len=30_000, seed=11)
    lists = ok.count_many(ds.files, 13, threads=4)
    assert all(np.array_equal(a[0], ok.count_kmers(f, 13)[0]) for a, f in zip(lists[:3], ds.files[:3]))
    u, rows = ok.union_and_rows([l[0] for l in lists], threads=4, chunk=50_000)
    u0 = ok.union([l[0] for l in lists])
    assert np.array_equal(u, u0)
    packed = ok.pack_rows(ok.presence_matrix(u0, lists))
    assert rows.shape[1] == 4 and np.array_equal(rows[:, :packed.shape[1]], packed) and not rows[:, packed.shape[1]:].any()
