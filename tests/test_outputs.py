"""Whole-path output parity: the files `phenotypeseeker modeling` writes (chi2_results_*.tsv,
t-test_results_*.tsv, *_top<N>.tsv, *_MLdf.csv) reproduced byte-for-byte against golden files
written by the UNMODIFIED reference CLI (tests/golden/cli_*). The CPU variant feeds the host-side
writer from the oracle; the GPU variant goes through the CUDA path."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD
from oracle import kmers as ok
from oracle import stats as ostats
from phenotypeseeker_b200 import synth
from phenotypeseeker_b200 import modeling_gpu as mg
from phenotypeseeker_b200.pipeline import PhenoResult


def _case(tag):
    d = os.path.join(GOLD, f"cli_{tag}")
    with open(os.path.join(d, "case.json")) as f:
        c = json.load(f)
    ds = synth.config(c["config"], tiny=True, n_samples=c["n_samples"], genome_len=c["genome_len"])
    a = c["args"]
    opt = lambda flag, default: a[a.index(flag) + 1] if flag in a else default
    return d, c, ds, dict(k=int(opt("-l", 13)), T=int(opt("-nt", 8)), limit=int(opt("--n_kmers", 1000)),
                          pvalue=float(opt("--pvalue", 0.05)), omit_b="--omit_B_correction" in a)


def _pheno_values(ds, j=0):
    out = []
    for v in ds.pheno[:, j]:
        out.append("NA" if np.isnan(v) else (int(v) if ds.binary else float(v)))
    return out


def _compare_files(gold_dir, files, outdir):
    for fn in files:
        with open(os.path.join(gold_dir, fn), "rb") as f:
            exp = f.read()
        with open(os.path.join(outdir, fn), "rb") as f:
            got = f.read()
        assert got == exp, fn


def _oracle_result(ds, o):
    lists = [ok.count_kmers(f, o["k"]) for f in ds.files]
    u = ok.union([l[0] for l in lists])
    pres = ok.presence_matrix(u, lists)
    N = ds.n_samples
    if ds.binary:
        code = np.where(np.isnan(ds.pheno[:, 0]), -1, ds.pheno[:, 0]).astype(np.int8)
        r = ostats.chi2_rows(pres, code, np.ones(N), 2, N - 2)
        thr = o["pvalue"] if o["omit_b"] else o["pvalue"] / len(u)
    else:
        r = ostats.welch_rows(pres, ds.pheno[:, 0], np.ones(N), 2, N - 2)
        thr = o["pvalue"] / len(u)
    keep = r["tested"] & (r["p"] < thr)
    rows = np.nonzero(keep)[0]
    z = np.zeros(len(rows))
    res = PhenoResult(name="pheno1", kmer=u[keep], row=rows.astype(np.uint64), stat=r["stat"][keep], p=r["p"][keep],
                      mean_x=r.get("mean_x", z)[keep] if "mean_x" in r else z,
                      mean_y=r.get("mean_y", z)[keep] if "mean_y" in r else z,
                      n_with=r["n_with"][keep], presence=pres[keep], na_mask=np.isnan(ds.pheno[:, 0]))
    return res


@pytest.mark.parametrize("tag", ["chi2", "ttest"])
def test_writer_reproduces_reference_files_from_oracle(tag, tmp_path):
    gold_dir, c, ds, o = _case(tag)
    res = _oracle_result(ds, o)
    df = mg.build_ml_df(res, o["k"], ds.names, o["T"], ds.binary)
    mg.write_outputs(df, "pheno1", ds.names, [1] * ds.n_samples, _pheno_values(ds), ds.binary, o["limit"], str(tmp_path))
    _compare_files(gold_dir, c["files"], str(tmp_path))


def test_stripe_major_order():
    rows = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
    assert list(rows[mg.stripe_major_order(rows, 4)]) == [0, 4, 8, 1, 5, 9, 2, 6, 10, 3, 7]
    assert list(rows[mg.stripe_major_order(rows, 1)]) == list(rows)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["chi2", "ttest"])
def test_gpu_path_reproduces_reference_files(tag, tmp_path):
    gold_dir, c, ds, o = _case(tag)
    N = ds.n_samples
    U, dfs = mg.run_hot_path(ds.files, ds.names, o["k"], 1, ds.pheno[:, :1], ["pheno1"], ds.binary, None, 2, N - 2,
                             o["pvalue"], o["omit_b"], o["T"])
    mg.write_outputs(dfs["pheno1"], "pheno1", ds.names, [1] * N, _pheno_values(ds), ds.binary, o["limit"], str(tmp_path))
    _compare_files(gold_dir, c["files"], str(tmp_path))


@pytest.mark.parametrize("tag", ["chi2", "ttest"])
def test_legacy_named_files_hold_the_same_rows(tag, tmp_path):
    """README.md:96-105 layout (no header, statistic / "%.2E" p / [means] / n / "| names"): the rows of the
    golden chi2_results / t-test_results TSV written by the real CLI, in test (stripe-major) order."""
    gold_dir, c, ds, o = _case(tag)
    res = _oracle_result(ds, o)
    df = mg.build_ml_df(res, o["k"], ds.names, o["T"], ds.binary)
    paths = mg.write_legacy_outputs(df, "pheno1", ds.binary, str(tmp_path))
    assert [os.path.basename(p) for p in paths] == [
        ("chi-squared_test_results_pheno1.txt" if ds.binary else "t-test_results_pheno1.txt"),
        "k-mers_filtered_by_pvalue_pheno1.txt"]
    got = open(paths[0]).read().splitlines()
    assert got == open(paths[1]).read().splitlines()
    stat_file = [f for f in c["files"] if f.endswith(".tsv") and "_top" not in f][0]
    gold = open(os.path.join(gold_dir, stat_file)).read().splitlines()[1:]       # drop the header
    assert sorted(got) == sorted(gold)                                            # same rows, other order
    assert list(df.columns) == [l.split("\t")[0] for l in got]                    # test order
    ncol = 5 if ds.binary else 7
    assert all(len(l.split("\t")) == ncol and l.split("\t")[-1].startswith("|") for l in got)
