"""oracle/ref_pipeline.py — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's hot path exactly as it runs today
(modeling.py:1644-1683): per-sample `glistmaker`, a tree of `glistcompare -u`
unions, `glistquery -l` text mapping + `split -n r/T`, then T worker processes
that zip the N stripe files line by line and run one scipy chi-square (or the
restated Welch test) per k-mer in pure Python.

The GenomeTester4 binaries are the reference's own (installed unmodified into
oracle/_ref/bin by oracle/build.py; /root/reference/bin in the build container).
modeling.py itself cannot travel to the GPU box, so stage 3 is restated here
(oracle/stats.py holds the arithmetic, pinned to the real methods by
tests/golden/stage3.json). Used by bench.py for `cpu_baseline` and
`--impl reference`, and by tests to cross-check whole-path results.
"""
import math
import multiprocessing as mp
import os
import shutil
import subprocess
import tempfile
import time

import numpy as np
from scipy import stats as _st

from . import build as _build


def _run(cmd, **kw):
    return subprocess.run(cmd, shell=True, **kw)


def _glistmaker(args):
    bindir, path, out_prefix, k, cutoff = args
    # modeling.py:309-310 passes `-c <cutoff>`; the shipped 4.2.3 binary accepts and ignores it
    _run(f"{bindir}/glistmaker {path} -o {out_prefix} -w {k} -c {cutoff}", capture_output=True)


def _map_sample(args):
    bindir, workdir, name, k, T = args
    mapped = f"{workdir}/{name}_mapped.txt"
    with open(mapped, "w") as f:
        _run(f"{bindir}/glistquery {workdir}/{name}_0_{k}.list -l {workdir}/feature_vector.list", stdout=f)
    _run(f"split -a 5 -d -n r/{T} {mapped} {workdir}/{name}_mapped_")
    os.remove(mapped)


def _union(args):
    bindir, workdir, names, rnd, k = args
    ins = " ".join(f"{workdir}/{n}_{rnd}_{k}" + ("_union" if rnd > 0 else "") + ".list" for n in names)
    _run(f"{bindir}/glistcompare -u -o {workdir}/{names[0]}_{rnd + 1} {ins}", capture_output=True)
    return f"{workdir}/{names[0]}_{rnd + 1}_{k}_union.list"


def _chi2_kmer(kmer, vec, pheno, weights, names, mn, mx, cutoff, omit_b, U):
    """conduct_chi_squared_test (modeling.py:759-858) for one k-mer."""
    a = b = c = d = 0
    n_wo = 0
    w_names = []
    for i, ph in enumerate(pheno):
        if ph == 1:
            if vec[i] != 0:
                a += weights[i]; w_names.append(names[i])
            else:
                b += weights[i]; n_wo += 1
        elif ph == 0:
            if vec[i] != 0:
                c += weights[i]; w_names.append(names[i])
            else:
                d += weights[i]; n_wo += 1
    n_w = len(w_names)
    if n_w < mn or n_wo < 2 or n_w > mx:
        return None
    w_ph, wo_ph, w_k, wo_k = a + b, c + d, a + c, b + d
    tot = float(w_ph + wo_ph)
    chi2, p = _st.chisquare([a, b, c, d],
                            [w_ph * w_k / tot, w_ph * wo_k / tot, wo_ph * w_k / tot, wo_ph * wo_k / tot], 1)
    if (omit_b and p < cutoff) or p < cutoff / U:
        return [kmer, round(chi2, 2), "%.2E" % p, n_w, " ".join(["|"] + w_names)] + vec
    return None


def _welch_kmer(kmer, vec, pheno, weights, names, mn, mx, cutoff, U):
    """conduct_t_test (modeling.py:716-757) for one k-mer."""
    from .stats import ttest_ind_weighted
    x, y, wx, wy, w_names = [], [], [], [], []
    for i, ph in enumerate(pheno):
        if ph is None:
            continue
        if vec[i] == 0:
            y.append(ph); wy.append(weights[i])
        else:
            x.append(ph); wx.append(weights[i]); w_names.append(names[i])
    if len(x) < mn or len(y) < 2 or len(x) > mx:
        return None
    t, p, _ = ttest_ind_weighted(x, y, wx, wy)
    if p < cutoff / U:
        return [kmer, round(t, 2), "%.2E" % p, round(float(np.average(x, weights=wx)), 2),
                round(float(np.average(y, weights=wy)), 2), len(w_names), " ".join(["|"] + w_names)] + vec
    return None


def _test_stripe(args):
    """get_kmers_tested (modeling.py:677-714): zip the N stripe files, one test per line."""
    (files, pheno, weights, names, binary, mn, mx, cutoff, omit_b, U) = args
    out = {}
    handles = [open(f) for f in files]
    with np.errstate(all="ignore"):
        for lines in zip(*handles):
            kmer = lines[0].split()[0]
            vec = [1 if int(j.split()[1]) > 0 else 0 for j in lines]
            if binary:
                r = _chi2_kmer(kmer, vec, pheno, weights, names, mn, mx, cutoff, omit_b, U)
            else:
                r = _welch_kmer(kmer, vec, pheno, weights, names, mn, mx, cutoff, U)
            if r:
                out[r[0]] = r[1:]
    for h in handles:
        h.close()
    return out


def run(paths, names, k, pheno_cols, binary, weights=None, min_samples=2, max_samples=None,
        pvalue_cutoff=0.05, omit_b=False, threads=None, workdir=None, keep=False, cutoff=1):
    """The reference hot path on CPU. pheno_cols: list of per-sample lists (1/0/None or float/None).

    Returns dict(U, results=[{kmer: row}], t_stage12, t_stage3, threads).
    """
    bindir = _build.ref_bin_dir()
    if bindir is None:
        raise RuntimeError("GenomeTester4 binaries not found (oracle/_ref/bin); run oracle/build.py "
                           "in the build container")
    N = len(paths)
    T = threads or os.cpu_count() or 1
    if max_samples is None:
        max_samples = N - 2
    weights = list(weights) if weights is not None else [1] * N
    own = workdir is None
    workdir = workdir or tempfile.mkdtemp(prefix="psref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    os.makedirs(workdir, exist_ok=True)
    t0 = time.time()
    with mp.Pool(T) as pool:
        pool.map(_glistmaker, [(bindir, p, f"{workdir}/{n}_0", k, cutoff) for p, n in zip(paths, names)])
        # ⌊log2 N⌋ rounds of pairwise (last group: up to 3) unions, modeling.py:351-365
        groups = [[n] for n in names]
        last = None
        for rnd in range(int(math.log(N, 2)) if N > 1 else 0):
            heads = [g[0] for g in groups]
            groups = []
            j = 0
            while j + 2 <= len(heads):
                take = 3 if len(heads) - j == 3 else 2   # an odd leftover joins the last pair
                groups.append(heads[j:j + take])
                j += take
            outs = pool.map(_union, [(bindir, workdir, g, rnd, k) for g in groups])
            last = outs[-1]
        if last is None:
            last = f"{workdir}/{names[0]}_0_{k}.list"
            shutil.copy(last, f"{workdir}/feature_vector.list")
        else:
            shutil.move(last, f"{workdir}/feature_vector.list")
        pool.map(_map_sample, [(bindir, workdir, n, k, T) for n in names])
    U = int(subprocess.run(f"{bindir}/glistquery {workdir}/feature_vector.list | wc -l", shell=True,
                           capture_output=True, text=True).stdout)
    t1 = time.time()
    results = []
    with mp.Pool(T) as pool:
        for col in pheno_cols:
            stripes = [[f"{workdir}/{n}_mapped_{t:05d}" for n in names] for t in range(T)]
            parts = pool.map(_test_stripe, [(s, col, weights, names, binary, min_samples, max_samples,
                                             pvalue_cutoff, omit_b, U) for s in stripes])
            merged = {}
            for d in parts:   # stripe-major order, like pd.concat(axis=1) at modeling.py:670-672
                merged.update(d)
            results.append(merged)
    t2 = time.time()
    if own and not keep:
        shutil.rmtree(workdir, ignore_errors=True)
    return {"U": U, "results": results, "t_stage12": t1 - t0, "t_stage3": t2 - t1, "threads": T}
