"""oracle/stats.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the per-k-mer association tests of PhenotypeSeeker
(modeling.py:716-858). Two forms:

  * `chi2_row` / `welch_row`: literal per-k-mer restatement (pure-Python loops
    in sample order, like the reference) — small cases and golden pinning;
  * `chi2_rows` / `welch_rows`: the same arithmetic vectorised over k-mers with
    numpy (still sequential over samples, so weighted sums round exactly like
    the reference's `+=` loop, modeling.py:809-823).

Pinned against the real reference methods by oracle/make_golden.py
(tests/golden/stage3.json). `ttest_ind` of statsmodels (modeling.py:44,734) is
not installed and its version is pinned nowhere in the reference: the
restatement below follows statsmodels' DescrStatsW/CompareMeans formulas
 and is "restated", checked against scipy.stats.ttest_ind for
unit / integer weights.

Phenotype coding used throughout: binary 1 / 0 / -1 (= "NA"); continuous
float with NaN = "NA".
"""
import math

import numpy as np
from scipy import stats as _st


# ---------------------------------------------------------------------------
# literal restatements (one k-mer)

def chi2_row(presence, pheno, weights, min_samples, max_samples):
    """modeling.py:759-858 for one k-mer -> (chi2, p, n_with) or None.

    presence: N ints (count or 0/1); pheno: N of 1/0/-1(NA); weights: N.
    Returns None when the min/max filter rejects (:770-772). chi2/p may be NaN.
    """
    a = b = c = d = 0  # int 0 like the reference; becomes float with float weights
    n_with = n_without = 0
    for i in range(len(presence)):
        if pheno[i] == 1:
            if presence[i] != 0:
                a += weights[i]; n_with += 1
            else:
                b += weights[i]; n_without += 1
        elif pheno[i] == 0:
            if presence[i] != 0:
                c += weights[i]; n_with += 1
            else:
                d += weights[i]; n_without += 1
    if n_with < min_samples or n_without < 2 or n_with > max_samples:
        return None
    w_pheno = a + b
    wo_pheno = c + d
    w_kmer = a + c
    wo_kmer = b + d
    total = w_pheno + wo_pheno
    with np.errstate(all="ignore"):
        tot = np.float64(total)
        exp = [np.float64(w_pheno * w_kmer) / tot, np.float64(w_pheno * wo_kmer) / tot,
               np.float64(wo_pheno * w_kmer) / tot, np.float64(wo_pheno * wo_kmer) / tot]
        obs = [np.float64(a), np.float64(b), np.float64(c), np.float64(d)]
        # scipy.stats.chisquare(obs, exp, ddof=1): terms summed in cell order,
        # p = chi2.sf(stat, 4 - 1 - 1 = 2) = exp(-stat/2)   (modeling.py:782-792)
        terms = [(o - e) ** 2 / e for o, e in zip(obs, exp)]
        chi2 = ((terms[0] + terms[1]) + terms[2]) + terms[3]
        p = float(_st.chi2.sf(chi2, 2))
    return float(chi2), p, n_with


def ttest_ind_weighted(x, y, wx, wy):
    """statsmodels.stats.weightstats.ttest_ind(x, y, usevar='unequal',
    weights=(wx, wy)) restated -> (t, p, dof)."""
    x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    wx = np.asarray(wx, dtype=np.float64); wy = np.asarray(wy, dtype=np.float64)
    with np.errstate(all="ignore"):
        n1 = wx.sum(); n2 = wy.sum()
        m1 = np.dot(x, wx) / n1; m2 = np.dot(y, wy) / n2
        v1 = np.dot((x - m1) ** 2, wx) / n1
        v2 = np.dot((y - m2) ** 2, wy) / n2
        s1 = v1 / (n1 - 1); s2 = v2 / (n2 - 1)
        t = (m1 - m2) / np.sqrt(s1 + s2)
        r1 = s1 / (s1 + s2); r2 = s2 / (s1 + s2)
        dof = 1.0 / (r1 ** 2 / (n1 - 1) + r2 ** 2 / (n2 - 1))
        p = _st.t.sf(np.abs(t), dof) * 2
    return float(t), float(p), float(dof)


def welch_row(presence, pheno, weights, min_samples, max_samples):
    """modeling.py:716-757 for one k-mer -> (t, p, mean_x, mean_y, n_with) or None."""
    x, y, wx, wy = [], [], [], []
    for i in range(len(presence)):
        ph = pheno[i]
        if ph is None or (isinstance(ph, float) and math.isnan(ph)):
            continue
        if presence[i] == 0:
            y.append(ph); wy.append(weights[i])
        else:
            x.append(ph); wx.append(weights[i])
    if len(x) < min_samples or len(y) < 2 or len(x) > max_samples:
        return None
    t, p, _ = ttest_ind_weighted(x, y, wx, wy)
    mean_x = float(np.average(x, weights=wx))
    mean_y = float(np.average(y, weights=wy))
    return t, p, mean_x, mean_y, len(x)


def passes(p, pvalue_cutoff, n_kmers, omit_b, binary):
    """The keep rule: chi2 honours --omit_B_correction (:795), the t-test never (:738)."""
    if binary and omit_b and p < pvalue_cutoff:
        return True
    return p < (pvalue_cutoff / n_kmers)


# ---------------------------------------------------------------------------
# vectorised over k-mers

def chi2_rows(presence, pheno, weights, min_samples, max_samples):
    """presence: U x N (0/!0); -> dict of arrays (chi2, p, n_with, tested mask)."""
    presence = np.asarray(presence) != 0
    pheno = np.asarray(pheno)
    U, N = presence.shape
    int_w = all(float(w) == int(w) for w in weights)
    acc_t = np.float64
    a = np.zeros(U, acc_t); b = np.zeros(U, acc_t); c = np.zeros(U, acc_t); d = np.zeros(U, acc_t)
    n_with = np.zeros(U, np.int64); n_without = np.zeros(U, np.int64)
    for s in range(N):  # sequential in sample order, like modeling.py:809-823
        w = np.float64(weights[s])
        col = presence[:, s]
        if pheno[s] == 1:
            a += np.where(col, w, 0.0); b += np.where(col, 0.0, w)
        elif pheno[s] == 0:
            c += np.where(col, w, 0.0); d += np.where(col, 0.0, w)
        else:
            continue
        n_with += col; n_without += ~col
    tested = ~((n_with < min_samples) | (n_without < 2) | (n_with > max_samples))
    with np.errstate(all="ignore"):
        w_pheno = a + b; wo_pheno = c + d; w_kmer = a + c; wo_kmer = b + d
        total = w_pheno + wo_pheno
        e = [w_pheno * w_kmer / total, w_pheno * wo_kmer / total,
             wo_pheno * w_kmer / total, wo_pheno * wo_kmer / total]
        o = [a, b, c, d]
        t = [(oo - ee) ** 2 / ee for oo, ee in zip(o, e)]
        chi2 = ((t[0] + t[1]) + t[2]) + t[3]
        p = _st.chi2.sf(chi2, 2)
    del int_w
    return {"stat": chi2, "p": p, "n_with": n_with, "tested": tested}


def welch_rows(presence, pheno, weights, min_samples, max_samples):
    """presence U x N; pheno float with NaN=NA -> dict(stat,p,mean_x,mean_y,n_with,tested)."""
    presence = np.asarray(presence) != 0
    ph = np.asarray(pheno, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float64)
    ok = ~np.isnan(ph)
    X = presence[:, ok]; phv = ph[ok]; wv = w[ok]
    Y = ~X
    with np.errstate(all="ignore"):
        n_with = X.sum(1); n_without = Y.sum(1)
        n1 = X @ wv; n2 = Y @ wv
        m1 = (X @ (wv * phv)) / n1; m2 = (Y @ (wv * phv)) / n2
        dx = (phv[None, :] - m1[:, None]) ** 2
        dy = (phv[None, :] - m2[:, None]) ** 2
        v1 = (np.where(X, dx, 0.0) @ wv) / n1
        v2 = (np.where(Y, dy, 0.0) @ wv) / n2
        s1 = v1 / (n1 - 1); s2 = v2 / (n2 - 1)
        t = (m1 - m2) / np.sqrt(s1 + s2)
        r1 = s1 / (s1 + s2); r2 = s2 / (s1 + s2)
        dof = 1.0 / (r1 ** 2 / (n1 - 1) + r2 ** 2 / (n2 - 1))
        p = _st.t.sf(np.abs(t), dof) * 2
    tested = ~((n_with < min_samples) | (n_without < 2) | (n_with > max_samples))
    return {"stat": t, "p": p, "mean_x": m1, "mean_y": m2, "n_with": n_with,
            "dof": dof, "tested": tested}
