"""Scale probes for the configs bench.py does not time (BASELINE.json configs 2-4):
  welch : N = 1000 continuous phenotype (2 % NA), weighted, on a loaded random matrix
  wide  : N = 5000, P = 10 binary phenotypes on a loaded random matrix (row = 640 B)
  reads : raw-read FASTQ samples at full depth (30x of 4.3 Mbp), per-sample counting + cutoff
Usage: python profiles/scale_probe.py [welch] [wide] [reads]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from phenotypeseeker_b200 import synth
from phenotypeseeker_b200._native import Context


def rand_rows(rng, U, N, W):
    """Random bit matrix with a realistic mix: 70 % rare rows, 25 % near-core rows, 5 % mid-frequency."""
    rows = np.zeros((U, W), dtype=np.uint32)
    kind = rng.random(U)
    full_words = N // 32
    dens = np.where(kind < 0.70, 2.0 / N, np.where(kind < 0.95, 0.98, 0.3))
    for w in range((N + 31) // 32):
        nb = 32 if w < full_words else N - 32 * full_words
        bits = rng.random((U, nb)) < dens[:, None]
        rows[:, w] = (bits.astype(np.uint64) << np.arange(nb, dtype=np.uint64)).sum(axis=1).astype(np.uint32)
    return rows


def timed(ctx, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.time()
        out = fn()
        best = min(best, time.time() - t0)
    return best, out


def welch(ctx):
    rng = np.random.default_rng(1)
    N, U = 1000, 2_000_000
    ctx.begin(16, N)
    W = ctx.row_words()
    rows = rand_rows(rng, U, N, W)
    ctx.load_matrix(rows)
    ph = np.round(rng.normal(0, 2, N), 3)
    ph[rng.random(N) < 0.02] = np.nan
    w = rng.gamma(2.0, 0.5, N) + 0.05
    ctx.profile(True)
    t, ns = timed(ctx, lambda: ctx.test_welch(ph, w, 2, N - 2, 0.05 / U))
    ctx.profile(False)
    tab = ctx.profile_table()
    k = tab["test_welch"]
    print(f"welch  N={N} U={U}: {1e3 * t:.2f} ms wall, kernel {k['ms'] / k['launches']:.3f} ms "
          f"({U / (k['ms'] / k['launches'] * 1e-3):.3e} rows/s, matrix {U * W * 4 / 1e6:.0f} MB -> "
          f"{U * W * 4 / (k['ms'] / k['launches'] * 1e-3) / 1e9:.0f} GB/s), survivors {ns}")


def wide(ctx):
    rng = np.random.default_rng(2)
    N, U, P = 5000, 400_000, 10
    ctx.begin(16, N)
    W = ctx.row_words()
    rows = rand_rows(rng, U, N, W)
    ctx.load_matrix(rows)
    ph = (rng.random((P, N)) < 0.4).astype(np.int8)
    ph[rng.random((P, N)) < 0.01] = -1
    ctx.profile_reset(); ctx.profile(True)
    t, ns = timed(ctx, lambda: ctx.test_chi2(ph, None, 2, N - 2, 0.05 / U))
    t2, ns2 = timed(ctx, lambda: ctx.test_chi2(ph, rng.gamma(2.0, 0.5, N), 2, N - 2, 0.05 / U))
    ctx.profile(False)
    tab = ctx.profile_table()
    for name in ("test_chi2", "test_chi2_w"):
        k = tab[name]
        ms = k["ms"] / k["launches"]
        print(f"wide   N={N} P={P} U={U} {name}: kernel {ms:.3f} ms ({U * P / (ms * 1e-3):.3e} tests/s, "
              f"matrix {U * W * 4 / 1e6:.0f} MB -> {U * W * 4 / (ms * 1e-3) / 1e9:.0f} GB/s)")
    print(f"       survivors {ns} / {ns2}")


def reads(ctx):
    from oracle import kmers as ok
    t0 = time.time()
    ds = synth.make_dataset(3, genome_len=4_300_000, seed=20260104, reads=True, coverage=30.0)
    print(f"reads  generated 3 FASTQ samples, {ds.total_bytes() / 1e6:.0f} MB in {time.time() - t0:.0f} s")
    for cutoff in (1, 3):
        ctx.begin(16, 3, cutoff)
        t0 = time.time()
        ctx.add_samples(0, ds.files)
        t1 = time.time()
        U = ctx.build_union()
        t2 = time.time()
        km, ct = ctx.sample_kmers(0, cutoff)
        print(f"reads  cutoff={cutoff}: ingest+count {1e3 * (t1 - t0):.0f} ms for 3 samples "
              f"({ds.total_bytes() / (t1 - t0) / 1e9:.1f} GB/s of FASTQ), union {1e3 * (t2 - t1):.1f} ms, "
              f"U={U}, distinct(sample0)={len(km)}")
        if cutoff == 3:
            okm, oct_ = ok.count_kmers(ds.files[0], 16, cutoff)
            print("       sample 0 vs oracle:", bool(np.array_equal(km, okm) and np.array_equal(ct, oct_)))


if __name__ == "__main__":
    which = sys.argv[1:] or ["welch", "wide", "reads"]
    ctx = Context(0)
    for w in which:
        {"welch": welch, "wide": wide, "reads": reads}[w](ctx)
