"""oracle/make_golden.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Generates tests/golden/* from the REAL reference in this container:

  kmer_lists.json  small FASTA/FASTQ inputs -> (k-mer, count) lists printed by the
                   shipped `glistmaker | glistquery` (GenomeTester4 4.2.3)
  union_map.json   `glistcompare -u` + `glistquery -l` outputs on two tiny lists
  stage3.json      rows returned by the real `phenotypes.conduct_chi_squared_test`
                   / `conduct_t_test` (modeling.py:716-858) on random presence
                   vectors, phenotypes (with NA) and weights
  cli_*/           whole-CLI runs of the unmodified reference on tiny synthetic
                   data sets (inputs regenerated from synth with the recorded seed)

Run:  python -m oracle.make_golden      (needs /root/reference; not needed at test time)
"""
import base64
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

from . import build, kmers, ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

A_FA = b">c1 test\nACGTTGCAAGGCTTAACCGGTTNACGTACGTAGGCTAGCTAGGATCC\nacgttgcaaggcttaa\n>c2\nTTTTTTTTTTTTTTTTTTTT\n"
B_FA = b">x\nACGTACGTRACGTTTGGCCAA-ACGT*ACGTAC\n"
R_FQ = (b"@r1\nACGTTGCAAGGC\n+\nIIIIIIIIIIII\n@r2\nACGTTGCAAGGN\n+r2\nIIIIIIIIIIII\n"
        b"@r3\nGCCTTGCAACGT\n+\nIIIIIIIIIIII\n")


def _rand_fasta(rng, n_rec, max_len, lower=0.1, bad=0.01, crlf=False, width=60):
    out = bytearray()
    for r in range(n_rec):
        out += b">rec%d some text > more\n" % r
        n = int(rng.integers(0, max_len))
        seq = rng.choice(list(b"ACGT"), size=n, p=[0.355, 0.145, 0.145, 0.355]).astype(np.uint8)
        lo = rng.random(n) < lower
        seq[lo] |= 0x20
        bd = rng.random(n) < bad
        seq[bd] = rng.choice(list(b"NRYKMnx-*. U u"), size=int(bd.sum()))
        for i in range(0, n, width):
            out += seq[i:i + width].tobytes() + (b"\r\n" if crlf else b"\n")
        if rng.random() < 0.3:
            out += b"\n"
    return bytes(out)


def _rand_fastq(rng, n_reads, read_len, crlf=False):
    out = bytearray()
    nl = b"\r\n" if crlf else b"\n"
    for r in range(n_reads):
        seq = rng.choice(list(b"ACGT"), size=read_len).astype(np.uint8)
        seq[rng.random(read_len) < 0.02] = ord("N")
        q = rng.choice(list(b"@>I#5+"), size=read_len).astype(np.uint8)
        out += b"@read%d" % r + nl + seq.tobytes() + nl + (b"+read%d" % r if r % 3 == 0 else b"+") + nl
        out += q.tobytes() + nl
    return bytes(out if rng.random() < 0.5 else out[:-len(nl)])


def gen_kmer_lists():
    rng = np.random.default_rng(7)
    cases = [("appendixA_a_fa", A_FA, ".fa", [5, 4, 13]), ("appendixA_b_fa", B_FA, ".fa", [4]),
             ("appendixA_r_fq", R_FQ, ".fq", [5])]
    edge = {
        "gt_midline": b">h\nACGTA>CGTAC\nGGGGG\n",
        "space_breaks_tab_skips": b">h\nACGT ACGT\nACGT\tACGT\n",
        "crlf": b">h\r\nACGT\r\nACGT\r\n",
        "blank_lines_no_trailing_nl": b">h\nACGT\n\n\nACGT",
        "leading_garbage": b"ACGTACGT\n;x\n\n >h\nGGGGG\n",
        "empty_record": b">h1\n>h2\nACGTAC\n",
        "bases_in_header": b">h ACGTACGTACGT\nTTTTT\n",
        "uracil": b">h\nACGU\nACGuAC\n",
        "cr_in_header": b">h\rACGTAC\nGGGGG\n",
        "only_header": b">h\n",
        "empty": b"",
        "no_records": b"ACGTACGT\n",
        "short_seq": b">h\nACG\n>g\nAC\nG\nT\n",
        "fastq_quality_at_gt": b"@r1\nACGTAC\n+\n@IIIII\n@r2\nGGGGG\n+r2\n>@>II\n@r3\nTTTTT\n+\nIIIII",
        "fastq_crlf": b"@r1\r\nACGTAC\r\n+\r\nIIIIII\r\n@r2\r\nGGGGG\r\n+\r\nIIIII\r\n",
        "fastq_leading_blank": b"\n\n@r1\nACGTAC\n+\nIIIIII\n",
        "fastq_tab_space": b"@r1\nACGT\tACGT ACGT\n+\nIIIIIIIIIIIIII\n",
    }
    for name, data in edge.items():
        sfx = ".fq" if name.startswith("fastq") else ".fa"
        cases.append((name, data, sfx, [4]))
    cases.append(("rand_fa_1", _rand_fasta(rng, 5, 700), ".fa", [13, 16]))
    cases.append(("rand_fa_crlf", _rand_fasta(rng, 4, 500, crlf=True), ".fa", [16, 21]))
    cases.append(("rand_fa_k32", _rand_fasta(rng, 3, 900, bad=0.002), ".fa", [32, 31, 17]))
    cases.append(("rand_fq_1", _rand_fastq(rng, 40, 50), ".fq", [13, 16]))
    cases.append(("rand_fq_crlf", _rand_fastq(rng, 25, 75, crlf=True), ".fq", [16, 24]))
    out = []
    for name, data, sfx, ks in cases:
        for k in ks:
            km, ct = kmers.run_glistmaker(data, k, sfx)
            okm, oct_ = kmers.count_kmers(data, k)
            agree = bool(np.array_equal(km, okm) and np.array_equal(ct, oct_))
            print(f"  {name:32s} k={k:2d} n={len(km):5d} oracle_agrees={agree}")
            out.append({"name": name, "k": k, "suffix": sfx,
                        "data_b64": base64.b64encode(data).decode(),
                        "kmers": [int(x) for x in km], "counts": [int(x) for x in ct]})
    with open(os.path.join(GOLD, "kmer_lists.json"), "w") as f:
        json.dump(out, f)


def gen_union_map():
    bindir = build.ref_bin_dir()
    res = {}
    with tempfile.TemporaryDirectory() as td:
        for nm, data in (("a.fa", A_FA), ("r.fq", R_FQ)):
            with open(os.path.join(td, nm), "wb") as f:
                f.write(data)
        run = lambda *a: subprocess.run([os.path.join(bindir, a[0])] + list(a[1:]), cwd=td,
                                        capture_output=True, text=True)
        run("glistmaker", "a.fa", "-o", "a", "-w", "5")
        run("glistmaker", "r.fq", "-o", "r", "-w", "5")
        run("glistcompare", "-u", "-o", "u", "a_5.list", "r_5.list")
        res["union"] = run("glistquery", "u_5_union.list").stdout
        res["map_r"] = run("glistquery", "r_5.list", "-l", "u_5_union.list").stdout
        res["map_a"] = run("glistquery", "a_5.list", "-l", "u_5_union.list").stdout
        with open(os.path.join(td, "a_5.list"), "rb") as f:
            res["a_5_list_b64"] = base64.b64encode(f.read()).decode()
    res["a_fa_b64"] = base64.b64encode(A_FA).decode()
    res["r_fq_b64"] = base64.b64encode(R_FQ).decode()
    with open(os.path.join(GOLD, "union_map.json"), "w") as f:
        json.dump(res, f)


def gen_gmer_counter():
    """prediction.py:72-100: `gmer_counter -db db.txt sample` on tiny samples (k = 13, 16)."""
    bindir = build.ref_bin_dir()
    rng = np.random.default_rng(23)
    out = []
    for k in (13, 16):
        files = [_rand_fasta(rng, 3, 900, bad=0.003), _rand_fastq(rng, 30, 60), _rand_fasta(rng, 2, 700, crlf=True)]
        pool = np.concatenate([kmers.count_kmers(f, k)[0] for f in files])
        q = [int(x) for x in rng.choice(pool, 40)]
        qs = [kmers.kmer_to_str(x, k) for x in q]
        # write half of them reverse-complemented: single-k-mer nodes are strand-agnostic (Appendix A7)
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        qs = [s if i % 2 else "".join(comp[c] for c in reversed(s)) for i, s in enumerate(qs)]
        qs += ["A" * k, "ACGT" * 4 if k == 16 else "ACGTACGTACGTA"]
        qs = list(dict.fromkeys(qs))
        rows = []
        with tempfile.TemporaryDirectory() as td:
            with open(os.path.join(td, "db.txt"), "w") as f:
                for s in qs:
                    f.write(f"{s}\t1\t{s}\n")
            for i, data in enumerate(files):
                sfx = ".fq" if data[:1] == b"@" else ".fa"
                with open(os.path.join(td, f"s{i}{sfx}"), "wb") as f:
                    f.write(data)
                txt = subprocess.run([os.path.join(bindir, "gmer_counter"), "-db", os.path.join(td, "db.txt"),
                                      os.path.join(td, f"s{i}{sfx}")], capture_output=True, text=True).stdout
                cnt = [int(l.split()[2]) for l in txt.splitlines() if not l.startswith("#")]
                rows.append(cnt)
        out.append({"k": k, "kmers": qs, "files_b64": [base64.b64encode(f).decode() for f in files], "counts": rows})
        print(f"  gmer_counter k={k}: {len(qs)} k-mers x {len(files)} samples, max count {max(map(max, rows))}")
    with open(os.path.join(GOLD, "gmer_counter.json"), "w") as f:
        json.dump(out, f)


def gen_stage3():
    m = ref_shim.load_modeling()
    rng = np.random.default_rng(11)
    out = []

    def run_case(kind, N, weights, pheno, presences, mn, mx, cutoff, omit_b, U):
        names = [f"s{i}" for i in range(N)]
        samples = [ref_shim.FakeSample(names[i], {"ph": pheno[i]}, weights[i]) for i in range(N)]
        m.Samples.min_samples, m.Samples.max_samples = mn, mx
        m.phenotypes.pvalue_cutoff = cutoff
        m.phenotypes.omit_B = omit_b
        obj = m.phenotypes("ph")
        obj.no_kmers_to_analyse = U
        rows = []
        for pv in presences:
            fn = obj.conduct_chi_squared_test if kind == "chi2" else obj.conduct_t_test
            r = fn("ACGT", list(pv), samples)
            if r is not None:
                r = [x.item() if hasattr(x, "item") else x for x in r]
            rows.append(r)
        out.append({"kind": kind, "N": N, "weights": [w for w in weights],
                    "pheno": ["NA" if p == "NA" else p for p in pheno],
                    "presence": [list(map(int, pv)) for pv in presences],
                    "min": mn, "max": mx, "cutoff": cutoff, "omit_B": omit_b, "U": U, "rows": rows})

    # Appendix A9 / A10 / A11 KATs
    W = [0.5, 1.5, 0.7, 1.3, 1.1, 0.9, 1.2, 0.8, 1.0, 1.0]
    run_case("chi2", 10, [1] * 10, [1, 1, 1, 1, 0, 0, 0, 0, 0, "NA"], [[1, 1, 1, 1, 0, 0, 0, 0, 0, 1]], 2, 8, 0.05, True, 100)
    run_case("chi2", 10, W, [1, 1, 1, 1, 0, 0, 0, 0, 0, "NA"], [[1, 1, 1, 0, 1, 0, 0, 0, 0, 1]], 2, 8, 1.1, True, 100)
    cont = [3.0, 4.0, 5.0, 2.0, 1.0, 0.0, -1.0, 0.5, 1.5, "NA"]
    run_case("t", 10, W, cont, [[1, 1, 1, 0, 1, 0, 0, 0, 0, 1]], 2, 8, 1e9, False, 100)
    run_case("t", 10, [1] * 10, cont, [[1, 1, 1, 0, 1, 0, 0, 0, 0, 1]], 2, 8, 1e9, False, 100)
    # random batteries
    for N, weighted in ((12, False), (37, True), (64, False), (100, True), (250, True)):
        ph = [int(x) for x in (rng.random(N) < 0.4)]
        for i in rng.choice(N, size=max(1, N // 15), replace=False):
            ph[int(i)] = "NA"
        w = list(np.round(rng.gamma(2.0, 0.5, N), 6)) if weighted else [1] * N
        pres = (rng.random((60, N)) < rng.random((60, 1))).astype(int)
        pres[:8] = (np.array([0 if p == "NA" else p for p in ph])[None, :] ^ (rng.random((8, N)) < 0.1)).astype(int)
        pres[8] = 0; pres[9] = 1
        run_case("chi2", N, w, ph, pres, 2, N - 2, 0.05, True, 1000)
        run_case("chi2", N, w, ph, pres, 2, N - 2, 0.05, False, 50)
        phc = [float(np.round(x, 3)) for x in rng.normal(0, 2, N)]
        for i in rng.choice(N, size=max(1, N // 15), replace=False):
            phc[int(i)] = "NA"
        base = np.array([0.0 if p == "NA" else p for p in phc])
        pres[:8] = ((base[None, :] + rng.normal(0, 1.0, (8, N))) > 0.5).astype(int)
        run_case("t", N, w, phc, pres, 2, N - 2, 0.05, False, 20)
        run_case("t", N, w, phc, pres, 2, N - 2, 1e9, False, 20)
    # degenerate variance cases for the t-test
    run_case("t", 8, [1] * 8, [2.0, 2.0, 2.0, 5.0, 5.0, 5.0, 5.0, 1.0],
             [[1, 1, 1, 0, 0, 0, 0, 0], [1, 1, 1, 0, 0, 0, 0, 1], [0, 0, 0, 1, 1, 1, 1, 0]], 2, 6, 1e9, False, 5)
    with open(os.path.join(GOLD, "stage3.json"), "w") as f:
        json.dump(out, f)
    print(f"  stage3: {len(out)} cases, {sum(len(c['rows']) for c in out)} rows")


def gen_cli(tag, cfg_idx, extra_args, n_samples=None, genome_len=None):
    """Whole-CLI golden: run the unmodified reference, keep the files the hot path writes."""
    from phenotypeseeker_b200 import synth
    ds = synth.config(cfg_idx, tiny=True, n_samples=n_samples, genome_len=genome_len)
    dst = os.path.join(GOLD, f"cli_{tag}")
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    with tempfile.TemporaryDirectory() as td:
        ph, _ = ds.write(os.path.join(td, "in"))
        args = ["modeling", ph] + extra_args
        code = ("import sys; sys.path.insert(0, %r); from oracle import ref_shim; "
                "ref_shim.run_cli(%r, %r)" % (ROOT, args, td))
        subprocess.run([sys.executable, "-c", code], cwd=td, capture_output=True)
        kept = []
        for fn in sorted(os.listdir(td)):
            if fn.endswith(".tsv") or fn.endswith("_MLdf.csv"):
                shutil.copy(os.path.join(td, fn), os.path.join(dst, fn))
                kept.append(fn)
    with open(os.path.join(dst, "case.json"), "w") as f:
        json.dump({"config": cfg_idx, "tiny": True, "n_samples": n_samples, "genome_len": genome_len,
                   "args": extra_args, "files": kept}, f)
    print(f"  cli_{tag}: {kept}")


def main():
    if not ref_shim.available():
        raise SystemExit("needs /root/reference")
    os.makedirs(GOLD, exist_ok=True)
    build.build_all()
    print("kmer lists (shipped glistmaker|glistquery):")
    gen_kmer_lists()
    gen_union_map()
    gen_gmer_counter()
    print("stage 3 (real conduct_chi_squared_test / conduct_t_test):")
    gen_stage3()
    print("whole CLI:")
    gen_cli("chi2", 0, ["-l", "16", "--omit_B_correction", "-nt", "4", "--n_kmers", "50"],
            n_samples=12, genome_len=20000)
    gen_cli("ttest", 2, ["-l", "13", "--pvalue", "5000", "-nt", "3", "--n_kmers", "50"],
            n_samples=10, genome_len=12000)


if __name__ == "__main__":
    main()
