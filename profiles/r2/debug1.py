import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from phenotypeseeker_b200._native import Context
from phenotypeseeker_b200 import synth
from phenotypeseeker_b200.pipeline import unpack_rows, kmer_to_str
from oracle import kmers as ok

def check(name, files, k, ctx):
    ctx.begin(k, len(files)); ctx.add_samples(0, files)
    lists = [ok.count_kmers(f, k) for f in files]
    u = ok.union([l[0] for l in lists])
    U = ctx.build_union()
    gu = ctx.get_union()
    extra = np.setdiff1d(gu, u); missing = np.setdiff1d(u, gu)
    print(f"== {name}: k={k} N={len(files)} U={U} expected={len(u)} extra={len(extra)} missing={len(missing)} sorted={bool(np.all(np.diff(gu.astype(np.int64))>0))}")
    for x in extra[:8]: print("   extra", hex(int(x)), kmer_to_str(x, k))
    for x in missing[:8]: print("   missing", hex(int(x)), kmer_to_str(x, k))
    if len(extra) == 0 and len(missing) == 0:
        pres = ok.presence_matrix(u, lists)
        got = unpack_rows(ctx.get_rows(), len(files))
        bad = np.argwhere(got != pres)
        print("   matrix cells differing:", len(bad), bad[:8].tolist())

ctx = Context(0)
files = [b"", b">only header\n", b">h\nACG\n", b"no records at all\n", b">h\n" + b"ACGT" * 5000,
         b">a\nAC\n>b\nGT\n", b"\n\n>x\n" + b"N" * 4096 + b"ACGTACGTACGTACGTACGT\n"]
check("ragged", files, 16, ctx)
check("one", [b">h\n" + b"ACGT" * 5000], 16, ctx)
check("one_small", [b">h\nACGTACGTACGTACGTACGTTTGACCA\n"], 16, ctx)
ds = synth.config(0, tiny=True)
check("cfg0tiny", ds.files, 16, ctx)
check("cfg0tiny13", ds.files, 13, ctx)
ds = synth.make_dataset(600, genome_len=2500, seed=78, n_clades=5, contigs=(1, 2))
check("n600", ds.files, 13, ctx)
