"""Quick per-kernel timing probe (CUDA events inside libpskmer): `python profiles/perf_probe.py N L`."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from phenotypeseeker_b200 import synth
from phenotypeseeker_b200.pipeline import KmerAssociation


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 4_300_000
    weighted = n >= 100
    t0 = time.time()
    ds = synth.make_dataset(n, genome_len=L, seed=20260102, weighted=weighted, n_clades=16, pos_rate=0.35)
    print(f"generated {n} x {L} in {time.time() - t0:.1f}s, {ds.total_bytes() / 1e6:.1f} MB", flush=True)
    ka = KmerAssociation(device=0)
    for rep in range(3):
        ka.ctx.profile(rep == 2)
        t0 = time.time()
        ka.count(ds.files, int(os.environ.get("PS_K", "16")))
        t1 = time.time()
        U = ka.build()
        t2 = time.time()
        res = ka.test(ds.pheno, True, ds.weights if weighted else None, max_samples=n - 2,
                      pvalue_cutoff=0.05, omit_b=(n < 40))
        t3 = time.time()
        print(f"rep{rep}: count {1e3 * (t1 - t0):.1f} ms  build {1e3 * (t2 - t1):.1f} ms  test+fetch {1e3 * (t3 - t2):.1f} ms"
              f"  U={U} survivors={len(res[0].kmer)} kmers/s={U / (t3 - t0):.3e}", flush=True)
    tab = ka.ctx.profile_table()
    tot = sum(v["ms"] for v in tab.values())
    for k, v in sorted(tab.items(), key=lambda kv: -kv[1]["ms"]):
        gbs = v["alg_bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0
        print(f"{k:16s} launches={v['launches']:5d} ms={v['ms']:9.3f} share={v['ms'] / tot:6.1%} alg_GB/s={gbs:8.1f}")
    print(f"device bytes held: {ka.ctx.device_bytes() / 1e9:.2f} GB; launches: {ka.ctx.launch_count()}")


if __name__ == "__main__":
    main()
