set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "chi2 or stage3 or end_to_end or ranged or select_top" 2>&1 | tail -5
run() { name=$1; shift; timeout 1500 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2_b11_$name.json 2> gpurun_out/r2_b11_$name.err; tail -c 600 gpurun_out/r2_b11_$name.err;
python - $name <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2_b11_{n}.json").read().strip().splitlines()[-1])
    c=d["config"]
    print(n, "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "U", c["union_kmers"], "surv", c["survivors_read_back"], "ranges", c["kmer_ranges"], "value %.3g"%d["value"], "devGB", round(c["device_bytes"]/1e9,1), c["result_digest"], c["digest_check"][:30])
    print("   ", {k:round(v["ms_per_step"],2) for k,v in list(d["kernels"].items())[:12]})
except Exception as e: print(n,"failed",e)
PY
}
run c2 --config 2 --steps 4 --write-digest
run c5s --config 5 --samples 600 --ranges 3 --steps 2
run c5 --config 5 --steps 2 --e2e-steps 1
