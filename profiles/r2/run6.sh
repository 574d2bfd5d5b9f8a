set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "select_top or ranged or page_route" 2>&1 | tail -5
run() { name=$1; shift; timeout 900 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2_b6_$name.json 2> gpurun_out/r2_b6_$name.err; tail -c 400 gpurun_out/r2_b6_$name.err;
python - $name <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2_b6_{n}.json").read().strip().splitlines()[-1])
    c=d["config"]
    print(n, "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "U", c["union_kmers"], "P", c["phenotype_columns"], "surv", c["survivors_read_back"], "ranges", c["kmer_ranges"], "value %.3g"%d["value"], "gen", round(d["gen_seconds"],1), "devGB", round(c["device_bytes"]/1e9,1))
    print("   ", {k:round(v["ms_per_step"],2) for k,v in list(d["kernels"].items())[:9]})
except Exception as e: print(n,"failed",e)
PY
}
run c2 --config 2 --steps 3
run c5s --config 5 --samples 600 --ranges 3 --steps 2
run c3s --config 3 --samples 300 --steps 2
run c4s --config 4 --samples 6 --steps 2
run c1 --config 1 --steps 3
