set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "chi2 or welch or stage3 or t_pvalue or end_to_end or ranged" 2>&1 | tail -5
run() { name=$1; shift; timeout 1500 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2_b7_$name.json 2> gpurun_out/r2_b7_$name.err; tail -c 600 gpurun_out/r2_b7_$name.err;
python - $name <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2_b7_{n}.json").read().strip().splitlines()[-1])
    c=d["config"]
    print(n, "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "U", c["union_kmers"], "P", c["phenotype_columns"], "surv", c["survivors_read_back"], "ranges", c["kmer_ranges"], "value %.3g"%d["value"], "gen", round(d["gen_seconds"],1), "devGB", round(c["device_bytes"]/1e9,1))
    print("   ", {k:round(v["ms_per_step"],2) for k,v in list(d["kernels"].items())[:10]})
except Exception as e: print(n,"failed",e)
PY
}
run c3s --config 3 --samples 300 --steps 2
run c5s --config 5 --samples 600 --ranges 3 --steps 2
run c3 --config 3 --steps 2 --write-digest
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
run c5 --config 5 --steps 2 --e2e-steps 1 --write-digest
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
free -g | head -2
