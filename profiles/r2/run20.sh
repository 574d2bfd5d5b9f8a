set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variant or row_builders or golden or union_and_matrix or end_to_end or page_route" ) 2>&1 | tail -8
summ() { python - $1 <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
    c=d["config"]
    print(n, "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "U", c["union_kmers"], "surv", c["survivors_read_back"], "value %.3g"%d["value"], c["digest_check"][:30])
    print("   ", {k:round(v["ms_per_step"],2) for k,v in list(d["kernels"].items())[:12]})
except Exception as e: print(n,"failed",e)
PY
}
run() { name=$1; shift; ( time timeout 1500 python bench.py "$@" ) > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err; tail -c 400 gpurun_out/r2_bench_$name.err; summ $name; }
run c2e --config 2 --steps 5 --e2e-steps 2 --no-cpu-baseline
PSKMER_SC1=lean run c2e_lean --config 2 --steps 5 --e2e-steps 2 --no-cpu-baseline
PSKMER_SC1=lean run c5e_lean --config 5 --steps 2 --e2e-steps 1 --no-cpu-baseline
PSKMER_SC1=lean run c3e_lean --config 3 --steps 3 --e2e-steps 1 --no-cpu-baseline
( time timeout 900 python bench.py --impl reference --config 1 --ref-genome-len 4300000 --steps 1 --warmup 0 ) > gpurun_out/r2_bench_reference_c1_full.json 2> gpurun_out/r2_bench_reference_c1_full.err
tail -c 300 gpurun_out/r2_bench_reference_c1_full.err; cut -c1-400 gpurun_out/r2_bench_reference_c1_full.json
