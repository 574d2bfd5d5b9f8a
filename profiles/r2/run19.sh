set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -q -m gpu -x --durations=3 ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu.log
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2_pytest_gpu.log | head -30
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
run c2d --config 2 --steps 5 --no-cpu-baseline
run c5d --config 5 --steps 2 --e2e-steps 1 --no-cpu-baseline
PSKMER_BK_ROW_KB=56 run c5d_row56 --config 5 --steps 2 --e2e-steps 1 --no-cpu-baseline
PSKMER_BK_ROW_KB=96 run c5d_row96 --config 5 --steps 2 --e2e-steps 1 --no-cpu-baseline
PSKMER_BK_ROW_KB=56 run c2d_row56 --config 2 --steps 5 --e2e-steps 1 --no-cpu-baseline
PSKMER_DECODE=bytes run c2d_bytes --config 2 --steps 5 --e2e-steps 1 --no-cpu-baseline
run c3d --config 3 --steps 3 --e2e-steps 1 --no-cpu-baseline
du -sh gpurun_out
