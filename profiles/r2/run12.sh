set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
( time timeout 1500 python -m pytest tests -q -m gpu --durations=15 ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -25 gpurun_out/r2_pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4
summ() { python - $1 <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
    c=d["config"]
    print(n, "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "U", c["union_kmers"], "P", c["phenotype_columns"], "surv", c["survivors_read_back"], "ranges", c["kmer_ranges"], "value %.3g"%d["value"], "gen", round(d["gen_seconds"],1), "devGB", round(c["device_bytes"]/1e9,1), c["digest_check"][:30])
    print("   roofline", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("kernel","achieved","frac","ms_per_launch","whole_step_frac","share_of_kernel_time")})
    print("   ", {k:round(v["ms_per_step"],2) for k,v in list(d["kernels"].items())[:14]})
    print("   cpu", d.get("cpu_baseline",{}).get("value"), d.get("clocks"))
except Exception as e: print(n,"failed",e)
PY
}
run() { name=$1; shift; ( time timeout 1500 python bench.py "$@" ) > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err; tail -c 700 gpurun_out/r2_bench_$name.err; summ $name; }
run c2 --config 2 --steps 5 --write-digest --no-cpu-baseline
run c5_default
run c3 --config 3 --steps 3 --no-cpu-baseline
run c4 --config 4 --steps 3 --no-cpu-baseline
run c1 --config 1 --steps 5 --write-digest --no-cpu-baseline
( time timeout 900 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/r2_bench_reference_c5.json 2> gpurun_out/r2_bench_reference_c5.err; tail -c 600 gpurun_out/r2_bench_reference_c5.err; cat gpurun_out/r2_bench_reference_c5.json | cut -c1-900
ls -la gpurun_out
