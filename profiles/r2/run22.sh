set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -q -m gpu --durations=5 ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -10 gpurun_out/r2_pytest_gpu.log
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2_pytest_gpu.log | head -30
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4
summ() { python - $1 <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
    c=d["config"]
    print(n, "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "U", c["union_kmers"], "surv", c["survivors_read_back"], "value %.3g"%d["value"], "devGB", round(c["device_bytes"]/1e9,1), c["digest_check"][:30])
    print("   roofline", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("kernel","achieved","frac","ms_per_launch","whole_step_frac","share_of_kernel_time","traffic")})
    print("   ", {k:round(v["ms_per_step"],2) for k,v in list(d["kernels"].items())[:14]})
    print("   cpu", d.get("cpu_baseline",{}).get("value"), d.get("clocks"))
except Exception as e: print(n,"failed",e)
PY
}
run() { name=$1; shift; ( time timeout 1500 python bench.py "$@" ) > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err; tail -c 400 gpurun_out/r2_bench_$name.err; summ $name; }
run final_default
( time timeout 900 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/r2_bench_final_reference.json 2> gpurun_out/r2_bench_final_reference.err; cut -c1-300 gpurun_out/r2_bench_final_reference.json
run final_c2 --config 2 --steps 5 --no-cpu-baseline
run final_c3 --config 3 --steps 3 --no-cpu-baseline
run final_c1 --config 1 --steps 5 --no-cpu-baseline
run final_c4 --config 4 --steps 2 --e2e-steps 1 --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 400 --csv --log-file gpurun_out/r2_launches_c2.csv \
    python bench.py --config 2 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2_launches_c2.log 2>&1
cap() { name=$1; regex=$2; skip=$3; cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none -k regex:"$regex" -s $skip -c $cnt -o gpurun_out/$name python bench.py "$@" --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
  gzip -9f gpurun_out/${name}_source.csv
  rm -f gpurun_out/$name.ncu-rep
}
cap r2_c2_scatter1_lean "k_scatter1" 3 1 --config 2
# DRAM bytes of the dominant kernel at the headline config itself (metrics-only pass: no replay, so config 5 fits)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_scatter1" -s 6 -c 2 --csv \
    --log-file gpurun_out/r2_c5_scatter1_dram.csv python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2_c5_scatter1_dram.log 2>&1
tail -8 gpurun_out/r2_c5_scatter1_dram.csv | cut -c1-300
du -sh gpurun_out
