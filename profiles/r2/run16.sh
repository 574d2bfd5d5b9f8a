set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -q -m gpu -x --durations=5 ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -12 gpurun_out/r2_pytest_gpu.log
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2_pytest_gpu.log | head -30
summ() { python - $1 <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
    c=d["config"]
    print(n, "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "U", c["union_kmers"], "surv", c["survivors_read_back"], "ranges", c["kmer_ranges"], "value %.3g"%d["value"], "devGB", round(c["device_bytes"]/1e9,1), c["digest_check"][:30])
    print("   roofline", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("kernel","achieved","frac","ms_per_launch","whole_step_frac","share_of_kernel_time")})
    print("   ", {k:round(v["ms_per_step"],2) for k,v in list(d["kernels"].items())[:14]})
except Exception as e: print(n,"failed",e)
PY
}
run() { name=$1; shift; ( time timeout 1500 python bench.py "$@" ) > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err; tail -c 500 gpurun_out/r2_bench_$name.err; summ $name; }
run c5b --config 5 --steps 3 --e2e-steps 2 --no-cpu-baseline
run c2b --config 2 --steps 5 --no-cpu-baseline
run c3b --config 3 --steps 3 --no-cpu-baseline
run c1b --config 1 --steps 5 --no-cpu-baseline
cap() { name=$1; regex=$2; skip=$3; cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none -k regex:"$regex" -s $skip -c $cnt -o gpurun_out/$name python bench.py "$@" --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/$name.log 2>&1
  tail -2 gpurun_out/$name.log | cut -c1-200
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
  gzip -9f gpurun_out/${name}_source.csv
  rm -f gpurun_out/$name.ncu-rep
}
# wide rows at config 3's own size (1,000 x 5 Mbp: the per-(bucket, sample group) work of config 5), new test kernels
cap r2_c3_full "k_bucket_build_pg|k_test_welch|k_scatter2" 12 4 --config 3
cap r2_c5s_chi2 "k_test_chi2" 3 1 --config 5 --genome-len 400000 --ranges 1
du -sh gpurun_out; ls gpurun_out | head -50
