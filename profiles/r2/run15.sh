set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
# 1. the full GPU suite
( time timeout 900 python -m pytest tests -q -m gpu --durations=8 ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -14 gpurun_out/r2_pytest_gpu.log
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2_pytest_gpu.log | head -40
# 2. launch lists (ncu timing pass; numbers printed under ncu are not bench values)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c2.csv \
    python bench.py --config 2 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2_launches_c2.log 2>&1
# 3. ncu --set full; reports are turned into text/CSV pages here (the .ncu-rep files are too large to travel)
cap() { name=$1; regex=$2; skip=$3; cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none -k regex:"$regex" -s $skip -c $cnt -o gpurun_out/$name python bench.py "$@" --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/$name.log 2>&1
  tail -2 gpurun_out/$name.log | cut -c1-200
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
  gzip -9 gpurun_out/${name}_source.csv
  rm -f gpurun_out/$name.ncu-rep
}
cap r2_c2_full "k_scatter1|k_scatter2|k_bucket_count_pg|k_bucket_build_pg|k_test_chi2|k_decode_write|k_decode_count" 24 8 --config 2
cap r2_c5s_full "k_scatter1|k_scatter2|k_bucket_build_pg|k_test_chi2" 15 5 --config 5 --genome-len 400000 --ranges 1
cap r2_c3s_welch "k_test_welch" 3 1 --config 3 --genome-len 1000000
# a range-restricted extraction (what config 5 runs on one GPU: half of the k-mers kept per pass)
cap r2_c5r_scatter1 "k_scatter1" 6 2 --config 5 --genome-len 400000 --ranges 4
du -sh gpurun_out; ls -la gpurun_out
