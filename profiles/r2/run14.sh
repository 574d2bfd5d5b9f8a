set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
# 1. the failing drop-in test, with its message
timeout 600 python -m pytest tests/test_dropin_wiring.py tests/test_outputs.py -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_pytest_dropin.log
tail -40 gpurun_out/r2_pytest_dropin.log
# 1b. config 5 with grouped ranges (2 extractions instead of 4)
( time timeout 900 python bench.py --steps 2 --e2e-steps 1 --no-cpu-baseline ) > gpurun_out/r2_bench_c5_grouped.json 2> gpurun_out/r2_bench_c5_grouped.err
tail -c 600 gpurun_out/r2_bench_c5_grouped.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_c5_grouped.json").read().strip().splitlines()[-1]); c=d["config"]
    print("c5 grouped ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "U", c["union_kmers"], "surv", c["survivors_read_back"], "devGB", round(c["device_bytes"]/1e9,1), c["digest_check"][:30])
    print("   ", {k:round(v["ms_per_step"],2) for k,v in list(d["kernels"].items())[:12]})
except Exception as e: print("c5 grouped failed", e)
PY
# 2. launch list of the config-2 step and of a config-5-shaped step (ncu timing pass; numbers printed under ncu are not bench values)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c2.csv \
    python bench.py --config 2 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2_launches_c2.log 2>&1
tail -2 gpurun_out/r2_launches_c2.log | cut -c1-300
# 3. ncu --set full, config 2: one steady-state step of every kernel that matters
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_scatter1|k_scatter2|k_bucket_count_pg|k_bucket_build_pg|k_test_chi2|k_decode_write|k_decode_count" \
    -s 24 -c 8 -o gpurun_out/r2_c2_full python bench.py --config 2 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2_ncu_c2.log 2>&1
tail -3 gpurun_out/r2_ncu_c2.log | cut -c1-300
# 4. config-5 shape (5,000 samples x 10 columns, 640-byte rows) at 400 kbp genomes so that one range fits and ncu can replay
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_scatter1|k_scatter2|k_bucket_build_pg|k_test_chi2" \
    -s 15 -c 5 -o gpurun_out/r2_c5s_full python bench.py --config 5 --genome-len 400000 --ranges 1 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2_ncu_c5s.log 2>&1
tail -3 gpurun_out/r2_ncu_c5s.log | cut -c1-300
# 5. config-3 shape (1,000 samples, Welch) at 1 Mbp genomes
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_test_welch" \
    -s 3 -c 1 -o gpurun_out/r2_c3s_welch python bench.py --config 3 --genome-len 1000000 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2_ncu_c3s.log 2>&1
tail -3 gpurun_out/r2_ncu_c3s.log | cut -c1-300
# 6. the same shapes without ncu (what the kernels take there)
for a in "5 --genome-len 400000 --ranges 1" "3 --genome-len 1000000"; do
  timeout 600 python bench.py --config $a --steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']
print(c['n_samples'], c['genome_len'], 'ms', round(d['ms_per_step'],2), 'U', c['union_kmers'], {k:round(v['ms_per_step'],2) for k,v in list(d['kernels'].items())[:10]})"
done
# 7. the full GPU suite with durations (the two slow oracle comparisons now run on host threads)
( time timeout 900 python -m pytest tests -q -m gpu --durations=12 ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -30 gpurun_out/r2_pytest_gpu.log
ls -la gpurun_out
