# round-1 capture B: ncu launch list of the bench command + one --set full capture of the main kernels
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/bench_ncu_c.json 2> gpurun_out/bench_ncu_c.err
timeout 700 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_part_pass|k_bucket_count|k_bucket_build|k_extract_direct|k_decode_write|k_test_chi2" --launch-skip 8 --launch-count 8 -o gpurun_out/r1_kernels3 -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/bench_ncu_d.json 2> gpurun_out/bench_ncu_d.err
tail -2 gpurun_out/bench_ncu_d.err
