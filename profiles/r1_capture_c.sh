# round-1 final capture: GPU test suite, default bench line, ncu launch list of the bench command
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/bench_ncu_e.json 2> gpurun_out/bench_ncu_e.err
tail -3 gpurun_out/pytest_gpu.log
