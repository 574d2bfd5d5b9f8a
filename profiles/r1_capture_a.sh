# round-1 capture A: GPU test suite, default bench line, compute-sanitizer memcheck on the new kernels
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "row_builders or wide_rows or partition_route" > gpurun_out/sanitizer_memcheck_b.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -4 gpurun_out/sanitizer_memcheck_b.log
