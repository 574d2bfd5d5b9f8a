set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_scatter1|k_scatter2|k_bucket_count_pg|k_bucket_build_pg" -s 5 -c 5 -o gpurun_out/r2_paged_v1 python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu1.log 2>&1
tail -5 gpurun_out/ncu1.log
ls -la gpurun_out/
