set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "union_and_matrix or row_builders or sample_groups or many_samples or range or end_to_end or fastq or ragged or tile_bound" 2>&1 | tail -5
for v in ""; do
env $v timeout 300 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/r2_b3.json 2> gpurun_out/r2_b3.err; tail -c 300 gpurun_out/r2_b3.err
python - "$v" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r2_b3.json").read().strip().splitlines()[-1])
print(sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["union_kmers"], d["config"]["survivors"])
for k,v in d["kernels"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
PY
done
