set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/r2_b4.json 2> gpurun_out/r2_b4.err; tail -c 300 gpurun_out/r2_b4.err
python - <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r2_b4.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["union_kmers"], d["config"]["survivors"])
for k,v in d["kernels"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
PY
