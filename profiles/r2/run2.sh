set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r2_b2_paged.json 2> gpurun_out/r2_b2_paged.err; tail -c 600 gpurun_out/r2_b2_paged.err
python - <<'PY'
import json
for n in ("paged",):
    try:
        d=json.loads(open(f"gpurun_out/r2_b2_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["union_kmers"], d["config"]["survivors"])
        for k,v in d["kernels"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
    except Exception as e: print(n, "failed", e)
PY
