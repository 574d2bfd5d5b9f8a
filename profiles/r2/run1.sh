set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "union_and_matrix or row_builders or sample_groups or many_samples or range or end_to_end or fastq or golden_kmer or ragged or tile_bound" 2>&1 | tail -25
timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r2_b1_paged.json 2> gpurun_out/r2_b1_paged.err; tail -c 600 gpurun_out/r2_b1_paged.err
PSKMER_PAGED=0 timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r2_b1_old.json 2> gpurun_out/r2_b1_old.err
python - <<'PY'
import json
for n in ("paged","old"):
    try:
        d=json.loads(open(f"gpurun_out/r2_b1_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["union_kmers"], d["config"]["survivors"])
        for k,v in d["kernels"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
    except Exception as e: print(n, "failed", e)
PY
