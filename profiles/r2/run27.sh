set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
run() { n=$1; name=$2; shift; shift;
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err
grep -v "^W1017\|^\*\*\*\|OMP_NUM" gpurun_out/r2_bench_$name.err | tail -c 600
python - $name <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
    c=d["config"]
    print(n, "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "U", c["union_kmers"], "surv", c["survivors_read_back"], "value %.3g"%d["value"], c["digest_check"][:40], "nvlink", d["roofline"].get("nvlink",{}).get("achieved"))
    print("   ", {k:round(v["ms_per_step"],2) for k,v in list(d["kernels"].items())[:10]})
except Exception as e: print(n,"failed",e)
PY
}
run 4 c5n4 --config 5 --steps 3 --e2e-steps 1
run 4 c2n4 --config 2 --steps 5 --e2e-steps 2
