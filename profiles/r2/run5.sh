set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
for n in 2; do
PS_DIST_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b5_n$n.json 2> gpurun_out/r2_b5_n$n.err
tail -c 1500 gpurun_out/r2_b5_n$n.err
python - $n <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f"gpurun_out/r2_b5_n{n}.json").read().strip().splitlines()[-1])
print("N",n, d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["union_kmers"], d["config"]["survivors"])
for k,v in d["kernels"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
PY
done
