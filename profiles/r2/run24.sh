set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "ranged or page_route or end_to_end or union_and_matrix" ) 2>&1 | tail -5
for a in "5 --steps 2 --e2e-steps 3" "2 --steps 5 --e2e-steps 5"; do
  timeout 900 python bench.py --config $a --no-cpu-baseline 2>gpurun_out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']
print(c['n_samples'], 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), c['digest_check'][:20], 'launches', d['gpu_launches'], {k:round(v['ms_per_step'],2) for k,v in list(d['kernels'].items())[:6]})"
  tail -c 300 gpurun_out/err.txt
done
