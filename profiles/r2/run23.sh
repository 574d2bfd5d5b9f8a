set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variant or row_builders or golden or union_and_matrix or end_to_end or page_route or ranged" ) 2>&1 | tail -3
for a in "2 --steps 6" "5 --steps 2 --e2e-steps 1" "3 --steps 3 --e2e-steps 1"; do
  timeout 900 python bench.py --config $a --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']
print(c['n_samples'], 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), c['digest_check'][:20], {k:round(v['ms_per_step'],2) for k,v in list(d['kernels'].items())[:6]})"
done
