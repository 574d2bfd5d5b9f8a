set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
PSKMER_TRACE=1 timeout 300 python profiles/r2/e2e_trace.py 2 2> gpurun_out/r2_e2e_trace_default.txt; grep -c "" gpurun_out/r2_e2e_trace_default.txt; grep "^step\|^{" gpurun_out/r2_e2e_trace_default.txt
PSKMER_TRACE=1 PSKMER_SC1=lean timeout 300 python profiles/r2/e2e_trace.py 2 2> gpurun_out/r2_e2e_trace_lean.txt; grep "^step\|^{" gpurun_out/r2_e2e_trace_lean.txt
for v in default lean; do
  if [ $v = lean ]; then export PSKMER_SC1=lean; else unset PSKMER_SC1; fi
  timeout 600 python bench.py --config 2 --steps 5 --e2e-steps 6 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v c2', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
for v in default lean; do
  if [ $v = lean ]; then export PSKMER_SC1=lean; else unset PSKMER_SC1; fi
  timeout 900 python bench.py --config 5 --steps 2 --e2e-steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v c5', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
gzip -9 gpurun_out/r2_e2e_trace_*.txt
