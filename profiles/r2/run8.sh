set -x; mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
run() { name=$1; shift; timeout 1500 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2_b8_$name.json 2> gpurun_out/r2_b8_$name.err; tail -c 900 gpurun_out/r2_b8_$name.err;
python - $name <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2_b8_{n}.json").read().strip().splitlines()[-1])
    c=d["config"]
    print(n, "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "U", c["union_kmers"], "P", c["phenotype_columns"], "surv", c["survivors_read_back"], "ranges", c["kmer_ranges"], "value %.3g"%d["value"], "gen", round(d["gen_seconds"],1), "devGB", round(c["device_bytes"]/1e9,1))
    print("   ", {k:round(v["ms_per_step"],2) for k,v in list(d["kernels"].items())[:12]})
except Exception as e: print(n,"failed",e)
PY
}
( while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader; sleep 5; done ) > gpurun_out/mem.log &
MP=$!
run c5 --config 5 --steps 2 --e2e-steps 1 --write-digest
kill $MP
sort -n gpurun_out/mem.log | tail -1
free -g | head -2
