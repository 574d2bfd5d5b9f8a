for v in base MB2; do
  if [ $v = base ]; then unset PSKMER_LIB; else export PSKMER_LIB=/root/repo/phenotypeseeker_b200/_variants/libpskmer_$v.so; fi
  timeout 200 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/bv_$v.json 2> gpurun_out/bv_$v.err
done
