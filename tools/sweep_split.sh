#!/bin/bash
# Split (reduce + solve kernels) vs fused step, and launch-bound variants of both halves (variants/*.so, WBC_LIB override).
out=gpurun_out/sweep_split.txt; : > $out
run() {  # label, env..., args
  local label=$1; shift
  r=$(env "$@" 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f M/s  p50 %.4f ms  e2e %.3f M/s launches %d' % (d['value']/1e6, d['p50_ms_per_step'], d['e2e']['value']/1e6, d['gpu_launches']))")
  echo "$label $r" | tee -a $out
}
for lib in variants/*.so; do
  for batch in 4096 262144; do
    steps=100; [ $batch -gt 4096 ] && steps=20
    run "$(basename $lib) split batch=$batch" WBC_LIB=$PWD/$lib python bench.py --no-cpu --steps $steps --warmup 3 --batch $batch
  done
done
for batch in 4096 262144 1048576; do
  steps=100; [ $batch -gt 4096 ] && steps=20
  run "base FUSED batch=$batch" WBC_SPLIT=0 python bench.py --no-cpu --steps $steps --warmup 3 --batch $batch
done
run "base split batch=1048576" python bench.py --no-cpu --steps 20 --warmup 3 --batch 1048576
run "base split clf walk 65536" python bench.py --no-cpu --steps 30 --warmup 3 --batch 65536 --controller clf --pattern walk
run "base FUSED clf walk 65536" WBC_SPLIT=0 python bench.py --no-cpu --steps 30 --warmup 3 --batch 65536 --controller clf --pattern walk
run "base split anymal trot tl 16384" python bench.py --no-cpu --steps 50 --warmup 3 --batch 16384 --robot anymal_b --pattern trot --torque-limits
run "base FUSED anymal trot tl 16384" WBC_SPLIT=0 python bench.py --no-cpu --steps 50 --warmup 3 --batch 16384 --robot anymal_b --pattern trot --torque-limits
