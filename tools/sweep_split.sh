#!/bin/bash
# Launch-bound variants of the library (variants/*.so, WBC_LIB override) at the BASELINE batch and at a large batch,
# plus the other BASELINE configs on the default build; one line per run into gpurun_out/sweep_split.txt.
out=gpurun_out/sweep_split.txt; : > $out
run() {  # label, env..., command
  local label=$1; shift
  r=$(env "$@" 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f M/s  p50 %.4f ms  e2e %.3f M/s launches %d iters %.2f' % (d['value']/1e6, d['p50_ms_per_step'], d['e2e']['value']/1e6, d['gpu_launches'], d['config']['mean_active_set_iterations']))")
  echo "$label $r" | tee -a $out
}
for lib in variants/*.so; do
  for batch in 4096 262144; do
    steps=100; [ $batch -gt 4096 ] && steps=20
    run "$(basename $lib) batch=$batch" WBC_LIB=$PWD/$lib python bench.py --no-cpu --steps $steps --warmup 3 --batch $batch
  done
done
if [ "$1" != "quick" ]; then
run "default batch=1048576" python bench.py --no-cpu --steps 20 --warmup 3 --batch 1048576
run "default clf walk 65536" python bench.py --no-cpu --steps 30 --warmup 3 --batch 65536 --controller clf --pattern walk
run "default pc walk 65536" python bench.py --no-cpu --steps 20 --warmup 3 --batch 65536 --controller pc --pattern walk
run "default anymal trot tl 16384" python bench.py --no-cpu --steps 50 --warmup 3 --batch 16384 --robot anymal_b --pattern trot --torque-limits
fi
