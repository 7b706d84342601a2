#!/bin/bash
out=gpurun_out/sweep_e2e2.txt; : > $out
run() {
  local label=$1; shift
  r=$(env "$@" 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dev %.3f M/s  p50 %.4f ms  e2e %.3f M/s' % (d['value']/1e6, d['p50_ms_per_step'], d['e2e']['value']/1e6))")
  echo "$label $r" | tee -a $out
}
for batch in 4096 65536 1048576; do
  steps=50; [ $batch -gt 4096 ] && steps=10
  for ch in 1 2 4 8 16; do
    run "staged chunks=$ch batch=$batch" WBC_HOST_ZEROCOPY=0 WBC_HOST_CHUNKS=$ch python bench.py --no-cpu --steps $steps --warmup 3 --batch $batch
  done
done
