#!/bin/bash
# End-to-end (wbc_step_host, page-locked buffers) under the chunking / split switches; one line per run.
out=gpurun_out/sweep_e2e.txt; : > $out
run() {
  local label=$1; shift
  r=$(env "$@" 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dev %.3f M/s  e2e %.3f M/s' % (d['value']/1e6, d['e2e']['value']/1e6))")
  echo "$label $r" | tee -a $out
}
for batch in 4096 65536 1048576; do
  steps=100; [ $batch -gt 4096 ] && steps=20
  for zc in 1 2 4 8; do
    run "split zc=$zc batch=$batch" WBC_ZC_CHUNKS=$zc python bench.py --no-cpu --steps $steps --warmup 3 --batch $batch
  done
  run "fused batch=$batch" WBC_SPLIT=0 python bench.py --no-cpu --steps $steps --warmup 3 --batch $batch
done
