#!/bin/bash
# End-to-end (wbc_step_host, page-locked buffers) and device rate under the input-staging switch; one line per run.
out=gpurun_out/sweep_e2e.txt; : > $out
run() {
  local label=$1; shift
  r=$(env "$@" 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dev %.3f M/s  p50 %.4f ms  e2e %.3f M/s' % (d['value']/1e6, d['p50_ms_per_step'], d['e2e']['value']/1e6))")
  echo "$label $r" | tee -a $out
}
for batch in 4096 65536 1048576; do
  steps=100; [ $batch -gt 4096 ] && steps=20
  for b in 1 0; do
    run "bulk_in=$b batch=$batch" WBC_BULK_IN=$b python bench.py --no-cpu --steps $steps --warmup 3 --batch $batch
  done
done
run "bulk_in=1 zc=2 batch=4096" WBC_ZC_CHUNKS=2 python bench.py --no-cpu --steps 100 --warmup 3 --batch 4096
run "bulk_in=1 zc=4 batch=1048576" WBC_ZC_CHUNKS=4 python bench.py --no-cpu --steps 20 --warmup 3 --batch 1048576
