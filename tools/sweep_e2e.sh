#!/bin/bash
# End-to-end (wbc_step_host, page-locked buffers): zero-copy against the staged two-stream pipeline over the batch size, and
# the programmatic-dependent-launch switch on the device-timed step; one line per run.
out=gpurun_out/sweep_e2e.txt; : > $out
run() {
  local label=$1; shift
  r=$(env "$@" 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dev %.3f M/s  p50 %.4f ms  e2e %.3f M/s' % (d['value']/1e6, d['p50_ms_per_step'], d['e2e']['value']/1e6))")
  echo "$label $r" | tee -a $out
}
for pdl in 1 0; do
  run "pdl=$pdl batch=4096" WBC_PDL=$pdl python bench.py --no-cpu --steps 200 --warmup 5
  run "pdl=$pdl batch=1024" WBC_PDL=$pdl python bench.py --no-cpu --steps 200 --warmup 5 --batch 1024
  run "pdl=$pdl batch=65536" WBC_PDL=$pdl python bench.py --no-cpu --steps 30 --warmup 3 --batch 65536
done
for batch in 16384 32768 65536 131072 262144; do
  run "zero-copy batch=$batch" WBC_HOST_ZEROCOPY=1 python bench.py --no-cpu --steps 20 --warmup 3 --batch $batch
  run "staged    batch=$batch" WBC_HOST_ZEROCOPY=0 python bench.py --no-cpu --steps 20 --warmup 3 --batch $batch
done
