#!/bin/bash
# Runs bench.py against every alternative build of the library under variants/ (WBC_LIB override),
# at the BASELINE batch (4096) and at a large batch; one line per run into gpurun_out/sweep.txt.
out=gpurun_out/sweep.txt; : > $out
for lib in variants/*.so; do
  for batch in 4096 262144; do
    steps=100; [ $batch -gt 4096 ] && steps=20
    r=$(WBC_LIB=$PWD/$lib python bench.py --no-cpu --steps $steps --warmup 3 --batch $batch 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f M/s  p50 %.4f ms  e2e %.3f M/s' % (d['value']/1e6, d['p50_ms_per_step'], d['e2e']['value']/1e6))")
    echo "$(basename $lib) batch=$batch $r" | tee -a $out
  done
done
