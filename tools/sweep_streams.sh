#!/bin/bash
# A/B of the number of chunks (= concurrent reduce -> solve chains) a device step is issued as: WBC_TWO_STREAMS=k.
for b in ${BATCHES:-4096 8192 16384}; do
  for k in ${KS:-0 2 3 4}; do
    WBC_TWO_STREAMS=$k python bench.py --no-cpu --no-e2e --no-aux --steps 200 --batch $b 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('chunks %d  batch %7d  value %7.2f M/s  p50 %.4f ms' % ($k, $b, d['value']/1e6, d['p50_ms_per_step']))"
  done
done
