#!/bin/bash
# A/B of the host-buffer path (wbc_step_host on page-locked buffers): WBC_ZC_STAGE (inputs through the copy engine) x WBC_ZC_CHUNKS.
for b in ${BATCHES:-4096 16384 65536}; do
  for s in ${STAGES:-0 1 3}; do
    for k in ${KS:-2 4}; do
      WBC_ZC_STAGE=$s WBC_ZC_CHUNKS=$k python bench.py --no-cpu --no-aux --steps 100 --batch $b 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('stage %d chunks %d  batch %7d  e2e %7.2f M/s  (device %7.2f M/s)' % ($s, $k, $b, d['e2e']['value']/1e6, d['value']/1e6))"
    done
  done
done
