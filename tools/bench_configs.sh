#!/bin/bash
# Device-timed throughput of the other BASELINE.json configs (experiments next to the headline line): one summary line each.
run() { python bench.py --no-cpu --steps 50 "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-110s value %7.2f M/s  e2e %7.2f M/s  p50 %.3f ms  iters %.2f' % (d['config']['workload'][:110], d['value']/1e6, d['e2e']['value']/1e6, d['p50_ms_per_step'], d['gpu_details']['mean_active_set_iterations']))"; }
run                                                                           # configs[1]: headline
run --robot anymal_b --pattern trot --batch 16384 --torque-limits             # configs[2]
run --controller clf --pattern walk --batch 65536                             # configs[3], CLF-QP
run --controller pc --pattern walk --batch 65536                              # configs[3], passivity-constrained
run --controller mptc --pattern walk --batch 65536
for b in 1024 16384 131072 1048576; do run --batch $b; done                   # configs[4] points (mini_cheetah stand)
run --robot anymal_b --pattern trot --batch 1048576
