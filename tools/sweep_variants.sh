#!/bin/bash
# A/B of build variants of the library on the GPU box: tools/sweep_variants.sh "name:-DFLAG=1 -DOTHER=2" ...
# Each variant is compiled here (cross-compile, before gpurun) into variants/<name>.so when called with BUILD=1, and
# benchmarked on the box (bench.py through WBC_LIB) otherwise. One summary line per variant and batch size.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  if [ -n "$BUILD" ]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -I$ROOT/include $flags \
      $ROOT/quadruped_drake_b200/csrc/wbc_api.cu -o $ROOT/variants/$name.so &
  else
    for b in ${BATCHES:-4096 1048576}; do
      steps=100; [ $b -gt 100000 ] && steps=20
      WBC_LIB=$ROOT/variants/$name.so python $ROOT/bench.py --no-cpu --steps $steps --batch $b ${BENCH_ARGS} 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['roofline'].get('kernels') or {}
print('%-28s batch %8d  value %7.2f M/s  e2e %7.2f M/s  p50 %.4f ms  reduce %.1f us  solve %.1f us' % ('$name', $b, d['value']/1e6, (d['e2e']['value'] or 0)/1e6, d['p50_ms_per_step'], 1e3*k.get('reduce_kernel_ms',0), 1e3*k.get('solve_kernel_ms',0)))"
    done
  fi
done
wait
