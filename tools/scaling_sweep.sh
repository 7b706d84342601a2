#!/bin/bash
# BASELINE.json configs[4]: scaling sweep 10^3 - 10^7 instances at 1/2/4/8 B200, mini_cheetah (stand) and anymal_b (trot).
#   tools/scaling_sweep.sh N          (run under `gpurun --gpus N`)
# N = 1: totals 10^3 .. 10^7 on one GPU. N > 1: strong scaling, 10^6 and 10^7 TOTAL instances split over the N ranks, plus the
# weak-scaling contract line (4096 per GPU). One bench.py JSON line per run -> gpurun_out/r2_scaling_N.jsonl
N=${1:-1}
out=gpurun_out/r2_scaling_$N.jsonl
: > $out
run() {
  if [ "$N" = "1" ]; then python bench.py --gpus 1 --no-cpu "$@" 2>>gpurun_out/r2_scaling_$N.err >> $out
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu "$@" 2>>gpurun_out/r2_scaling_$N.err >> $out; fi
}
if [ "$N" = "1" ]; then
  for robot in "mini_cheetah stand" "anymal_b trot"; do set -- $robot
    for t in 1000 10000 100000 1000000; do run --robot $1 --pattern $2 --batch $t --steps 30; done
    run --robot $1 --pattern $2 --batch 10000000 --steps 5 --warmup 3 --no-e2e
  done
else
  run --steps 100                                                  # weak scaling, contract line
  for robot in "mini_cheetah stand" "anymal_b trot"; do set -- $robot
    run --robot $1 --pattern $2 --total 1000000 --steps 30
    run --robot $1 --pattern $2 --total 10000000 --steps 5 --warmup 3 --no-e2e
  done
fi
python - <<PY
import json
for l in open("$out"):
    d = json.loads(l)
    c = d["config"]
    print("%-13s %-6s gpus %d  per-gpu %9d  total %9s  %8.2f M steps/s  e2e %s  scaling %s" % (c["robot"], c["contact_pattern"], d["n_gpus"], c["instances_per_step_per_gpu"],
          c["total_instances"], d["value"] / 1e6, ("%.2f M" % (d["e2e"]["value"] / 1e6)) if d["e2e"]["value"] else "-", d["scaling"]))
PY
