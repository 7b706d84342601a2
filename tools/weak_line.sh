#!/bin/bash
# Weak-scaling contract line (4096 instances per GPU) at N = 2, 4, 8 ranks of one box + the 10^6-total strong line at 8:
#   gpurun --gpus 8 -- bash tools/weak_line.sh     -> gpurun_out/r2_weak_line.jsonl
out=gpurun_out/r2_weak_line.jsonl; : > $out
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --no-cpu --no-aux --steps 100 2>>gpurun_out/r2_weak_line.err >> $out
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --no-cpu --no-aux --total 1000000 --steps 30 2>>gpurun_out/r2_weak_line.err >> $out
python - <<PY
import json
for l in open("$out"):
    d = json.loads(l); c = d["config"]
    print("gpus %d per-gpu %8d  %8.2f M steps/s  e2e %.2f M  scaling %s" % (d["n_gpus"], c["instances_per_step_per_gpu"], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["scaling"]))
PY
