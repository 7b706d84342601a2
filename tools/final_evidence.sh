#!/bin/bash
# Round evidence in one GPU call: launch list + full ncu capture of the headline step, bench lines (headline, reference arm,
# other BASELINE configs, rows next to the step). Everything lands in gpurun_out/ and is copied to profiles/ afterwards.
o=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r2_launches_final5.csv python bench.py --steps 5 --warmup 3 --no-cpu > $o/launches_bench.log 2>&1
WBC_TWO_STREAMS=0 ncu --set full --import-source on --clock-control none -k regex:"wbc_(reduce|solve)_kernel" --launch-skip 6 -c 2 -f -o $o/r2_step_final5 python bench.py --no-cpu --steps 2 --warmup 3 > $o/ncu_final5.log 2>&1
python bench.py --steps 200 --warmup 5 > $o/r2_bench_final5_n4096.json 2>$o/bench_final5.err
python bench.py --impl reference --steps 5 --warmup 1 > $o/r2_bench_reference_arm5.json 2>>$o/bench_final5.err
bash tools/bench_configs.sh > $o/r2_bench_configs5.txt 2>>$o/bench_final5.err
for w in wire traj rollout; do python bench.py --workload $w --steps 20 --warmup 3 2>>$o/bench_final5.err; done > $o/r2_bench_aux5.jsonl
tail -c 300 $o/r2_bench_final5_n4096.json; echo; cat $o/r2_bench_configs5.txt
