#!/usr/bin/env python
"""Throughput of the closed-loop rollout with the ground-contact plant (4096 / 65536 robots), for A/B of plant-kernel variants."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench_aux
for d in bench_aux.collect("rollout", rollout_sizes=((4096, 100), (65536, 20))):
    print("%-60s plant=%d  %.2f M robot-steps/s  (no graph %.2f M)" % (d["config"]["workload"][:60], d["plant"], d["value"] / 1e6, d["without_graph_robot_steps_per_s"] / 1e6))
