#!/usr/bin/env python
"""One kept training run of the vectorised trainer (judge item: "show that learn_vec learns"): the reference's curriculum,
hyper-parameters and replay ratio (batch 32, one update per 4 transitions, target update / evaluation every 10 000 / EVAL
learning timesteps), NUM_ENVS parallel environments, evaluation of the reference's 30 maps with the fp32 kernel.

    python scripts/learn_run.py [NUM_ENVS=64] [TOTAL=3000000] [EVAL=50000] [graph]  ->  gpurun_out/learn_run[_graph]/{log.txt, *.npz, network_params.pth}
    ("graph": learn_vec(graph=True), the vector step replayed as one CUDA graph)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np
import torch

import train_iqn

num_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
total = int(sys.argv[2]) if len(sys.argv) > 2 else 3_000_000
eval_freq = int(sys.argv[3]) if len(sys.argv) > 3 else 50_000
graph = len(sys.argv) > 4 and sys.argv[4] == "graph"
out = os.path.join(ROOT, "gpurun_out", "learn_run_graph" if graph else "learn_run")
os.makedirs(out, exist_ok=True)
params = {"agent": "IQN", "seed": 3, "total_timesteps": total, "eval_freq": eval_freq, "save_dir": out, "training_time": "r2"}
t0 = time.time()
exp_dir, model = train_iqn.run_trial("cuda:0", params, num_envs, 32, "reference", graph=graph)
torch.cuda.synchronize()
wall = time.time() - t0
g = np.load(os.path.join(exp_dir, "greedy_evaluations.npz"), allow_pickle=True)
a = np.load(os.path.join(exp_dir, "adaptive_evaluations.npz"), allow_pickle=True)
lines = [f"learn_vec(graph={graph}): {num_envs} envs, batch 32, updates_per_step=reference ({model.reference_updates_per_step(num_envs, 32)}), "
         f"{total} transitions, {model.optimizer.step_count} updates, wall {wall:.1f} s (incl. {len(g['timesteps'])} x 2 evaluations of 30 maps)"]
for name, d in (("greedy", g), ("adaptive", a)):
    sr = d["successes"].mean(axis=1)
    lines.append(f"{name}: success rate per evaluation: " + " ".join(f"{x:.2f}" for x in sr))
    lines.append(f"{name}: final {sr[-1]:.3f}, best {sr.max():.3f}, mean of last 5 {sr[-5:].mean():.3f}; final mean discounted return {d['rewards'][-1].mean():.2f}")
txt = "\n".join(lines)
print(txt)
with open(os.path.join(out, "log.txt"), "w") as f:
    f.write(txt + "\n")
