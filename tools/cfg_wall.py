"""Scratch: in-process wall time of a shipped config through run_experiment (second and third of three runs)."""
import sys, time
sys.path.insert(0, ".")
import torch
import auction_gym_b200 as ag
cfg = sys.argv[1]
w = []
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ag.run_experiment(cfg); torch.cuda.synchronize(); w.append(round(time.perf_counter() - t0, 3))
print(cfg, w)
