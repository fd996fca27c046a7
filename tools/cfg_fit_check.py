"""Scratch: epochs and time of the allocator fits in the shipped SP_Truthful_TS config (18 fits of ~1 667 rows)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import auction_gym_b200 as ag
rng, config, agent_configs, a2i, a2v, num_runs, max_slots, D, var, Do = ag.parse_config("config/SP_Truthful_TS.json")
agents = ag.instantiate_agents(rng, agent_configs, a2v, a2i)
auction, num_iter, T, out = ag.instantiate_auction(rng, config, a2i, a2v, agents, max_slots, D, var, Do, num_runs=num_runs, seed=config["random_seed"], rounds_capacity=config["rounds_per_iter"])
for it in range(4):
    auction.simulate_rounds(T)
    eng = auction.engine
    torch.cuda.synchronize(); t0 = time.time()
    info = eng.update_allocators(want_info=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    inf = info.cpu().numpy()
    print(f"it {it}: fit {dt*1e3:.1f} ms, rows {inf[...,3].mean():.0f}, epochs mean {inf[...,1].mean():.0f} (min {inf[...,1].min():.0f}, max {inf[...,1].max():.0f}), stop {inf[...,0].min():.0f}..{inf[...,0].max():.0f}, loss {inf[...,2].mean():.3f}")
    auction.end_iteration()
