"""Scratch: run a shipped config with many runs and print mean +- std over runs of the per-iteration results."""
import json, sys, tempfile, os
import numpy as np
sys.path.insert(0, ".")
import auction_gym_b200 as ag
cfg = json.load(open(sys.argv[1])); cfg["num_runs"] = int(sys.argv[2]); cfg["num_iter"] = int(sys.argv[3])
d = tempfile.mkdtemp(); cfg["output_dir"] = d + "/"; p = os.path.join(d, "c.json"); json.dump(cfg, open(p, "w"))
r = ag.run_experiment(p, fit_mode=os.environ.get("FIT_MODE", "adam_ref"))
m, rev = r["metrics"], r["revenue"]
sur, wel, gam = m[..., 0].sum(axis=2), m[..., 1].sum(axis=2), m[..., 9].mean(axis=2)
for i in range(cfg["num_iter"]):
    print(f"iter {i}: revenue {rev[:, i].mean():8.1f} +- {rev[:, i].std():6.1f}   surplus {sur[:, i].mean():8.1f} +- {sur[:, i].std():6.1f}   "
          f"welfare {wel[:, i].mean():8.1f} +- {wel[:, i].std():6.1f}   gamma {np.nanmean(gam[:, i]):.4f}")
