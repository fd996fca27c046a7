"""TEST INFRASTRUCTURE ONLY (build container) -- runs the UNMODIFIED reference's main.py on a config under the shims of
oracle/ref_harness.py and prints per-iteration means of results_*.csv; used to collect distribution-level evidence for the
learning configs (python -m oracle.run_reference_config <config.json> [num_runs] [num_iter] [out_dir])."""
import glob
import json
import os
import runpy
import sys
import tempfile
import time

from . import ref_harness as rh


def main():
    cfg_path = sys.argv[1]
    cfg = json.load(open(cfg_path))
    if len(sys.argv) > 2:
        cfg["num_runs"] = int(sys.argv[2])
    if len(sys.argv) > 3:
        cfg["num_iter"] = int(sys.argv[3])
    out = sys.argv[4] if len(sys.argv) > 4 else tempfile.mkdtemp(prefix="agym_ref_")
    cfg["output_dir"] = out.rstrip("/") + "/"
    tmp = os.path.join(out, "config.json")
    os.makedirs(out, exist_ok=True)
    json.dump(cfg, open(tmp, "w"))
    rh.load_reference()  # installs the shims
    import builtins
    import tqdm as _tq

    _tq.tqdm = lambda it, **k: it
    sys.path.insert(0, rh.REF_SRC)
    sys.argv = ["main.py", tmp]
    t0 = time.time()
    _print = builtins.print
    builtins.print = lambda *a, **k: None
    try:
        runpy.run_path(os.path.join(rh.REF_SRC, "main.py"), run_name="__main__")
    finally:
        builtins.print = _print
    wall = time.time() - t0
    import pandas as pd

    r = pd.read_csv(glob.glob(os.path.join(out, "results_*.csv"))[0])
    p = r.pivot_table(index="Iteration", columns="Measure Name", values="Measure", aggfunc="mean")
    print(f"{os.path.basename(cfg_path)}: reference wall {wall:.0f} s, runs {cfg['num_runs']}, iters {cfg['num_iter']}")
    print(p.round(1).to_string())


if __name__ == "__main__":
    main()
