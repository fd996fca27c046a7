#!/bin/bash
# Runs the six shipped experiment configs end to end on one GPU and records wall time + a result summary.
mkdir -p gpurun_out/configs
for c in SP_Oracle SP_Truthful_TS FP_DM_Oracle FP_DM_TS FP_DR_TS FP_IPS_TS; do
  s=$(date +%s%N)
  python auction_gym_b200/src/main.py config/$c.json --output-dir gpurun_out/configs/$c > gpurun_out/configs/$c.log 2>&1
  rc=$?
  e=$(date +%s%N)
  echo "$c rc=$rc wall=$(( (e - s) / 1000000 )) ms"
done
python - <<'PY'
import glob, pandas as pd
for d in sorted(glob.glob("gpurun_out/configs/*/")):
    f = glob.glob(d + "results_*.csv")
    if not f: continue
    r = pd.read_csv(f[0])
    p = r.pivot_table(index="Iteration", columns="Measure Name", values="Measure", aggfunc="mean")
    last = p.index.max()
    print(d.split("/")[-2], "iter 0:", p.loc[0].round(1).to_dict(), f"iter {last}:", p.loc[last].round(1).to_dict())
PY
