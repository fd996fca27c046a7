"""Scratch: group an `ncu --page source --csv` export into contiguous SASS regions with similar execution counts."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
norm = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0   # fits * epochs, to print executions per fit-epoch
hdr, data = rows[1], rows[2:]
isrc, iex, ith, ism = (hdr.index(k) for k in ("Source", "Instructions Executed", "Avg. Threads Executed", "# Samples"))
tot = sum(int(r[iex]) for r in data if r[iex].isdigit())
tsmp = sum(int(r[ism]) for r in data if r[ism].isdigit())
seg, cur = [], None
for i, r in enumerate(data):
    if not r[iex].isdigit():
        continue
    e = int(r[iex])
    if e < tot * 2e-5:
        if cur: seg.append(cur); cur = None
        continue
    if cur and abs(e - cur["e"]) <= 0.15 * cur["e"]:
        cur["n"] += 1; cur["sum"] += e; cur["end"] = i; cur["smp"] += int(r[ism]); cur["thr"] += float(r[ith])
    else:
        if cur: seg.append(cur)
        cur = {"start": i, "end": i, "e": e, "n": 1, "sum": e, "smp": int(r[ism]), "thr": float(r[ith])}
if cur: seg.append(cur)
print("total warp instructions", tot, "per unit", tot / norm)
for s in seg:
    if s["sum"] > tot * 0.004:
        ops = {}
        for r in data[s["start"]:s["end"] + 1]:
            op = r[isrc].split()[0] if not r[isrc].lstrip().startswith("@") else r[isrc].split()[1]
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
        print(f"{s['start']:5d}-{s['end']:5d} n={s['n']:4d} x{s['e'] / norm:7.3f} instr/unit={s['sum'] / norm:7.1f} share={s['sum'] / tot * 100:5.1f}% "
              f"samples={s['smp'] / tsmp * 100:5.1f}% thr={s['thr'] / s['n']:5.1f}  {top}")
