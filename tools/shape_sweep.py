"""Scratch: sweep seeds of tests/test_gpu_shapes.py::test_odd_shapes_replay_and_fit to find failing cases."""
import sys, traceback
sys.path.insert(0, ".")
import tests.test_gpu_shapes as ts
orig = ts._case
for name in (sys.argv[2:] or ("D12_Do9_P5", "D3_Do1_P17", "D20_Do20_P2")):
    for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
        ts._case = lambda s, **kw: orig(seed, **kw)
        try:
            ts.test_odd_shapes_replay_and_fit(name)
        except Exception as e:
            msg = str(e).strip().splitlines()
            tb = traceback.extract_tb(e.__traceback__)
            where = "; ".join(f"{f.name}:{f.lineno} {f.line}" for f in tb[-2:])
            print(name, seed, type(e).__name__, where[:300], " | ".join(msg[:6])[:600], flush=True)
print("sweep done")
