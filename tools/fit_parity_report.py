"""Scratch: how far the CUDA allocator fit lands from the reference-run goldens (tests/golden/fit_*.npz) and from the
fit oracle -- the numbers behind the tolerances written in tests/test_gpu_fit.py."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import fit_oracle as fo
from tests import test_gpu_fit as tf
from tests.conftest import GOLDEN_DIR
for name in ("fit_ref_shape", "fit_64x64"):
    z = np.load(f"{GOLDEN_DIR}/{name}.npz")
    agents = [int(a) for a in z["fit_agents"]]
    for it in (0, 1):
        pre = [f"it{it}_a{a}_" for a in agents]
        I, K = z[pre[0] + "m0"].shape
        Do = K - 1
        rows = []
        for j, p in enumerate(pre):
            X, items, y = z[p + "X"], z[p + "items"], z[p + "y"]
            rows += [(r, j, X[r, :Do], items[r], y[r]) for r in range(len(y))]
        rows.sort(key=lambda t: (t[0], t[1]))
        T = len(rows)
        for mode in (0, 1):
            eng = tf._engine_for_fits(None, len(agents), I, Do, T)
            eng.fit_ctx[0, :T].copy_(torch.from_numpy(np.stack([r[2] for r in rows]).astype(np.float32)))
            meta = tf._pack_meta(np.array([r[1] for r in rows]), np.array([r[3] for r in rows]), np.array([r[4] for r in rows]) > 0)
            eng.fit_meta[0, :T].copy_(torch.from_numpy(meta.view(np.int32)))
            eng._check(eng.lib.agym_set_rounds_in_iteration(eng.handle, T))
            eng.set_allocator_state(np.stack([z[p + "m0"] for p in pre])[None], np.stack([z[p + "q0"] for p in pre])[None],
                                    np.stack([z[p + "m_prev"] for p in pre])[None])
            info = eng.update_allocators(fit_mode=mode).cpu().numpy()[0]
            m1, q1 = eng.m.cpu().numpy()[0], eng.q.cpu().numpy()[0]
            for j, p in enumerate(pre):
                xs = np.concatenate([np.random.default_rng(1).standard_normal((256, Do)), np.ones((256, 1))], axis=1).astype(np.float32)
                used = np.unique(z[p + "items"])
                est_c = 1 / (1 + np.exp(-(xs @ m1[j].T)))[:, used]
                est_r = 1 / (1 + np.exp(-(xs @ z[p + "m1"].T)))[:, used]
                orc = fo.fit_allocator(z[p + "X"], z[p + "items"], z[p + "y"], z[p + "m0"], z[p + "q0"], z[p + "m_prev"]) if mode == 0 else None
                line = (f"{name} it{it} mode{mode} agent {agents[j]:2d} rows {len(z[p + 'y']):5d}  stop cuda {int(info[j, 0]):6d} ref {int(z[p + 'stop_epoch']):6d}"
                        f"  |dm| {np.abs(m1[j] - z[p + 'm1']).max():.2e}  q rel {np.abs(q1[j] / z[p + 'q1'] - 1).max():.2e}"
                        f"  map-ctr rel {np.abs(est_c / est_r - 1).max():.2e}  loss rel {abs(info[j, 2] / z[p + 'losses_tail'][-1] - 1):.1e}")
                if orc is not None:
                    est_o = 1 / (1 + np.exp(-(xs @ orc["m"].T)))[:, used]
                    line += (f"  || oracle-vs-ref: stop {orc['stop_epoch']} |dm| {np.abs(orc['m'] - z[p + 'm1']).max():.2e} q {np.abs(orc['q'] / z[p + 'q1'] - 1).max():.2e}"
                             f" ctr {np.abs(est_o / est_r - 1).max():.2e}")
                print(line, flush=True)
            eng.close()
