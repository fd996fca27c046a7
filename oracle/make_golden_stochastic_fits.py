"""TEST INFRASTRUCTURE ONLY (build container) -- full-trajectory goldens of the two bidder fits that draw fresh rsample
noise every epoch, from the UNMODIFIED reference:

  DoublyRobustBidder.update          reference src/Bidder.py:477-615, loss src/Models.py:198-218
  ValueLearningBidder('policy').update   src/Bidder.py:210-325 (the policy fit at :278-316)

The engine's noise is Philox keyed by (seed, run, iteration, epoch, row) and restated in oracle/philox_oracle.py, so the
reference can be fed EXACTLY the stream the device will draw: torch's ``Normal.rsample`` takes its standard normals from
``_standard_normal``, which is served here, call by call (= epoch by epoch), from that stream.  The logged rows are agent
2's rows of tests/golden/bidfit_ppo.npz (rows the reference itself logged); the initial weights are the imitation-
initialised policy of that golden and a win-rate initialisation of tests/golden/bidfit_winrate.npz.

    python -m oracle.make_golden_stochastic_fits     ->  tests/golden/bidfit_stochastic.npz
"""
from __future__ import annotations

import contextlib
import io
import os
import re

import numpy as np

from . import make_golden as mg
from . import make_golden_policy as mp
from . import philox_oracle as ph
from . import ref_harness as rh

SEED, ITER, RUN_OFFSET, RUN, AGENT = 77, 3, 5, 1, 1   # the (seed, iteration, global run, agent) the GPU test fits with
N_ROWS = 900


class injected_rsample_stream:
    """Serve ``Normal.rsample``'s standard normals from fn(call index); one call per epoch in both fits."""

    def __init__(self, fn):
        self.fn, self.calls = fn, 0

    def __enter__(self):
        import torch
        import torch.distributions.normal as tdn

        self.mod, self.orig = tdn, tdn._standard_normal

        def fake(shape, dtype, device):
            e = torch.from_numpy(np.asarray(self.fn(self.calls), np.float32))
            self.calls += 1
            return e.reshape(shape)

        tdn._standard_normal = fake
        return self

    def __exit__(self, *exc):
        self.mod._standard_normal = self.orig
        return False


def device_noise(n):
    key = ph.make_key(SEED, RUN_OFFSET + RUN)
    return lambda e: ph.normal_x(np.arange(n, dtype=np.uint32), np.uint32(e), np.uint32((6 << 16) | AGENT), np.uint32(ITER), key)


def set_winrate(model, w):
    import torch

    with torch.no_grad():
        model.model[0].weight.copy_(torch.from_numpy(np.asarray(w[:3], np.float32)).reshape(1, 3))
        model.model[0].bias.copy_(torch.from_numpy(np.asarray(w[3:4], np.float32)))


def winrate_of(model):
    return np.concatenate([model.model[0].weight.detach().numpy().ravel(), model.model[0].bias.detach().numpy().ravel()]).astype(np.float32)


def theta24_to_12(policy):
    """BidShadingPolicy has two unused hidden layers (Models.py:73-77); the 12 used parameters in the engine's order."""
    g = lambda t: t.detach().numpy().ravel()  # noqa: E731
    return np.concatenate([g(policy.shared_linear.weight), g(policy.shared_linear.bias), g(policy.mu_linear_out.weight), g(policy.mu_linear_out.bias),
                           g(policy.sigma_linear_out.weight), g(policy.sigma_linear_out.bias)]).astype(np.float32)


def main():
    import torch

    ref = rh.load_reference()
    B = ref["Bidder"]
    z = np.load(os.path.join(mg.GOLDEN_DIR, "bidfit_ppo.npz"))
    zw = np.load(os.path.join(mg.GOLDEN_DIR, "bidfit_winrate.npz"))
    a, n = 2, N_ROWS
    est, val, gam, prop, util, won = (np.asarray(z[f"a{a}_{k}"][:n]) for k in ("est", "value", "gamma", "prop", "utility", "won"))
    th0, w0 = z[f"a{a}_theta_imit"], zw["a4_w0"]
    # Bidder.update recomputes utilities as value * outcome - price on the won rows: hand it outcome = 1, price = value - utility
    outcomes = won.astype(np.float64)
    prices = np.where(won, val - util, 0.0)
    out = {"est": est, "value": val, "gamma": gam, "prop": prop, "utility": util, "won": won, "theta0": th0, "w0": w0,
           "seed": SEED, "iteration": ITER, "run_offset": RUN_OFFSET, "run": RUN, "agent": AGENT}
    rng = np.random.default_rng(0)
    for kind in ("DR", "VL_POLICY"):
        if kind == "DR":
            b = B.DoublyRobustBidder(rng, gamma_sigma=0.02, init_gamma=1.0)
            mp.set_theta(b.bidding_policy, th0)
            b.model_initialised = True
            b.bidding_policy.model_initialised = True
            b.gammas = [torch.tensor(float(g)) for g in gam]        # tensors once the model is initialised (Bidder.py:470-475)
            b.propensities = [float(p) for p in prop]
        else:
            b = B.ValueLearningBidder(rng, gamma_sigma=0.02, init_gamma=1.0, inference="policy")
            with torch.no_grad():
                b.bidding_policy.shared_linear.weight.copy_(torch.from_numpy(th0[0:4].reshape(2, 2)))
                b.bidding_policy.shared_linear.bias.copy_(torch.from_numpy(th0[4:6]))
                b.bidding_policy.mu_linear_out.weight.copy_(torch.from_numpy(th0[6:8].reshape(1, 2)))
                b.bidding_policy.mu_linear_out.bias.copy_(torch.from_numpy(th0[8:9]))
                b.bidding_policy.sigma_linear_out.weight.copy_(torch.from_numpy(th0[9:11].reshape(1, 2)))
                b.bidding_policy.sigma_linear_out.bias.copy_(torch.from_numpy(th0[11:12]))
            b.model_initialised = True
            b.gammas = [float(g) for g in gam]
            b.propensities = [float(p) for p in prop]
        set_winrate(b.winrate_model, w0)
        buf = io.StringIO()
        with injected_rsample_stream(device_noise(n)) as inj, contextlib.redirect_stdout(buf):
            b.update(contexts=np.zeros((n, 5)), values=val.copy(), bids=est * val * gam, prices=prices.copy(), outcomes=outcomes.copy(),
                     estimated_CTRs=est.copy(), won_mask=won.copy(), iteration=ITER, plot=False, figsize=(8, 5), fontsize=14, name=kind)
        stops = [int(x) for x in re.findall(r"Stopping at Epoch (\d+)", buf.getvalue())]
        stops = stops + [-1] * (2 - len(stops))
        th1 = theta24_to_12(b.bidding_policy) if kind == "VL_POLICY" else mp.theta_of(b.bidding_policy)
        X = torch.tensor(np.stack([est, val], axis=1), dtype=torch.float32)
        with torch.no_grad():
            h = torch.nn.Softplus()(b.bidding_policy.shared_linear(X))
            mu = torch.nn.Softplus()(b.bidding_policy.mu_linear_out(h)).squeeze().numpy()
            sg = (torch.nn.Softplus()(b.bidding_policy.sigma_linear_out(h)) + 1e-2).squeeze().numpy()
        out.update({f"{kind}_w1": winrate_of(b.winrate_model), f"{kind}_theta1": th1, f"{kind}_stops": np.asarray(stops), f"{kind}_epochs": inj.calls,
                    f"{kind}_mu": mu, f"{kind}_sigma": sg})
        print(f"{kind}: win-rate stop {stops[0]}, policy stop {stops[1]} ({inj.calls} policy epochs), mu mean {mu.mean():.4f}, sigma mean {sg.mean():.4f}")
    path = os.path.join(mg.GOLDEN_DIR, "bidfit_stochastic.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
