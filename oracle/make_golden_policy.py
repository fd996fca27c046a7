"""TEST INFRASTRUCTURE ONLY -- golden vectors for the bid-shading policy models, from the UNMODIFIED reference.

    python -m oracle.make_golden_policy

Writes
  tests/golden/policy_grad.npz     loss values + torch-autograd gradients of every BidShadingContextualBandit loss
                                   (Models.py:167-218), of initialise_policy's objective (Models.py:122-124) and of the
                                   'policy' objective of ValueLearningBidder (Bidder.py:292-302) at random parameters;
  tests/golden/bidfit_ppo.npz      PolicyLearningBidder.update (Bidder.py:369-431) on rows the reference logged:
                                   parameters before / after initialise_policy / after the PPO fit, stop epochs;
  tests/golden/rounds_fp_bandit.npz  a replayed iteration whose bids come from the fitted bandit (Bidder.py:357-362).
"""
from __future__ import annotations

import contextlib
import io
import os
import re

import numpy as np

from . import auction_oracle as ao
from . import make_golden as mg
from . import ref_harness as rh


def theta_of(model):
    g = lambda t: t.detach().numpy().ravel()  # noqa: E731
    return np.concatenate([g(model.shared_linear.weight), g(model.shared_linear.bias), g(model.mu_linear_out.weight),
                           g(model.mu_linear_out.bias), g(model.sigma_linear_out.weight), g(model.sigma_linear_out.bias)]).astype(np.float32)


def grad_of(model):
    g = lambda t: t.grad.detach().numpy().ravel()  # noqa: E731
    return np.concatenate([g(model.shared_linear.weight), g(model.shared_linear.bias), g(model.mu_linear_out.weight),
                           g(model.mu_linear_out.bias), g(model.sigma_linear_out.weight), g(model.sigma_linear_out.bias)]).astype(np.float32)


def set_theta(model, th):
    import torch

    th = torch.from_numpy(np.asarray(th, np.float32))
    with torch.no_grad():
        model.shared_linear.weight.copy_(th[0:4].reshape(2, 2))
        model.shared_linear.bias.copy_(th[4:6])
        model.mu_linear_out.weight.copy_(th[6:8].reshape(1, 2))
        model.mu_linear_out.bias.copy_(th[8:9])
        model.sigma_linear_out.weight.copy_(th[9:11].reshape(1, 2))
        model.sigma_linear_out.bias.copy_(th[11:12])


class injected_rsample_noise:
    """torch.distributions.Normal.rsample draws eps through ``_standard_normal``; serve it from a queue."""

    def __init__(self, queue):
        self.queue = list(queue)

    def __enter__(self):
        import torch
        import torch.distributions.normal as tdn

        self.mod, self.orig = tdn, tdn._standard_normal

        def fake(shape, dtype, device):
            e = torch.from_numpy(np.asarray(self.queue.pop(0), np.float32))
            return e.reshape(shape)

        tdn._standard_normal = fake
        return self

    def __exit__(self, *exc):
        self.mod._standard_normal = self.orig
        return False


def make_grad_golden():
    import torch

    ref = rh.load_reference()
    M = ref["Models"]
    rng = np.random.default_rng(51)
    n = 257
    out = {}
    X = np.stack([rng.uniform(0.02, 0.4, n), rng.lognormal(0.1, 0.2, n)], axis=1).astype(np.float32)
    gam = rng.uniform(0.2, 1.05, n).astype(np.float32)
    lp = rng.uniform(0.05, 20.0, n).astype(np.float32)
    lp[:5] = 1e-15
    u = (rng.normal(0.0, 0.3, n) * (rng.random(n) < 0.5)).astype(np.float32)
    uh = rng.normal(0.0, 0.1, n).astype(np.float32)
    ww = np.array([2.5, 0.9, 5.0, -7.0], np.float32)
    out.update(X=X, gammas=gam, logging_prop=lp, utility=u, utility_estimates=uh, winrate_w=ww)
    wr = M.PyTorchWinRateEstimator()
    with torch.no_grad():
        wr.model[0].weight.copy_(torch.from_numpy(ww[:3]).reshape(1, 3))
        wr.model[0].bias.copy_(torch.from_numpy(ww[3:]))
    tX, tg, tlp, tu, tuh = (torch.from_numpy(v) for v in (X, gam, lp, u, uh))
    for case in range(4):
        th = rng.uniform(-0.9, 0.9, 12).astype(np.float32)
        if case == 3:
            th[8] = 3.0  # a large mu: exercises the clip of the sampled gamma at 1
        eps = rng.standard_normal(n).astype(np.float32)
        out[f"c{case}_theta"], out[f"c{case}_eps"] = th, eps
        for name in ("REINFORCE", "REINFORCE_offpolicy", "TRPO", "PPO", "Doubly Robust"):
            model = M.BidShadingContextualBandit(loss=name)
            set_theta(model, th)
            with injected_rsample_noise([eps]):
                loss = model.loss(tX, tg, tlp, tu, utility_estimates=tuh, winrate_model=wr, importance_weight_clipping_eps=50.0)
            loss.backward()
            key = name.replace(" ", "_")
            out[f"c{case}_{key}_loss"], out[f"c{case}_{key}_grad"] = np.float32(loss.item()), grad_of(model)
        # initialise_policy objective (Models.py:122-124)
        model = M.BidShadingContextualBandit(loss="PPO")
        set_theta(model, th)
        sp = torch.nn.Softplus()
        pm = sp(model.mu_linear_out(sp(model.shared_linear(tX))))
        ps = sp(model.sigma_linear_out(sp(model.shared_linear(tX))))
        crit = torch.nn.MSELoss()
        loss = crit(pm.squeeze(), tg) + crit(ps.squeeze(), torch.ones_like(tg) * .05)
        loss.backward()
        out[f"c{case}_imitation_loss"], out[f"c{case}_imitation_grad"] = np.float32(loss.item()), grad_of(model)
        # ValueLearningBidder 'policy' objective (Bidder.py:292-302) on BidShadingPolicy
        pol = M.BidShadingPolicy()
        set_theta(pol, th)
        with injected_rsample_noise([eps]):
            sg_, _ = pol(tX)
        Xg = torch.hstack((tX, sg_))
        pw = wr(Xg).squeeze()
        values = Xg[:, 0].squeeze() * Xg[:, 1].squeeze()
        loss = -(pw * (values - values * sg_.squeeze())).mean()
        loss.backward()
        out[f"c{case}_DM_loss"], out[f"c{case}_DM_grad"] = np.float32(loss.item()), grad_of(pol)
    path = os.path.join(mg.GOLDEN_DIR, "policy_grad.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def make_ppo_golden():
    import torch

    ref = rh.load_reference()
    O, GA = ao.ALLOC_ORACLE, ao.BID_GAUSS
    kw = dict(seed=61, A=3, n_items=12, D=5, Do=4, P=2, mechanism=ao.MECH_FIRST, alloc_kinds=[O] * 3, bidder_kinds=[GA] * 3, T=3000)
    case, noise, cfg = mg.build_case(**kw)
    for ac in cfg["agents"]:
        ac["bidder"] = {"type": "PolicyLearningBidder", "kwargs": {"gamma_sigma": 0.02, "init_gamma": 1.0, "loss": "\"PPO\""}}
    torch.manual_seed(11)
    rec, met, auction, agents = mg.run_reference(case, noise, cfg)
    out = {}
    thetas = np.zeros((3, 12), np.float32)
    for a, ag in enumerate(agents):
        b = ag.bidder
        won = np.array([o.won for o in ag.logs], bool)
        est = np.array([o.estimated_CTR for o in ag.logs])
        val = np.array([o.value for o in ag.logs])
        price = np.array([o.price for o in ag.logs])
        outc = np.array([o.outcome for o in ag.logs])
        util = np.zeros_like(val)
        util[won] = val[won] * outc[won] - price[won]
        pre = f"a{a}_"
        out[pre + "est"], out[pre + "value"], out[pre + "gamma"] = est, val, np.array(b.gammas)
        out[pre + "prop"], out[pre + "utility"], out[pre + "won"] = np.array(b.propensities), util, won
        out[pre + "theta0"] = theta_of(b.model)
        snap = {}
        orig_init = b.model.initialise_policy

        def wrapped(ctx, gam, _o=orig_init, _m=b.model, _s=snap):
            _o(ctx, gam)
            _s["theta"] = theta_of(_m)

        b.model.initialise_policy = wrapped
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ag.update(iteration=0)
        stops = [int(x) for x in re.findall(r"Stopping at Epoch (\d+)", buf.getvalue())]
        out[pre + "theta_imit"], out[pre + "theta1"] = snap["theta"], theta_of(b.model)
        out[pre + "stops"] = np.asarray(stops + [-1] * (2 - len(stops)))
        thetas[a] = theta_of(b.model)
        print(f"ppo fit agent {a}: rows {len(won)}, stops {stops}, theta1 {thetas[a][:4]}...")
    path = os.path.join(mg.GOLDEN_DIR, "bidfit_ppo.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")
    # second iteration: bids from the fitted bandit, replayed (rsample noise injected per bid)
    for ag in agents:
        ag.clear_utility()
        ag.clear_logs()
    auction.clear_revenue()
    T2 = 400
    rng2 = np.random.default_rng(6161)
    noise2 = ao.draw_replay_inputs(rng2, T2, 3, 2, 5, want_gamma=True)
    rr = rh.ReplayRNG(noise2["ctx"], noise2["parts"], noise2["u"], noise2["gamma_z"])
    auction.rng = rr
    for ag in agents:
        ag.bid = type(ag).bid.__get__(ag)
        ag.bidder.rng = rr
    rh.wrap_bid_slots(agents, rr)
    queue = [np.array([z]) for z in noise2["gamma_z"].ravel()]  # one draw per (round, slot), in bid order
    with injected_rsample_noise(queue):
        rec2 = rh.run_reference_rounds(auction, agents, rr, T2, None)
    met2 = rh.reference_metrics(auction, agents)
    case2 = dict(case)
    case2["bidder_kind"] = np.full(3, ao.BID_BANDIT, np.int32)
    case2["policy_w"] = thetas
    # gammas / propensities are torch tensors on this path (Bidder.py:365-366): read them back as floats
    mg.save_case("rounds_fp_bandit", case2, noise2, rec2, met2)


if __name__ == "__main__":
    make_grad_golden()
    make_ppo_golden()
