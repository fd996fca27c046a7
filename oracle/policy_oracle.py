"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float32) of the bid-shading policy models and their fits.

Restates ``BidShadingContextualBandit`` / ``BidShadingPolicy`` (reference src/Models.py:65-218) and the policy parts of
``PolicyLearningBidder.update`` (src/Bidder.py:369-431), ``DoublyRobustBidder.update`` (src/Bidder.py:557-615) and
``ValueLearningBidder.update`` with inference='policy' (src/Bidder.py:278-316) with hand-written gradients, so the CUDA
kernels (which cannot use autograd) have something to be compared with.  ``tests/test_oracle_golden.py`` checks the
loss values and gradients here against torch autograd of the UNMODIFIED reference (tests/golden/policy_grad.npz) and the
deterministic fits (imitation initialisation, PPO) against the reference's own fitted parameters.

Parameter vector theta (12 floats, the layout of bidder_w[4:16] in include/agym.h):
  W1[0,0], W1[0,1], W1[1,0], W1[1,1], b1[0], b1[1], w_mu[0], w_mu[1], b_mu, w_sigma[0], w_sigma[1], b_sigma
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
MIN_SIGMA = f32(1e-2)          # Models.py:80,104
SQRT_2PI = f32(np.sqrt(2 * np.pi))
BETA1, BETA2, ADAM_EPS = 0.9, 0.999, 1e-8
LOSSES = ("REINFORCE", "REINFORCE_offpolicy", "TRPO", "PPO", "Doubly Robust", "DM")


def softplus(x):
    x = np.asarray(x, f32)
    return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, f32(20))))).astype(f32)


def sigmoid(x):
    x = np.asarray(x, f32)
    return (f32(1) / (f32(1) + np.exp(-x))).astype(f32)


def dsoftplus(x):
    return np.where(np.asarray(x) > 20, f32(1), sigmoid(x)).astype(f32)


def forward(theta, X):
    """Models.py:146-151 (and :82-85): returns a dict with mu, sigma (incl. min_sigma) and the intermediates."""
    th = np.asarray(theta, f32)
    X = np.asarray(X, f32)
    h0 = X[:, 0] * th[0] + X[:, 1] * th[1] + th[4]
    h1 = X[:, 0] * th[2] + X[:, 1] * th[3] + th[5]
    s0, s1 = softplus(h0), softplus(h1)
    a_mu = s0 * th[6] + s1 * th[7] + th[8]
    a_sg = s0 * th[9] + s1 * th[10] + th[11]
    mu = softplus(a_mu)
    sg_raw = softplus(a_sg)
    return dict(h0=h0, h1=h1, s0=s0, s1=s1, a_mu=a_mu, a_sg=a_sg, mu=mu, sg_raw=sg_raw, sigma=(sg_raw + MIN_SIGMA).astype(f32))


def backward(theta, X, fw, d_mu, d_sg):
    """Gradient of sum_rows(L) w.r.t. theta given dL/dmu and dL/dsigma per row."""
    th = np.asarray(theta, f32)
    X = np.asarray(X, f32)
    d_amu = (d_mu * dsoftplus(fw["a_mu"])).astype(f32)
    d_asg = (d_sg * dsoftplus(fw["a_sg"])).astype(f32)
    d_s0 = d_amu * th[6] + d_asg * th[9]
    d_s1 = d_amu * th[7] + d_asg * th[10]
    d_h0 = (d_s0 * dsoftplus(fw["h0"])).astype(f32)
    d_h1 = (d_s1 * dsoftplus(fw["h1"])).astype(f32)
    g = np.array([
        (d_h0 * X[:, 0]).sum(dtype=f32), (d_h0 * X[:, 1]).sum(dtype=f32), (d_h1 * X[:, 0]).sum(dtype=f32), (d_h1 * X[:, 1]).sum(dtype=f32),
        d_h0.sum(dtype=f32), d_h1.sum(dtype=f32),
        (d_amu * fw["s0"]).sum(dtype=f32), (d_amu * fw["s1"]).sum(dtype=f32), d_amu.sum(dtype=f32),
        (d_asg * fw["s0"]).sum(dtype=f32), (d_asg * fw["s1"]).sum(dtype=f32), d_asg.sum(dtype=f32)], f32)
    return g


def imitation_loss_grad(theta, X, gammas):
    """Models.py:122-124: MSE(mu, logged gamma) + MSE(softplus(sigma_out) (no min_sigma), 0.05)."""
    fw = forward(theta, X)
    n = f32(len(gammas))
    e_mu = fw["mu"] - np.asarray(gammas, f32)
    e_sg = fw["sg_raw"] - f32(0.05)
    loss = (e_mu * e_mu).sum(dtype=f32) / n + (e_sg * e_sg).sum(dtype=f32) / n
    return f32(loss), backward(theta, X, fw, (f32(2) * e_mu / n).astype(f32), (f32(2) * e_sg / n).astype(f32))


def winrate(w, est, value, gamma):
    """Models.py:61-62 on features [est, value, gamma]."""
    z = est * f32(w[0]) + value * f32(w[1]) + gamma * f32(w[2]) + f32(w[3])
    return sigmoid(z)


def policy_loss_grad(theta, X, gammas, logging_prop, utility, loss_name, utility_estimates=None, winrate_w=None, eps=None,
                     clip=50.0, kl_weight=5e-2):
    """BidShadingContextualBandit.loss (Models.py:167-218) and the 'DM' objective of Bidder.py:290-303, with gradients.

    eps [n]: the standard normals of ``dist.rsample()`` (Models.py:208 / :87) for the stochastic losses."""
    X = np.asarray(X, f32)
    fw = forward(theta, X)
    mu, sg = fw["mu"], fw["sigma"]
    n = f32(X.shape[0])
    d_mu = np.zeros_like(mu)
    d_sg = np.zeros_like(sg)
    if loss_name != "DM":
        g = np.asarray(gammas, f32)
        u = np.asarray(utility, f32)
        diff = ((mu - g) / sg).astype(f32)
        t_raw = (np.exp(-(diff * diff) / f32(2)) / (sg * SQRT_2PI)).astype(f32)
        live = t_raw > f32(1e-30)                       # torch.clip(min=1e-30): no gradient below the clip
        t = np.maximum(t_raw, f32(1e-30))
        dt_dmu = np.where(live, t * (-(mu - g) / (sg * sg)), f32(0)).astype(f32)
        dt_dsg = np.where(live, t * ((mu - g) ** 2 / (sg ** 3) - f32(1) / sg), f32(0)).astype(f32)
        lp = np.asarray(logging_prop, f32)
        iw = (t / lp).astype(f32)
    if loss_name == "REINFORCE":
        loss = (-(t * u)).sum(dtype=f32) / n
        c = (-u / n).astype(f32)
        d_mu, d_sg = c * dt_dmu, c * dt_dsg
    elif loss_name == "REINFORCE_offpolicy":
        loss = (-(iw * u)).sum(dtype=f32) / n
        c = (-u / (lp * n)).astype(f32)
        d_mu, d_sg = c * dt_dmu, c * dt_dsg
    elif loss_name == "TRPO":
        kl = ((sg * sg + (mu - g) ** 2) / (f32(2) * sg * sg) - f32(0.5)).astype(f32)
        loss = -(iw * u).sum(dtype=f32) / n + kl.sum(dtype=f32) / n * f32(kl_weight)
        c = (-u / (lp * n)).astype(f32)
        dkl_dmu = ((mu - g) / (sg * sg)).astype(f32)
        dkl_dsg = (-((mu - g) ** 2) / (sg ** 3)).astype(f32)
        d_mu = c * dt_dmu + f32(kl_weight) / n * dkl_dmu
        d_sg = c * dt_dsg + f32(kl_weight) / n * dkl_dsg
    elif loss_name == "PPO":
        lo, hi = f32(1.0 / clip), f32(clip)
        cl = np.clip(iw, lo, hi)
        loss = -np.minimum(iw * u, cl * u).sum(dtype=f32) / n
        passes = ((iw >= lo) & (iw <= hi)) | ((iw > hi) & (u < 0)) | ((iw < lo) & (u > 0))
        c = np.where(passes, -u / (lp * n), f32(0)).astype(f32)
        d_mu, d_sg = c * dt_dmu, c * dt_dsg
    elif loss_name in ("Doubly Robust", "DM"):
        e = np.asarray(eps, f32)
        raw = (mu + sg * e).astype(f32)
        inside = (raw > 0) & (raw < 1)
        gs = np.clip(raw, f32(0), f32(1))
        W = winrate(winrate_w, X[:, 0], X[:, 1], gs)
        V = (X[:, 0] * X[:, 1]).astype(f32)
        dm = (W * (V - V * gs)).astype(f32)
        ddm_dgs = (V * (W * (f32(1) - W) * f32(winrate_w[2]) * (f32(1) - gs) - W)).astype(f32)
        ddm_draw = np.where(inside, ddm_dgs, f32(0)).astype(f32)
        d_mu = (-ddm_draw / n).astype(f32)
        d_sg = (-ddm_draw * e / n).astype(f32)
        loss = -dm.sum(dtype=f32) / n
        if loss_name == "Doubly Robust":
            lo, hi = f32(1.0 / clip), f32(clip)
            cl = np.clip(iw, lo, hi)
            du = (u - np.asarray(utility_estimates, f32)).astype(f32)
            loss = loss - (du * cl).sum(dtype=f32) / n
            c = np.where((iw >= lo) & (iw <= hi), -du / (lp * n), f32(0)).astype(f32)
            d_mu = d_mu + c * dt_dmu
            d_sg = d_sg + c * dt_dsg
    else:
        raise ValueError(loss_name)
    return f32(loss), backward(theta, X, fw, d_mu.astype(f32), d_sg.astype(f32))


class AdamAmsgrad:
    """torch.optim.Adam(weight_decay, amsgrad=True), single-tensor path, on one flat float32 vector."""

    def __init__(self, n, lr, weight_decay):
        self.lr, self.wd, self.t = lr, weight_decay, 0
        self.ea, self.es, self.mx = np.zeros(n, f32), np.zeros(n, f32), np.zeros(n, f32)

    def step(self, theta, grad):
        self.t += 1
        g = (grad + f32(self.wd) * theta).astype(f32)
        self.ea += (g - self.ea) * f32(1 - BETA1)
        self.es = (self.es * f32(BETA2) + f32(1 - BETA2) * g * g).astype(f32)
        self.mx = np.maximum(self.mx, self.es)
        bc1, bc2 = 1 - BETA1 ** self.t, 1 - BETA2 ** self.t
        denom = (np.sqrt(self.mx) / f32(np.sqrt(bc2)) + f32(ADAM_EPS)).astype(f32)
        return (theta - f32(self.lr / bc1) * (self.ea / denom)).astype(f32)


class Plateau:
    """ReduceLROnPlateau('min') with the reference's keyword arguments."""

    def __init__(self, opt, patience, factor, min_lr, threshold=1e-4):
        self.opt, self.patience, self.factor, self.min_lr, self.threshold = opt, patience, factor, min_lr, threshold
        self.best, self.bad = np.inf, 0

    def step(self, cur):
        if cur < self.best * (1.0 - self.threshold):
            self.best, self.bad = cur, 0
        else:
            self.bad += 1
        if self.bad > self.patience:
            new_lr = max(self.opt.lr * self.factor, self.min_lr)
            if self.opt.lr - new_lr > 1e-8:
                self.opt.lr = new_lr
            self.bad = 0


def run_fit(theta0, loss_grad, lr, weight_decay, max_epochs, stop_after, plateau=None, noise=None):
    """The reference's training loop skeleton (Bidder.py:394-409 etc.): Adam step, scheduler step, then
    "if best - loss > 1e-6: remember; elif epoch - best_epoch > stop_after: break"."""
    theta = np.array(theta0, f32, copy=True)
    opt = AdamAmsgrad(len(theta), lr, weight_decay)
    sched = Plateau(opt, **plateau) if plateau else None
    best_epoch, best_loss, stop_epoch, losses = -1, np.inf, -1, []
    for epoch in range(max_epochs):
        loss, grad = loss_grad(theta, None if noise is None else noise(epoch))
        theta = opt.step(theta, grad)
        cur = float(loss)
        losses.append(cur)
        if sched:
            sched.step(cur)
        if (best_loss - cur) > 1e-6:
            best_epoch, best_loss = epoch, cur
        elif epoch - best_epoch > stop_after:
            stop_epoch = epoch
            break
    return {"theta": theta, "stop_epoch": stop_epoch, "n_epochs": len(losses), "final_loss": losses[-1], "losses": np.asarray(losses)}


def fit_imitation(theta0, X, gammas, max_epochs=8192 * 2):
    """BidShadingContextualBandit.initialise_policy (Models.py:110-133)."""
    return run_fit(theta0, lambda th, _: imitation_loss_grad(th, X, gammas), lr=1e-3, weight_decay=1e-4, max_epochs=max_epochs,
                   stop_after=512)


def fit_policy_ppo(theta0, X, gammas, logging_prop, utility, loss_name="PPO", max_epochs=8192 * 2):
    """PolicyLearningBidder.update after initialisation (Bidder.py:384-409)."""
    lp = np.maximum(np.asarray(logging_prop, f32), f32(1e-15))  # Bidder.py:385
    return run_fit(theta0, lambda th, _: policy_loss_grad(th, X, gammas, lp, utility, loss_name), lr=2e-3, weight_decay=1e-4,
                   max_epochs=max_epochs, stop_after=512, plateau=dict(patience=100, factor=0.2, min_lr=1e-8))
