"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float32) of the per-iteration allocator fit.

Restates ``PyTorchLogisticRegressionAllocator.update`` (reference src/BidderAllocation.py:29-65)
together with ``PyTorchLogisticRegression.{predict_item,loss,laplace_approx,update_prior}``
(src/Models.py:35-48), ``torch.optim.Adam`` (single-tensor path, no amsgrad / weight decay) and
``torch.optim.lr_scheduler.ReduceLROnPlateau('min', factor=0.5)`` as one explicit state machine.
torch 1.13.1 is the reference's pin (requirements.txt); torch 2.11.0 is what the golden vectors in
``tests/golden/fit_*.npz`` were generated with (oracle/make_golden.py).

Only tests / smoke / the bench's CPU-baseline legs may import this module.

Tolerance note: the fit stops where the scheduler says, not at the optimum,
and the reference moves its own result by ~1e-3 when rows are merely permuted, so parity on fitted
parameters is |dm| <= 1e-2, q <= 1e-3 rel, stop epoch +-1 %.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

MAX_EPOCHS = 8192 * 2          # BidderAllocation.py:38
LR0 = 2e-3                     # BidderAllocation.py:39
BETA1, BETA2, ADAM_EPS = 0.9, 0.999, 1e-8
PLATEAU_FACTOR, PLATEAU_PATIENCE, PLATEAU_THRESHOLD, PLATEAU_EPS = 0.5, 10, 1e-4, 1e-8
STOP_AFTER, STOP_WINDOW, STOP_TOL = 1024, 100, 1e-6   # BidderAllocation.py:53


def bce_sum32(p, y):
    """torch.nn.BCELoss(reduction='sum'): log terms clamped at -100 (Models.py:25)."""
    with np.errstate(divide="ignore"):  # log(0) = -inf is the case the clamp exists for
        lp = np.maximum(np.log(p), f32(-100.0))
        l1p = np.maximum(np.log(f32(1.0) - p), f32(-100.0))
    return -(y * lp + (f32(1.0) - y) * l1p).astype(f32).sum(dtype=f32)


def fit_allocator(X, items, y, m0, q0, m_prev, max_epochs=MAX_EPOCHS, return_losses=False):
    """One ``allocator.update`` call on the won rows of one agent.

    X [n, Do+1] (context rows incl. the trailing 1), items [n] int, y [n] in {0,1};
    m0, q0, m_prev [I, Do+1] float32.  Returns dict(m, q, stop_epoch, n_epochs, final_loss, lr).
    """
    X = np.asarray(X, f32)
    items = np.asarray(items, np.int64)
    y = np.asarray(y, f32)
    m = np.array(m0, f32, copy=True)
    q = np.array(q0, f32, copy=True)
    m_prev = np.asarray(m_prev, f32)
    n = len(y)
    out = {"m": m, "q": q, "stop_epoch": -1, "n_epochs": 0, "final_loss": np.nan, "lr": LR0}
    if n < 2:  # BidderAllocation.py:33
        return out
    I, K = m.shape
    exp_avg = np.zeros_like(m)
    exp_avg_sq = np.zeros_like(m)
    lr = LR0
    best = np.inf
    bad = 0
    losses = []
    onehot = np.zeros((n, I), f32)
    onehot[np.arange(n), items] = 1
    stop_epoch = -1
    for epoch in range(max_epochs):
        # forward: predict_item (Models.py:37) + loss (Models.py:39-41)
        z = (X * m[items]).sum(axis=1, dtype=f32)
        p = (f32(1.0) / (f32(1.0) + np.exp(-z))).astype(f32)
        diff = (m_prev[:, :-1] - m[:, :-1]).astype(f32)
        prior = (q[:, :-1] * diff * diff).sum(dtype=f32)
        loss = f32(0.5) * prior + bce_sum32(p, y)
        # backward
        g_row = (p - y).astype(f32)  # d BCE / dz  (away from the log clamp)
        grad = (onehot.T @ (g_row[:, None] * X)).astype(f32)
        grad[:, :-1] += (q[:, :-1] * (m[:, :-1] - m_prev[:, :-1])).astype(f32)
        # Adam step (torch/optim/adam.py _single_tensor_adam)
        step = epoch + 1
        exp_avg += (grad - exp_avg) * f32(1 - BETA1)
        exp_avg_sq *= f32(BETA2)
        exp_avg_sq += f32(1 - BETA2) * grad * grad
        bc1 = 1 - BETA1 ** step
        bc2 = 1 - BETA2 ** step
        step_size = lr / bc1
        denom = (np.sqrt(exp_avg_sq) / f32(np.sqrt(bc2)) + f32(ADAM_EPS)).astype(f32)
        m -= (f32(step_size) * (exp_avg / denom)).astype(f32)
        losses.append(float(loss))
        # ReduceLROnPlateau.step(loss)
        cur = float(loss)
        if cur < best * (1.0 - PLATEAU_THRESHOLD):
            best = cur
            bad = 0
        else:
            bad += 1
        if bad > PLATEAU_PATIENCE:
            new_lr = lr * PLATEAU_FACTOR
            if lr - new_lr > PLATEAU_EPS:
                lr = new_lr
            bad = 0
        # early stop (BidderAllocation.py:53-55)
        if epoch > STOP_AFTER and abs(losses[-STOP_WINDOW] - losses[-1]) < STOP_TOL:
            stop_epoch = epoch
            break
    # Laplace approximation (BidderAllocation.py:58-62, Models.py:43-45) -- note the literal "1 -"
    for it in range(I):
        Xi = X[items == it]
        if len(Xi) == 0:
            continue
        P = (f32(1.0) / (f32(1.0) + np.exp(f32(1.0) - Xi @ m[it]))).astype(f32)
        q[it] += ((P * (f32(1.0) - P))[:, None] * Xi * Xi).sum(axis=0, dtype=f32)
    out.update(m=m, q=q, stop_epoch=stop_epoch, n_epochs=len(losses), final_loss=losses[-1], lr=lr)
    if return_losses:
        out["losses"] = np.asarray(losses)
    return out


def allocator_objective(X, items, y, m, q, m_prev, prior_on_intercept=False):
    """The reference's training objective (Models.py:39-41) in float64: BCE sum + 0.5 * sum over the context columns of
    q (prev_iter_m - m)^2 (``prior_on_intercept``: over all columns, AGYM_FIT_NEWTON's objective), and its gradient."""
    X, y = np.asarray(X, np.float64), np.asarray(y, np.float64)
    m, q, m_prev = (np.asarray(a, np.float64) for a in (m, q, m_prev))
    z = (X * m[items]).sum(axis=1)
    nll = np.logaddexp(0.0, -z) * y + np.logaddexp(0.0, z) * (1 - y)
    qq = q.copy()
    if not prior_on_intercept:
        qq[:, -1] = 0.0
    d = m - m_prev
    grad = np.zeros_like(m)
    np.add.at(grad, items, (1 / (1 + np.exp(-z)) - y)[:, None] * X)
    grad += qq * d
    return float(nll.sum() + 0.5 * (qq * d * d).sum()), grad


def fit_allocator_newton(X, items, y, m0, q0, m_prev, max_passes=50):
    """Restatement (float64) of the OPT-IN fit mode AGYM_FIT_NEWTON (csrc/agym_fit_newton.cu) -- NOT a reference algorithm.

    Per item with rows: damped Newton on the reference's likelihood with the Gaussian prior N(m_prev, 1/q) on ALL columns (the
    reference's loss, Models.py:39-41, leaves the intercept out, which puts the optimum of an all-click / no-click item at
    infinity), step halving when the objective does not decrease, stop when the Newton decrement g . d < 1e-9; then the
    reference's Laplace update with its literal exp(1 - z) (Models.py:43-45).  Returns dict(m, q, passes [I])."""
    X = np.asarray(X, np.float64)
    items = np.asarray(items, np.int64)
    y = np.asarray(y, np.float64)
    m = np.array(m0, np.float64, copy=True)
    q = np.array(q0, np.float64, copy=True)
    m_prev = np.asarray(m_prev, np.float64)
    I, K = m.shape
    passes = np.zeros(I, np.int64)
    if len(y) < 2:  # BidderAllocation.py:33
        return {"m": m.astype(f32), "q": q.astype(f32), "passes": passes}
    for it in range(I):
        sel = items == it
        if not sel.any():
            continue
        Xi, yi = X[sel], y[sel]
        qi = q[it].copy()
        acc, loss_acc, delta, alpha, trial = m[it].copy(), np.inf, np.zeros(K), 1.0, m[it].copy()
        while True:
            z = Xi @ trial
            e = np.exp(-np.abs(z))
            p1 = np.where(z >= 0, 1 / (1 + e), e / (1 + e))
            w = e / (1 + e) ** 2
            d = trial - m_prev[it]
            loss = (np.log1p(e) + np.where((yi > 0.5) == (z >= 0), 0.0, np.abs(z))).sum() + 0.5 * (qi * d * d).sum()
            g = Xi.T @ (p1 - yi) + qi * d
            H = (Xi * w[:, None]).T @ Xi + np.diag(qi)
            passes[it] += 1
            if not loss <= loss_acc + 1e-6 * (1.0 + abs(loss_acc)):
                alpha *= 0.5
                if alpha < 1.0 / 1024 or passes[it] >= max_passes:
                    break
                trial = acc - alpha * delta
                continue
            acc, loss_acc = trial, loss
            delta = np.linalg.solve(H * (1 + 1e-12 * np.eye(K)) + 1e-12 * np.eye(K), g)
            if g @ delta < 1e-9 or passes[it] >= max_passes:
                break
            alpha, trial = 1.0, acc - delta
        m[it] = acc
        P = 1.0 / (1.0 + np.exp(1.0 - Xi @ acc))
        q[it] += ((P * (1 - P))[:, None] * Xi * Xi).sum(axis=0)
    return {"m": m.astype(f32), "q": q.astype(f32), "passes": passes}


# ----------------------------------------------------------------------------------------------
# ValueLearningBidder / DoublyRobustBidder win-rate model  (Bidder.py:218-260, 501-538; Models.py:51-62)
# ----------------------------------------------------------------------------------------------
def fit_winrate(est, value, gamma, won, w0, lr=3e-3, weight_decay=1e-6, patience=100, factor=0.1, min_lr=1e-7,
                stop_after=512, max_epochs=8192 * 4, return_losses=False):
    """P(win | est CTR, value, gamma) = sigmoid(w[0:3] . x + w[3]) fitted on the logged rows plus the augmentation
    "had you shaded to gamma = 0 you would have lost" (Bidder.py:223-236).  Adam(lr, weight_decay, amsgrad=True) +
    ReduceLROnPlateau(patience, factor, min_lr) + "no 1e-6 improvement for `stop_after` epochs" (Bidder.py:239-260).
    ValueLearningBidder: patience 100, factor 0.1, stop 512; DoublyRobustBidder: patience 256, factor 0.2, stop 1024.
    Returns dict(w [4] float32, stop_epoch, n_epochs, final_loss)."""
    est, value, gamma = (np.asarray(v, f32) for v in (est, value, gamma))
    n = len(est)
    X = np.stack([est, value, gamma], axis=1)
    Xn = X.copy()
    Xn[:, 2] = 0.0
    X = np.concatenate([X, Xn]).astype(f32)
    y = np.concatenate([np.asarray(won, f32), np.zeros(n, f32)])
    w = np.array(w0, f32, copy=True)
    N = f32(len(y))
    ea, es, mx = np.zeros(4, f32), np.zeros(4, f32), np.zeros(4, f32)
    best_sched, bad = np.inf, 0
    best_epoch, best_loss = -1, np.inf
    losses = []
    stop_epoch = -1
    for epoch in range(max_epochs):
        z = (X @ w[:3] + w[3]).astype(f32)
        p = (f32(1.0) / (f32(1.0) + np.exp(-z))).astype(f32)
        lp = np.maximum(np.log(p), f32(-100.0))
        l1p = np.maximum(np.log(f32(1.0) - p), f32(-100.0))
        loss = f32(-(y * lp + (f32(1.0) - y) * l1p).sum(dtype=f32) / N)  # BCELoss(reduction='mean')
        g = ((p - y) / N).astype(f32)
        grad = np.concatenate([X.T @ g, [g.sum(dtype=f32)]]).astype(f32)
        grad = (grad + f32(weight_decay) * w).astype(f32)  # Adam's L2 weight decay
        step = epoch + 1
        ea += (grad - ea) * f32(1 - BETA1)
        es *= f32(BETA2)
        es += f32(1 - BETA2) * grad * grad
        mx = np.maximum(mx, es)  # amsgrad
        bc1, bc2 = 1 - BETA1 ** step, 1 - BETA2 ** step
        denom = (np.sqrt(mx) / f32(np.sqrt(bc2)) + f32(ADAM_EPS)).astype(f32)
        w -= (f32(lr / bc1) * (ea / denom)).astype(f32)
        cur = float(loss)
        losses.append(cur)
        if cur < best_sched * (1.0 - PLATEAU_THRESHOLD):
            best_sched, bad = cur, 0
        else:
            bad += 1
        if bad > patience:
            new_lr = max(lr * factor, min_lr)
            if lr - new_lr > PLATEAU_EPS:
                lr = new_lr
            bad = 0
        if (best_loss - cur) > 1e-6:
            best_epoch, best_loss = epoch, cur
        elif epoch - best_epoch > stop_after:
            stop_epoch = epoch
            break
    out = {"w": w, "stop_epoch": stop_epoch, "n_epochs": len(losses), "final_loss": losses[-1]}
    if return_losses:
        out["losses"] = np.asarray(losses)
    return out
