"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of AuctionGym's round loop in replay mode.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module; the product path (``auction_gym_b200/``) never does.

Parity pin: the reference ships no tests or golden vectors, so this
restatement is pinned against OUTPUTS OF THE UNMODIFIED REFERENCE run in the build container
through ``oracle/ref_harness.py``; the vectors live in ``tests/golden/*.npz`` and were written by
``oracle/make_golden.py`` (committed).  ``tests/test_oracle_golden.py`` checks every function here
against them.

Two entry points restate the same arithmetic:
  * ``simulate_rounds``        -- vectorised over rounds (fast; used by parity tests)
  * ``simulate_rounds_scalar`` -- one opportunity at a time in the reference's own operation order
                                  (used as the CPU baseline "port": same per-round Python/numpy
                                  call pattern as src/Auction.py:28-74)

Precision follows the reference: the Oracle path is float64 end to end;
the learnt path computes the CTR estimate in float32 on a float32 copy of the observed context
and everything downstream (est x value, argmax, bid, price, utilities) in float64.
"""
from __future__ import annotations

import numpy as np

# allocator kinds (BidderAllocation.py:21,71)
ALLOC_ORACLE, ALLOC_TS, ALLOC_MAP = 0, 1, 2
# bidder kinds at bid time (Bidder.py:28,38,156,336,442)
BID_TRUTHFUL, BID_GAUSS, BID_GAUSS_CLIP, BID_SEARCH, BID_BANDIT, BID_POLICY = 0, 1, 2, 3, 4, 5
MECH_SECOND, MECH_FIRST = 0, 1

# columns of the per-agent accumulator block (Agent.py:70-118, main.py:131-148)
M_NET, M_GROSS, M_ALLOC_REG, M_ESTIM_REG, M_OVERBID, M_UNDERBID, M_SQERR, M_BIAS, M_NPART, M_NWON, M_BEST_EV, M_GAMMA = range(12)
NUM_METRICS = 12


def sigmoid64(x):
    """Models.py:10-12 (numba, float64)."""
    return 1.0 / (1.0 + np.exp(-x))


def sigmoid32(x):
    """torch.sigmoid on float32 (Models.py:31,33)."""
    x = np.asarray(x, dtype=np.float32)
    one = np.float32(1.0)
    return (one / (one + np.exp(-x))).astype(np.float32)


def softplus32(x):
    """torch.nn.Softplus (beta=1, threshold=20) on float32 (Models.py:84-85,149-150)."""
    x = np.asarray(x, dtype=np.float32)
    return np.where(x > np.float32(20.0), x, np.log1p(np.exp(np.minimum(x, np.float32(20.0))))).astype(np.float32)


def ts_weights(m, q, eps):
    """Models.py:31 -- m + eps * (1/sqrt(q)), each op rounded to float32 separately."""
    m = np.asarray(m, np.float32)
    std = (np.float32(1.0) / np.sqrt(np.asarray(q, np.float32))).astype(np.float32)
    return (m + (np.asarray(eps, np.float32) * std).astype(np.float32)).astype(np.float32)


def linear32(x, w):
    """F.linear(x, w) for x [n, K], w [n, I, K] -> [n, I]; float32 sequential fma-free accumulation."""
    acc = np.zeros(w.shape[:-1], np.float32)
    for k in range(w.shape[-1]):
        acc = (acc + (w[..., k] * x[:, None, k]).astype(np.float32)).astype(np.float32)
    return acc


def make_contexts(ctx):
    """Auction.py:33,36 -- true context = [ctx, 1];  observed = [ctx[:Do], 1] (built by the caller)."""
    T = ctx.shape[0]
    return np.concatenate([ctx, np.ones((T, 1))], axis=1)


def estimate_and_select(case, a, true_ctx, obs_ctx, eps):
    """Agent.select_item (Agent.py:29-42) for a batch of rounds of agent ``a``.

    true_ctx [n, D+1] f64, obs_ctx [n, Do+1] f64, eps [n, I, Do+1] f32 or None.
    Returns item [n] int, est [n] f64 (value of the f64 or f32 estimate the reference logs) and the
    relative decision margin (best - runner-up) / |best| of the arg-max (1.0 when there is one item).
    """
    nI = int(case["n_items"][a])
    V = case["V"][a, :nI]
    kind = int(case["alloc_kind"][a])
    if kind == ALLOC_ORACLE:
        # BidderAllocation.py:81-82
        est_all = sigmoid64(true_ctx @ case["E"][a, :nI].T)
        item = np.argmax(est_all * V[None, :], axis=1)
        est = est_all[np.arange(len(item)), item]
        return item, est, _margin(est_all * V[None, :])
    x32 = obs_ctx.astype(np.float32)  # BidderAllocation.py:68
    m = case["m"][a, :nI].astype(np.float32)
    n = x32.shape[0]
    mb = np.broadcast_to(m, (n,) + m.shape)
    est_map = sigmoid32(linear32(x32, mb))
    if kind == ALLOC_TS:
        w = ts_weights(mb, np.broadcast_to(case["q"][a, :nI], mb.shape), eps[:, :nI])
        est_ts = sigmoid32(linear32(x32, w))
        score = est_ts.astype(np.float64) * V[None, :]  # Agent.py:33-35
    else:
        score = est_map.astype(np.float64) * V[None, :]
    item = np.argmax(score, axis=1)
    est = est_map[np.arange(n), item].astype(np.float64)  # Agent.py:38-40 (MAP of the TS-chosen item)
    return item, est, _margin(score)


def _margin(score):
    if score.shape[1] < 2:
        return np.ones(score.shape[0])
    srt = np.sort(score, axis=1)
    return (srt[:, -1] - srt[:, -2]) / np.abs(srt[:, -1])


def gaussian_pdf(prev_gamma, sigma, g):
    """Bidder.py:178,355,462."""
    return np.exp(-((prev_gamma - g) / sigma) ** 2 / 2) / (sigma * np.sqrt(2 * np.pi))


def winrate32(w, x):
    """PyTorchWinRateEstimator.forward (Models.py:61-62): sigmoid(w[0:3].x + w[3]) in float32."""
    x = np.asarray(x, np.float32)
    z = np.zeros(x.shape[:-1], np.float32)
    for k in range(3):
        z = (z + (x[..., k] * np.float32(w[k])).astype(np.float32)).astype(np.float32)
    z = (z + np.float32(w[3])).astype(np.float32)
    return sigmoid32(z)


def bandit_mu_sigma32(w, x):
    """BidShadingContextualBandit.forward / BidShadingPolicy.forward (Models.py:82-85,146-151).

    w is the flat float32 parameter vector
      [W1(2x2 row-major), b1(2), w_mu(2), b_mu, w_sigma(2), b_sigma]   (12 floats)
    x [..., 2] = [estimated_CTR, value].  Returns (mu, sigma) float32, sigma includes min_sigma 1e-2.
    """
    x = np.asarray(x, np.float32)
    w = np.asarray(w, np.float32)
    h0 = ((x[..., 0] * w[0]).astype(np.float32) + (x[..., 1] * w[1]).astype(np.float32)).astype(np.float32) + w[4]
    h1 = ((x[..., 0] * w[2]).astype(np.float32) + (x[..., 1] * w[3]).astype(np.float32)).astype(np.float32) + w[5]
    s0, s1 = softplus32(h0), softplus32(h1)
    mu = softplus32(((s0 * w[6]).astype(np.float32) + (s1 * w[7]).astype(np.float32)).astype(np.float32) + w[8])
    sg = softplus32(((s0 * w[9]).astype(np.float32) + (s1 * w[10]).astype(np.float32)).astype(np.float32) + w[11])
    return mu.astype(np.float32), (sg + np.float32(1e-2)).astype(np.float32)


def compute_bids(case, a, value, est, gamma_z, grid_u):
    """Bidder.bid for a batch of rounds of agent ``a``.  Returns bid, gamma, propensity (f64; NaN if n/a)."""
    kind = int(case["bidder_kind"][a])
    n = len(value)
    nan = np.full(n, np.nan)
    if kind == BID_TRUTHFUL:
        return value * est, nan, nan  # Bidder.py:34-35
    bp = case["bidder_f"][a]
    prev_gamma, sigma = float(bp[0]), float(bp[1])
    bid = value * est
    if kind in (BID_GAUSS, BID_GAUSS_CLIP):
        gamma = prev_gamma + sigma * gamma_z  # Bidder.py:177,354,461 (unclipped) / :51
        if kind == BID_GAUSS_CLIP:
            gamma = np.clip(gamma, 0.0, 1.0)  # Bidder.py:52-55
            prop = nan
        else:
            prop = gaussian_pdf(prev_gamma, sigma, gamma)
        return bid * gamma, gamma, prop
    if kind == BID_SEARCH:
        # Bidder.py:180-196
        G = grid_u.shape[1]
        grid = np.sort(0.1 + (1.0 - 0.1) * grid_u, axis=1)
        x = np.stack([np.broadcast_to(est[:, None], (n, G)), np.broadcast_to(value[:, None], (n, G)), grid], axis=-1)
        pw = winrate32(case["winrate_w"][a], x.astype(np.float32))
        util = pw.astype(np.float64) * (bid[:, None] - bid[:, None] * grid)
        gamma = grid[np.arange(n), np.argmax(util, axis=1)]
        return bid * gamma, gamma, np.ones(n)
    if kind in (BID_BANDIT, BID_POLICY):
        # Bidder.py:198-203,357-362,464-470 ; Models.py:82-90,146-155
        x = np.stack([est, value], axis=-1).astype(np.float32)
        mu, sg = bandit_mu_sigma32(case["policy_w"][a], x)
        z = gamma_z.astype(np.float32)
        sample = (mu + (z * sg).astype(np.float32)).astype(np.float32)
        logp = (-((sample - mu) ** 2) / (np.float32(2.0) * sg * sg) - np.log(sg) - np.float32(0.5 * np.log(2 * np.pi))).astype(np.float32)
        prop = np.exp(logp).astype(np.float64)
        gamma = np.clip(sample, np.float32(0.0), np.float32(1.0)).astype(np.float64)
        return bid * gamma, gamma, prop
    raise ValueError(f"unknown bidder kind {kind}")


def resolve(bids, mechanism):
    """AuctionAllocation.py:18-23,32-35 with num_slots == 1.

    bids [T, P] f64 -> winner slot [T] (lowest slot among maxima), price [T], second [T], valid [T].
    P == 1: the reference's price array is empty, nobody is charged (Auction.py:68 zip) -> valid False.
    """
    T, P = bids.shape
    winner = np.argmax(bids, axis=1)
    if P < 2:
        return winner, np.zeros(T), np.zeros(T), np.zeros(T, bool)
    srt = -np.sort(-bids, axis=1)
    second = srt[:, 1]
    price = srt[:, 0] if mechanism == MECH_FIRST else second
    return winner, price, second, np.ones(T, bool)


def resolve_slots(bids, mechanism, num_slots):
    """AuctionAllocation.py:18-23,32-35 with num_slots [T] >= 1 (Auction.py:30,60-74; `max_slots` > 1).

    Returns rank [T, P] (0 = highest bid; equal bids keep slot order: numpy's argsort of a handful of elements is an insertion
    sort), srt [T, P] bids in descending order, K [T] = slots actually charged.  The reference zips winners, prices and
    second_prices (Auction.py:68); second_prices = sorted[1 : num_slots + 1] has only P - 1 entries at most, and for
    SecondPrice so has prices, so K = min(num_slots, P - 1): the slot whose runner-up does not exist is silently dropped.
    """
    T, P = bids.shape
    order = np.argsort(-bids, axis=1, kind="stable")
    rank = np.empty_like(order)
    np.put_along_axis(rank, order, np.arange(P)[None, :], axis=1)
    srt = np.take_along_axis(bids, order, axis=1)
    K = np.minimum(np.asarray(num_slots, np.int64), P - 1)
    return rank, srt, K


def _finish_multislot(rec, parts, u, mechanism, num_slots, A):
    """Charging with several slots per round (Auction.py:60-74, Agent.py:70-77).  Slot k's winner is charged price_k and
    logs (price_k, second_k, outcome_k, won); every other participant's logged price is overwritten by set_price(price_k) --
    so after the last slot EVERY participant's logged price is the last slot's price, while utilities were charged slot by
    slot.  The log-derived regrets (Agent.py:104-112) therefore see the last price; net utility and revenue do not."""
    T, P = parts.shape
    rank, srt, K = resolve_slots(rec["bid"], mechanism, num_slots)
    srt_pad = np.concatenate([srt, np.zeros((T, 1))], axis=1)
    ar = np.arange(T)[:, None]
    won = rank < K[:, None]
    own_price = np.take_along_axis(srt_pad, rank + (0 if mechanism == MECH_FIRST else 1), axis=1)  # price of the slot the participant wins
    own_second = np.take_along_axis(srt_pad, rank + 1, axis=1)
    last = np.clip(K - 1, 0, None)
    last_price = np.where(K > 0, srt_pad[np.arange(T), last + (0 if mechanism == MECH_FIRST else 1)], 0.0)
    u2 = np.asarray(u, np.float64).reshape(T, -1)
    uk = np.take_along_axis(u2, np.minimum(rank, u2.shape[1] - 1), axis=1)  # the click uniform of the slot the participant wins
    outcome = won & (uk < rec["true_ctr"])
    rec["rank"] = rank.astype(np.int32)
    rec["won"] = won.astype(np.uint8)
    rec["outcome"] = outcome.astype(np.uint8)
    rec["price"] = np.broadcast_to(last_price[:, None], (T, P)).copy()
    rec["second"] = np.where(won, own_second, 0.0)
    rec["paid"] = np.where(won, own_price, 0.0)
    rec["winner"] = np.argmin(rank, axis=1).astype(np.int32)
    rec["n_charged"] = K.astype(np.int32)
    srt2 = np.sort(rec["bid"], axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        gaps = np.diff(srt2, axis=1) / np.maximum(np.abs(srt2[:, 1:]), 1e-300)
    rec["bid_margin"] = np.nan_to_num(gaps.min(axis=1), nan=0.0) if P >= 2 else np.ones(T)
    metrics = accumulate_metrics(rec, parts, A)
    tv = rec["true_ctr"] * rec["value"]
    flat = parts.ravel()
    metrics["acc"][:, M_NET] = np.bincount(flat, weights=np.where(won, rec["value"] * outcome - rec["paid"], 0.0).ravel(), minlength=A)[:A]
    metrics["revenue"] = np.float64(rec["paid"].sum())  # Auction.py:74: += price per charged slot
    return rec, metrics


def simulate_rounds(case, ctx, parts, u, ts_eps=None, gamma_z=None, grid_u=None, num_slots=None):
    """Vectorised restatement of T calls of Auction.simulate_opportunity (Auction.py:28-74).
    ``num_slots`` [T] (with ``u`` [T, max_slots]) switches to several slots per round (Auction.py:30, `max_slots` > 1).

    ctx [T, D] f64 (already scaled by embedding_var), parts [T, P] int, u [T] f64 click uniforms,
    ts_eps [T, P, I, Do+1] f32, gamma_z [T, P] f64, grid_u [T, P, G] f64.
    Returns (rec, metrics): rec has per-(round, slot) arrays, metrics the [A, NUM_METRICS] accumulator
    block plus revenue.
    """
    T, P = parts.shape
    A = len(case["alloc_kind"])
    Do = int(case["Do"])
    true_ctx = make_contexts(ctx)
    obs_ctx = np.concatenate([ctx[:, :Do], np.ones((T, 1))], axis=1)
    rec = {k: np.zeros((T, P), np.float64) for k in ("est", "value", "bid", "true_ctr", "best_ev", "price", "second")}
    rec["gamma"] = np.full((T, P), np.nan)
    rec["propensity"] = np.full((T, P), np.nan)
    rec["item_margin"] = np.ones((T, P))
    rec["item"] = np.zeros((T, P), np.int32)
    rec["outcome"] = np.zeros((T, P), np.uint8)
    rec["won"] = np.zeros((T, P), np.uint8)
    for s in range(P):
        for a in range(A):
            rows = np.nonzero(parts[:, s] == a)[0]
            if len(rows) == 0:
                continue
            nI = int(case["n_items"][a])
            eps = None if ts_eps is None else ts_eps[rows, s]
            item, est, margin = estimate_and_select(case, a, true_ctx[rows], obs_ctx[rows], eps)
            rec["item_margin"][rows, s] = margin
            value = case["V"][a, item]  # Agent.py:49
            gz = None if gamma_z is None else gamma_z[rows, s]
            gu = None if grid_u is None else grid_u[rows, s]
            bid, gamma, prop = compute_bids(case, a, value, est, gz, gu)
            true_all = sigmoid64(true_ctx[rows] @ case["E"][a, :nI].T)  # Auction.py:52
            rec["item"][rows, s] = item
            rec["est"][rows, s] = est
            rec["value"][rows, s] = value
            rec["bid"][rows, s] = bid
            rec["gamma"][rows, s] = gamma
            rec["propensity"][rows, s] = prop
            rec["best_ev"][rows, s] = np.max(true_all * case["V"][a, :nI][None, :], axis=1)  # Auction.py:53
            rec["true_ctr"][rows, s] = true_all[np.arange(len(rows)), item]
    if num_slots is not None:
        return _finish_multislot(rec, parts, u, int(case["mechanism"]), num_slots, A)
    winner, price, second, valid = resolve(rec["bid"], int(case["mechanism"]))
    ar = np.arange(T)
    outcome = (u < rec["true_ctr"][ar, winner]) & valid  # Auction.py:65 (replay click rule)
    rec["winner"] = winner.astype(np.int32)
    if P >= 2:
        srt = np.sort(rec["bid"], axis=1)
        with np.errstate(invalid="ignore", divide="ignore"):
            rec["bid_margin"] = np.nan_to_num((srt[:, -1] - srt[:, -2]) / np.abs(srt[:, -1]), nan=0.0)
    else:
        rec["bid_margin"] = np.ones(T)
    rec["won"][ar[valid], winner[valid]] = 1
    rec["outcome"][ar, winner] = outcome.astype(np.uint8)
    rec["price"][:] = np.where(valid, price, 0.0)[:, None]  # Agent.py:70-77 (losers log the price too)
    rec["second"][ar, winner] = np.where(valid, second, 0.0)
    metrics = accumulate_metrics(rec, parts, A)
    metrics["revenue"] = np.float64(np.sum(price[valid]))  # Auction.py:74
    return rec, metrics


def accumulate_metrics(rec, parts, A):
    """Agent.charge / metric getters (Agent.py:70-118) and main.py:131-148 as per-agent sums."""
    acc = np.zeros((A, NUM_METRICS), np.float64)
    won = rec["won"].astype(bool)
    tv = rec["true_ctr"] * rec["value"]
    cols = {
        M_NET: np.where(won, rec["value"] * rec["outcome"] - rec["price"], 0.0),
        M_GROSS: np.where(won, rec["value"] * rec["outcome"], 0.0),
        M_ALLOC_REG: rec["best_ev"] - tv,
        M_ESTIM_REG: rec["est"] * rec["value"] - tv,
        M_OVERBID: np.where(won, rec["price"] - rec["second"], 0.0),
        M_UNDERBID: np.where(~won & (rec["price"] < tv), rec["price"] - rec["bid"], 0.0),
        M_SQERR: (rec["true_ctr"] - rec["est"]) ** 2,
        M_BIAS: np.where(won, rec["est"] / rec["true_ctr"], 0.0),
        M_NPART: np.ones_like(tv),
        M_NWON: won.astype(np.float64),
        M_BEST_EV: rec["best_ev"],
        M_GAMMA: np.nan_to_num(rec["gamma"], nan=0.0),
    }
    flat = parts.ravel()
    for c, v in cols.items():
        acc[:, c] = np.bincount(flat, weights=v.ravel(), minlength=A)[:A]
    return {"acc": acc}


def derived_metrics(acc):
    """The ten per-agent numbers main.py:131-148 appends each iteration, from the accumulator block."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return {
            "net": acc[..., M_NET], "gross": acc[..., M_GROSS],
            "alloc_regret": acc[..., M_ALLOC_REG], "estim_regret": acc[..., M_ESTIM_REG],
            "overbid_regret": acc[..., M_OVERBID], "underbid_regret": acc[..., M_UNDERBID],
            "ctr_rmse": np.sqrt(acc[..., M_SQERR] / acc[..., M_NPART]),
            "ctr_bias": acc[..., M_BIAS] / acc[..., M_NWON],
            "best_ev_mean": acc[..., M_BEST_EV] / acc[..., M_NPART],
            "gamma_mean": acc[..., M_GAMMA] / acc[..., M_NPART],
        }


def simulate_rounds_scalar(case, ctx, parts, u, ts_eps=None, gamma_z=None, grid_u=None):
    """One opportunity at a time, in the reference's operation order (Auction.py:28-74).

    This is the "port" timed as the CPU baseline: one Python iteration per round with the same
    handful of small numpy calls per participant that the reference makes.  Results equal
    ``simulate_rounds``.
    """
    T, P = parts.shape
    A = len(case["alloc_kind"])
    Do = int(case["Do"])
    mech = int(case["mechanism"])
    acc = np.zeros((A, NUM_METRICS), np.float64)
    revenue = 0.0
    winners = np.zeros(T, np.int32)
    prices = np.zeros(T, np.float64)
    outcomes = np.zeros(T, np.uint8)
    items = np.zeros((T, P), np.int32)
    one = np.ones(1)
    for t in range(T):
        true_ctx = np.concatenate((ctx[t], one))
        obs_ctx = np.concatenate((true_ctx[:Do], one))
        bids = np.empty(P)
        ctrs = np.empty(P)
        rows = []
        for s in range(P):
            a = int(parts[t, s])
            nI = int(case["n_items"][a])
            V = case["V"][a, :nI]
            eps = None if ts_eps is None else ts_eps[t, s][None]
            item, est, _ = estimate_and_select(case, a, true_ctx[None], obs_ctx[None], eps)
            item, est = int(item[0]), est[0]
            value = V[item]
            gz = None if gamma_z is None else gamma_z[t, s:s + 1]
            gu = None if grid_u is None else grid_u[t, s][None]
            bid, gamma, _ = compute_bids(case, a, np.array([value]), np.array([est]), gz, gu)
            true_all = sigmoid64(true_ctx @ case["E"][a, :nI].T)
            bids[s] = bid[0]
            ctrs[s] = true_all[item]
            items[t, s] = item
            rows.append((a, est, value, bid[0], np.max(true_all * V), true_all[item], gamma[0]))
        w, price, second, valid = resolve(bids[None], mech)
        w, price, second, valid = int(w[0]), price[0], second[0], bool(valid[0])
        outcome = bool(u[t] < ctrs[w]) and valid
        winners[t], prices[t], outcomes[t] = w, price if valid else 0.0, outcome
        for s, (a, est, value, bid, best_ev, true_sel, gamma) in enumerate(rows):
            tv = true_sel * value
            won = valid and s == w
            r = acc[a]
            if won:
                r[M_NET] += value * outcome - price
                r[M_GROSS] += value * outcome
                r[M_OVERBID] += price - second
                r[M_BIAS] += est / true_sel
                r[M_NWON] += 1
            else:
                p_logged = price if valid else 0.0
                if p_logged < tv:
                    r[M_UNDERBID] += p_logged - bid
            r[M_ALLOC_REG] += best_ev - tv
            r[M_ESTIM_REG] += est * value - tv
            r[M_SQERR] += (true_sel - est) ** 2
            r[M_NPART] += 1
            r[M_BEST_EV] += best_ev
            if not np.isnan(gamma):
                r[M_GAMMA] += gamma
        if valid:
            revenue += price
    return {"winner": winners, "price": prices, "outcome": outcomes, "item": items}, {"acc": acc, "revenue": np.float64(revenue)}


# ----------------------------------------------------------------------------------------------
# Production-mode statistical helpers (distribution-level parity)
# ----------------------------------------------------------------------------------------------

def draw_replay_inputs(rng, T, A, P, D, I=None, Do=None, embedding_var=1.0, want_eps=False, want_gamma=False, grid=0):
    """Host-draw the noise one replay case needs, with numpy's own generator."""
    ctx = rng.normal(0.0, embedding_var, size=(T, D))
    parts = np.stack([rng.choice(A, P, replace=False) for _ in range(T)]).astype(np.int32)
    u = rng.random(T)
    out = {"ctx": ctx, "parts": parts, "u": u}
    if want_eps:
        out["ts_eps"] = rng.standard_normal((T, P, I, Do + 1)).astype(np.float32)
    if want_gamma:
        out["gamma_z"] = rng.standard_normal((T, P))
    if grid:
        out["grid_u"] = rng.random((T, P, grid))
    return out


def make_catalog(rng, A, I, D, embedding_var=1.0):
    """main.py:60-72 -- embeddings N(0, var), values LogNormal(0.1, 0.2), intercept -3 - U[0,1)."""
    E = rng.normal(0.0, embedding_var, size=(A, I, D))
    V = rng.lognormal(0.1, 0.2, size=(A, I))
    b = -3.0 - rng.random((A, I, 1))
    return np.concatenate([E, b], axis=2), V
