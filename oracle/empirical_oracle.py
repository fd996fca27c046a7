"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64) of EmpiricalShadedBidder.update (reference
src/Bidder.py:60-125): bucketised estimate of the utility per shading factor and the move of prev_gamma to the bucket with
the best lower confidence bound.  Pinned against the reference's own result in tests/golden/bidfit_empirical.npz."""
import numpy as np


def fit_empirical(gammas, utilities, grid_delta=0.005, critical_value=1.96):
    gammas = np.asarray(gammas, np.float64)
    utilities = np.asarray(utilities, np.float64)
    lo, hi = gammas.min(), gammas.max()
    num_buckets = int((hi - lo) // grid_delta) + 1          # Bidder.py:82
    edges = np.linspace(lo, hi, num_buckets)                # Bidder.py:83
    centres, lower = [], []
    for b_lo, b_hi in zip(edges[:-1], edges[1:]):
        centres.append((b_hi - b_lo) / 2.0 + b_lo)          # Bidder.py:90
        u = utilities[(gammas < b_hi) & (b_lo <= gammas)]   # Bidder.py:92
        lower.append(u.mean() - critical_value * u.std() / np.sqrt(len(u)) if len(u) > 1 else np.nan)
    lower = np.asarray(lower)
    best = len(centres) - np.nanargmax(lower[::-1]) - 1     # Bidder.py:119: the highest bucket among ties
    return float(np.clip(centres[best], 0.0, 1.0)), int(best), len(centres)
