"""TEST INFRASTRUCTURE ONLY -- golden vectors for EmpiricalShadedBidder.update (Bidder.py:60-125) from the UNMODIFIED
reference:  python -m oracle.make_golden_empirical  ->  tests/golden/bidfit_empirical.npz"""
import os

import numpy as np

from . import auction_oracle as ao
from . import make_golden as mg


def main():
    O, GC = ao.ALLOC_ORACLE, ao.BID_GAUSS_CLIP
    kw = dict(seed=71, A=4, n_items=8, D=5, Do=4, P=2, mechanism=ao.MECH_FIRST, alloc_kinds=[O] * 4, bidder_kinds=[GC] * 4, T=4000,
              bidder_variants=[(0.8, 0.1), (0.6, 0.15), (0.95, 0.05), (0.5, 0.3)])
    case, noise, cfg = mg.build_case(**kw)
    rec, met, auction, agents = mg.run_reference(case, noise, cfg)
    out = {}
    for a, ag in enumerate(agents):
        won = np.array([o.won for o in ag.logs], bool)
        val = np.array([o.value for o in ag.logs])
        price = np.array([o.price for o in ag.logs])
        outc = np.array([o.outcome for o in ag.logs])
        util = np.zeros_like(val)
        util[won] = val[won] * outc[won] - price[won]
        out[f"a{a}_gamma"], out[f"a{a}_utility"], out[f"a{a}_won"] = np.array(ag.bidder.gammas), util, won
        out[f"a{a}_value"], out[f"a{a}_price"], out[f"a{a}_outcome"] = val, price, outc
        ag.update(iteration=0)
        out[f"a{a}_best_gamma"] = np.float64(ag.bidder.prev_gamma)
        print(f"empirical agent {a}: rows {len(won)}, best gamma {ag.bidder.prev_gamma:.6f}")
    path = os.path.join(mg.GOLDEN_DIR, "bidfit_empirical.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
