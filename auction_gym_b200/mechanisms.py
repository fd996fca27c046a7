"""Auction allocation mechanisms -- same names and call signature as the reference's
src/AuctionAllocation.py.  Inside the engine the rule is a flag of the resolution kernel
(``agym_shape.mechanism``); ``allocate`` is kept for callers that resolve a single bid vector on the
host (notebooks), and follows AuctionAllocation.py:18-23,32-35 (lowest slot wins ties)."""
import numpy as np

from . import _lib


class AllocationMechanism:
    """Base class for allocation mechanisms (AuctionAllocation.py:3-9)."""

    code = None

    def allocate(self, bids, num_slots):
        raise NotImplementedError

    @staticmethod
    def _order(bids, num_slots):
        bids = np.asarray(bids, dtype=np.float64)
        order = np.argsort(-bids, kind="stable")  # stable: ties keep the lowest slot first
        return bids, order[:num_slots], bids[order]


class FirstPrice(AllocationMechanism):
    """(Generalised) first-price: winners pay their own bid (AuctionAllocation.py:12-23)."""

    code = _lib.FIRST_PRICE

    def allocate(self, bids, num_slots):
        _, winners, ranked = self._order(bids, num_slots)
        return winners, ranked[:num_slots], ranked[1:num_slots + 1]


class SecondPrice(AllocationMechanism):
    """(Generalised) second-price: winners pay the next bid (AuctionAllocation.py:26-35)."""

    code = _lib.SECOND_PRICE

    def allocate(self, bids, num_slots):
        _, winners, ranked = self._order(bids, num_slots)
        prices = ranked[1:num_slots + 1]
        return winners, prices, prices
