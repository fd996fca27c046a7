"""Bare-name module, same file name as the reference's src/Bidder.py so that ``from Bidder import ...`` keeps working
when this directory is on sys.path.  The implementation lives in the auction_gym_b200 package."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from auction_gym_b200.bidders import (Bidder, DoublyRobustBidder, EmpiricalShadedBidder, PolicyLearningBidder,  # noqa: E402,F401
                                      TruthfulBidder, ValueLearningBidder)
