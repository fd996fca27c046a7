"""auction-gym_b200 -- B200-native batched engine for AuctionGym's round loop.

The directory name carries a hyphen (the project's name); import it as ``auction_gym_b200`` (the
alias package at the repository root) or put this directory's ``src`` on ``sys.path`` to get the
reference's bare module names (``from Auction import Auction`` ...), as the reference does.
"""
from . import _lib  # noqa: F401
from ._lib import AgymError  # noqa: F401
from .engine import Engine  # noqa: F401

__all__ = ["Engine", "AgymError", "_lib"]
