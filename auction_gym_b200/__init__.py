"""auction_gym_b200 -- B200-native batched engine for AuctionGym's round loop.

``import auction_gym_b200`` for the package, or put ``auction_gym_b200/src`` on ``sys.path`` to get the
reference's bare module names (``from Auction import Auction`` ...), exactly as the reference is used.
(``auction-gym_b200`` at the repository root is a symbolic link to this directory: the project's name
carries a hyphen, which a Python module name cannot.)
"""
from . import _lib  # noqa: F401
from ._lib import AgymError  # noqa: F401
from .engine import Engine  # noqa: F401
from .mechanisms import AllocationMechanism, FirstPrice, SecondPrice  # noqa: F401
from .allocators import Allocator, LogisticTSAllocator, OracleAllocator, PyTorchLogisticRegressionAllocator  # noqa: F401
from .bidders import (Bidder, DoublyRobustBidder, EmpiricalShadedBidder, PolicyLearningBidder, TruthfulBidder,  # noqa: F401
                      ValueLearningBidder)
from .impression import ImpressionOpportunity  # noqa: F401
from .agent import Agent  # noqa: F401
from .auction import Auction  # noqa: F401
from . import driver  # noqa: F401
from .driver import instantiate_agents, instantiate_auction, parse_config, run_experiment, shard_runs, write_csvs  # noqa: F401
