"""Per-(agent, round) log record with the reference's field names (src/Impression.py:4-31).

Inside the engine the log is SoA on the device (``agym_round_log``); records of this class are only
materialised when a caller iterates ``agent.logs`` (notebooks, src/main.py:147)."""
from dataclasses import dataclass

import numpy as np


@dataclass
class ImpressionOpportunity:
    __slots__ = ["context", "item", "value", "bid", "best_expected_value", "true_CTR", "estimated_CTR", "price",
                 "second_price", "outcome", "won"]

    context: np.ndarray
    item: int
    value: float
    bid: float
    best_expected_value: float
    true_CTR: float
    estimated_CTR: float
    price: float
    second_price: float
    outcome: bool
    won: bool
