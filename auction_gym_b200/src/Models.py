"""Bare-name module, same file name as the reference's src/Models.py so that ``from Models import ...`` keeps working
when this directory is on sys.path.  The implementation lives in the auction_gym_b200 package."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
import numpy as _np  # noqa: E402


def sigmoid(x):
    """Models.py:10-12."""
    return 1.0 / (1.0 + _np.exp(-x))
