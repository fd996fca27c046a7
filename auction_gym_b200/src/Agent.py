"""Bare-name module, same file name as the reference's src/Agent.py so that ``from Agent import ...`` keeps working
when this directory is on sys.path.  The implementation lives in the auction_gym_b200 package."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from auction_gym_b200.agent import Agent  # noqa: E402,F401
