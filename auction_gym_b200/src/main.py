"""Bare-name module, same file name as the reference's src/main.py so that ``from main import ...`` keeps working
when this directory is on sys.path.  The implementation lives in the auction_gym_b200 package."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from auction_gym_b200.driver import (instantiate_agents, instantiate_auction, main, parse_config, parse_kwargs,  # noqa: E402,F401
                                     run_experiment, simulation_run, write_csvs)

if __name__ == "__main__":
    main()
