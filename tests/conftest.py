import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """Load tests/golden/<name>.npz -> (case, inputs, ref_records, ref_metrics)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    case, inp, rec, met = {}, {}, {}, {}
    for k in z.files:
        if k.startswith("case_"):
            v = z[k]
            case[k[5:]] = v.item() if v.ndim == 0 else v
        elif k.startswith("in_"):
            inp[k[3:]] = z[k]
        elif k.startswith("ref_"):
            rec[k[4:]] = z[k]
        elif k.startswith("met_"):
            met[k[4:]] = z[k]
    return case, inp, rec, met


def round_golden_names():
    """Single-slot round goldens (the several-slots-per-round ones have their own tests: multislot_golden_names)."""
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "rounds_*.npz"))) if "_slots" not in n)


def multislot_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "rounds_*_slots*.npz")))


@pytest.fixture(scope="session")
def golden_loader():
    return load_golden
