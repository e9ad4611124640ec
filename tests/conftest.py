import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
# the multi-band tests run several bands of ONE process on one GPU, each with two streams that spin on each other's flags:
# with the default 8 hardware queues two such streams can share a queue and serialise behind a spinning kernel
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")


def golden_names():
    """Trajectory fixtures of oracle/gen_golden.py (big_*: full-size summaries, es_*: ES fitness values, ref_*: other recordings have their own tests)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith(("big_", "es_", "ref_")))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
