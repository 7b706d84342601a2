import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Build the native pieces once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return True


def oracle_fk(plant):
    """Forward-kinematics callable for synth.generate backed by the oracle."""
    import numpy as np

    def fk(q, v):
        n = len(q)
        P, V = np.zeros((n, 4, 3)), np.zeros((n, 4, 3))
        for i in range(n):
            for k, f in enumerate(plant.foot_frames):
                p, J, _ = plant.frame_position_quantities(q[i], v[i], f)
                P[i, k], V[i, k] = p, J @ v[i]
        return P, V
    return fk
