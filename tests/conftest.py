import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def egnn():
    import egnn_b200
    return egnn_b200


@pytest.fixture(scope="session")
def small_graph():
    """Elliptic-shaped graph small enough for the CPU oracle to finish in well under a second."""
    from egnn_b200 import synthetic
    return synthetic.make_elliptic_like(n_nodes=6000, n_edges=7000, n_feats=166, n_timesteps=12, seed=3,
                                        hub_degree=200, t_train_end=8, t_val_end=10, train_window_k=6)
