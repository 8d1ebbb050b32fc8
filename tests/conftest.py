import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests never silently pass on a box without a device: they are deselected by `-m "not gpu"`,
    # and fail loudly if selected where no device exists.
    pass


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ai_model_dir(tmp_path_factory):
    """A directory laid out like the encoder's cwd: available AI checkpoints + Thr_info.txt."""
    from oracle import assets

    d = str(tmp_path_factory.mktemp("ai_models"))
    present = assets.materialize(d, "AI")
    return d, present


@pytest.fixture(scope="session")
def ldp_model_dir(tmp_path_factory):
    from oracle import assets

    d = str(tmp_path_factory.mktemp("ldp_models"))
    present = assets.materialize(d, "LDP")
    return d, present


@pytest.fixture(scope="session")
def eb():
    import ethcnn_b200

    return ethcnn_b200
