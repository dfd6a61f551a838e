import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ASSETS = "/root/reference/assistive_gym/envs/assets"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "assets: needs the reference asset tree (build container only)")


def pytest_collection_modifyitems(config, items):
    have_assets = os.path.isdir(ASSETS)
    for item in items:
        if "assets" in item.keywords and not have_assets:
            item.add_marker(pytest.mark.skip(reason="reference assets not present on this machine"))


@pytest.fixture(scope="session")
def env_data():
    from assistive_vr_gym_b200.envs import load_env_data
    return load_env_data("ScratchItchJaco.npz")


@pytest.fixture(scope="session")
def oracles(env_data):
    from oracle.oracle import Oracle
    return [Oracle(b) for b in env_data[0]]
