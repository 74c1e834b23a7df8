"""pytest configuration: registers the `gpu` marker and shared fixtures."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: longer CPU test")


@pytest.fixture(scope="session")
def scene_dir(tmp_path_factory):
    """Directory with the scene files written out in the reference's text format."""
    import scenes
    d = tmp_path_factory.mktemp("scenes")
    scenes.write_all(str(d))
    return str(d)
