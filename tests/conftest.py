import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden():
    from tests import helpers
    return helpers.load_golden()


@pytest.fixture(scope="session")
def fixture_dir(tmp_path_factory, golden):
    """Materialise tests/golden/fixtures.json as files (gz re-compressed) and return the directory."""
    from tests import helpers
    return helpers.materialise_fixtures(tmp_path_factory.mktemp("fixtures"), golden[0])
