import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# Property tests run the same examples on every machine (what passed here is what the driver runs); set
# CLIFT_HYPOTHESIS_RANDOM=1 to explore fresh examples.
try:
    from hypothesis import settings as _hyp_settings

    _hyp_settings.register_profile("pinned", derandomize=True, database=None, deadline=None)
    _hyp_settings.register_profile("explore", deadline=None)
    _hyp_settings.load_profile("explore" if os.environ.get("CLIFT_HYPOTHESIS_RANDOM") == "1" else "pinned")
except ImportError:      # hypothesis absent: the property tests fail at import, everything else runs
    pass


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
