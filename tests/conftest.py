import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def weights():
    from oracle import xfeat_oracle as xo
    return xo.load_weights()


@pytest.fixture(scope="session")
def xfb_small():
    """A context for frames up to 128x160 (GPU tests only)."""
    from xfeatslam_b200.capi import XFeatB200
    ctx = XFeatB200(max_h=128, max_w=160, max_batch=4, max_topk=1024)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def xfb_vga():
    from xfeatslam_b200.capi import XFeatB200
    ctx = XFeatB200(max_h=480, max_w=640, max_batch=4, max_topk=4096)
    yield ctx
    ctx.close()
