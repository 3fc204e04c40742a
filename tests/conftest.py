import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A GPU test that does not come back (a kernel waiting on a barrier that never flips, a collective a peer never entered)
    would sit there until whatever launched pytest gives up — minutes of a GPU box per occurrence.  With pytest-timeout installed
    every GPU test gets 15 minutes (the largest takes under two); the thread method ends the whole run, which is the right
    outcome: the context is wedged anyway."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") is not None and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(900, method="thread"))


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything is built in-tree once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()
