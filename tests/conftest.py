import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A GPU test that hangs (a kernel that never ends) must end the run, not sit on the GPU box until somebody kills it:
    with pytest-timeout present every GPU test gets 15 minutes, enforced from a watchdog thread (it can end the process
    while the main thread is stuck inside a CUDA call)."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(900, method="thread"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py


def pytest_terminal_summary(terminalreporter):
    """Tie / grazing counts of every closest-hit comparison of the run (tests/gpu_util.py SOFT_LOG)."""
    mod = sys.modules.get("tests.gpu_util")
    log = getattr(mod, "SOFT_LOG", None) if mod else None
    if not log:
        return
    terminalreporter.write_sep("-", "closest-hit comparisons against the oracle: rays / exact ids / inside the tie band / grazing (allowed)")
    for what, n, exact, tie, grazing, allowed in log:
        terminalreporter.write_line(f"{what:60s} {n:9d} {exact:9d} {tie:6d} {grazing:6d}  ({allowed})")
