import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Which branch of helpers.assert_parity decided each comparison: 'direct' = within tol of the fp32 oracle,
    'arbiter' = not, but no farther from the fp64 oracle than 4x the fp32 oracle itself is, 'family' = no farther
    from the fp64 oracle than the fp32 oracle's worst tensor of the same model (helpers.assert_grads_parity)."""
    try:
        from helpers import parity_summary
    except Exception:
        return
    n, worst = parity_summary()
    if sum(n.values()) == 0:
        return
    terminalreporter.write_line(f"assert_parity: {n['direct']} direct, {n['arbiter']} by fp64 arbiter, {n['family']} by "
                                f"family bound (deep-model gradients), {n['FAIL']} failed")
    for what, tol, e32, e64, eref, branch in worst:
        terminalreporter.write_line(f"  [{branch}] {what}: |cuda-oracle32| {e32:.2e} |cuda-exact| {e64:.2e} "
                                    f"|oracle32-exact| {eref:.2e} (tol {tol:.0e})")
