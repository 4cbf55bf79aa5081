import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def glass_lib():
    """The C-ABI library; GPU tests call the product only through it."""
    from glass_text_spotting_b200 import lib
    return lib.load()


def pytest_sessionfinish(session, exitstatus):
    """Every tolerance check of the session (tap name, max error, misses of the literal / scaled bound) goes to
    gpurun_out/parity_report.json, so the status of each tap is on record next to the pass/fail verdict."""
    try:
        import parity_common
        if parity_common.REPORT:
            parity_common.dump_report(os.path.join(ROOT, "gpurun_out", "parity_report.json"))
    except Exception:
        pass
