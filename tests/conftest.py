import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multiple-quadrotor-slam_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "harness"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def _have_cuda_device():
    """GPU tests are skipped on a box without a GPU; on a box WITH one they run -- and fail loudly if the library is
    missing or broken (the product has no CPU fallback)."""
    if any(os.path.exists("/dev/nvidia%d" % i) for i in range(16)):
        return True
    try:
        import triangl_cuda
        return triangl_cuda.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_cuda_device():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run with gpurun / on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- parity reports: counted over ALL points, printed at the end of the run (also with -q) -----------------------------
PARITY_REPORTS = []


def pytest_terminal_summary(terminalreporter):
    if not PARITY_REPORTS:
        return
    terminalreporter.write_sep("-", "parity reports (every point counted; ill-posed / knife-edge points are a separate class)")
    for line in PARITY_REPORTS:
        terminalreporter.write_line(line)
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_reports.txt"), "w") as f:
            f.write("\n".join(PARITY_REPORTS) + "\n")
    except OSError:
        pass
