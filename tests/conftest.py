import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_reference: needs the unmodified reference under /root/reference")
    # the unmodified reference indexes tensors with lists (episode_buffer.py) and builds tensors from lists of ndarrays
    config.addinivalue_line("filterwarnings", "ignore:Using a non-tuple sequence:UserWarning")
    config.addinivalue_line("filterwarnings", "ignore:Creating a tensor from a list of numpy:UserWarning")
    config.addinivalue_line("filterwarnings", "ignore:invalid escape sequence:SyntaxWarning")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a CPU-only box skips the gpu-marked tests instead of failing in them."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    have_lib = os.path.exists(os.path.join(ROOT, "homophily_marl_b200", "libssd_b200.so"))
    if have_gpu and have_lib:
        return
    why = "no CUDA device" if not have_gpu else "libssd_b200.so has not been built"
    skip = pytest.mark.skip(reason=why + " (the product path has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    from oracle import oracle
    oracle.build()
    yield


@pytest.fixture(scope="session", autouse=True)
def _checked_build_reports_no_violation():
    """With SSD_B200_CHECKED=1 the whole GPU suite runs on libssd_b200_check.so; no index violation may be seen."""
    yield
    if os.environ.get("SSD_B200_CHECKED") == "1":
        import torch
        if torch.cuda.is_available():
            from homophily_marl_b200 import _capi
            n = _capi.load().ssd_debug_oob_count()
            assert n == 0, f"checked build counted {n} shared-memory index violations"
