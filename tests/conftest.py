import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "quadrupedal-agility_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(autouse=True)
def _parity_linear_mode():
    """The product default for the dense layers is the tcgen05 TF32 path (tests/test_ppo_plan_gpu.py, test_linear_gpu.py test it
    against derived TF32 bounds); the fp32 parity tests compare against the reference's fp32 results, so EVERY test starts in
    full-fp32 mode -- per test, so that a test that dies in "tc" mode cannot leak the mode into the tests after it."""
    import torch
    from qa_b200.rsl_rl import linear
    linear.set_mode("fp32")
    # `import bench` switches the framework's TF32 modes on for the whole process; the parity references are fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    linear.set_mode("fp32")
