import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture
def first_rows():
    """Row numbering of SparseConvNet (first appearance in the point list) for tests that compare rows with the oracle
    one to one; the product default is Morton order, compared through the canonical sort (test_gpu_baseline_size.py,
    test_gpu_morton.py)."""
    from sparse_rcnn_b200 import scn
    prev = scn.get_row_order()
    scn.set_row_order("first")
    yield
    scn.set_row_order(prev)
