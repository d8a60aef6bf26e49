import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')
    config.addinivalue_line('markers', 'segment_mode(mode): GGLB200_SEGMENT policy for this GPU test (default in tests: csr)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _segment_mode(request):
    """GPU tests run the unsorted segment ops on the CSR path (fp32 sums bit-identical to the serial reference loop) unless a
    test asks for another policy with @pytest.mark.segment_mode('auto' | 'fast' | 'oneshot'): the sort-free path is covered by
    its own tests (tests/test_gpu_oneshot.py)."""
    if 'gpu' not in request.keywords:
        yield
        return
    from gammagl_b200 import mpops
    m = request.node.get_closest_marker('segment_mode')
    old = mpops._SEGMENT_MODE
    mpops._SEGMENT_MODE = m.args[0] if m else 'csr'
    try:
        yield
    finally:
        mpops._SEGMENT_MODE = old
