import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _restore_library_math_flags():
    """Tests switch torch's global tf32 flags (cuBLAS / cuDNN library math around the path); none may leak into the next."""
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev


@pytest.fixture(scope='session')
def golden():
    """Fixtures generated from the reference's own unmodified files (tests/golden/make_golden.py)."""
    path = os.path.join(ROOT, 'tests', 'golden', 'hotpath_golden.npz')
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def params(golden, prefix):
    return {k[len(prefix):]: v for k, v in golden.items() if k.startswith(prefix)}


def rel_err(a, b):
    """max |a-b| / max |b| -- the relative-error measure the parity bar is stated in."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    # NaN rows are part of the reference's behaviour (zero-norm vectors are divided without an epsilon):
    # they must appear at the same places; the rest is compared numerically.
    assert torch.equal(torch.isnan(a), torch.isnan(b)), 'NaN pattern differs'
    a, b = torch.nan_to_num(a, nan=0.0), torch.nan_to_num(b, nan=0.0)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope='session')
def rpn_golden():
    """RPN proposal-stage vectors made by the reference's RPNHead._get_bboxes (tests/golden/make_rpn_golden.py)."""
    import numpy as np
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'rpn_golden.npz')
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) for k in z.files}
