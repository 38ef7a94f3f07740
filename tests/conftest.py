import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def assert_close(got, ref, rtol=1e-5, atol=1e-6, what=""):
    """The float parity gate of the path: |got - ref| <= atol + rtol * |ref|."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    err = np.abs(got - ref)
    bound = atol + rtol * np.abs(ref)
    bad = err > bound
    if bad.any():
        i = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError("%s: %d of %d outside tolerance; worst at %s: got %r ref %r (|d|=%.3g, bound %.3g)"
                             % (what, int(bad.sum()), bad.size, i, got[i], ref[i], err[i], bound[i]))


def assert_exact(got, ref, what=""):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    bad = got != ref
    if bad.any():
        i = np.unravel_index(np.argmax(bad), bad.shape)
        raise AssertionError("%s: %d of %d differ; first at %s: got %r ref %r" % (what, int(bad.sum()), bad.size, i, got[i], ref[i]))
