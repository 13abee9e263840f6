import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_cim():
    return np.load(os.path.join(GOLDEN, "cim_layer.npz"))


@pytest.fixture(scope="session")
def golden_masks():
    return np.load(os.path.join(GOLDEN, "mask_overlap.npz"))


@pytest.fixture(scope="session")
def golden_scores():
    return np.load(os.path.join(GOLDEN, "score_heads.npz"))


def assert_close_elementwise(got, want, rtol=1e-5, atol_rms=1e-5, what=""):
    """ELEMENT-WISE tolerance next to the max-norm bounds of the test modules (which divide by max|want| and would
    let elements far below the maximum be wrong by large factors): |got - want| <= rtol * |want| + atol_rms * rms(want)
    for every element.  The rms term is what a sum of O(10..1000) fp32 products can lose to cancellation; NaN must
    match NaN."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan), f"{what}: NaN positions differ"
    w = np.where(nan, 0.0, want)
    rms = float(np.sqrt(np.mean(w * w))) if w.size else 0.0
    err = np.abs(np.where(nan, 0.0, got) - w)
    bound = rtol * np.abs(w) + atol_rms * rms
    bad = err > bound
    if bad.any():
        i = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError(f"{what}: {int(bad.sum())} of {err.size} elements outside rtol={rtol:g} + {atol_rms:g}*rms "
                             f"(rms {rms:.3e}); worst at {i}: got {got[i]!r}, want {want[i]!r}, err {err[i]:.3e}, "
                             f"bound {bound[i]:.3e}")


def cim_case_names(npz):
    return sorted({k.split("/")[0] for k in npz.files})


def assert_f16_bits_equal(got_u16, want_u16):
    """Bit-exact comparison of float16 payloads, except that any NaN equals any NaN: x86 0/0
    yields 0xFE00 (the reference's numpy), CUDA yields 0x7FFF; the payload carries no meaning."""
    got_u16, want_u16 = np.asarray(got_u16), np.asarray(want_u16)
    is_nan = lambda u: ((u & 0x7C00) == 0x7C00) & ((u & 0x03FF) != 0)
    gn, wn = is_nan(got_u16), is_nan(want_u16)
    np.testing.assert_array_equal(gn, wn, err_msg="NaN positions differ")
    np.testing.assert_array_equal(np.where(gn, 0, got_u16), np.where(wn, 0, want_u16))
