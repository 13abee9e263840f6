"""The C restatement (oracle/roi_oracle.c) against torchvision's CPU ops, which implement the same
Detectron-derived algorithms as the absent mmcv ops."""
import numpy as np
import pytest
import torch
from torchvision.ops import roi_align as tv_roi_align, roi_pool as tv_roi_pool

from oracle import roi_oracle


def make_case(seed, B=2, C=5, H=13, W=17, K=40, scale=0.25, wild=False):
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(B, C, H, W, generator=g)
    b = torch.randint(0, B, (K,), generator=g).float().sort().values
    span_w, span_h = W / scale, H / scale
    x1 = torch.rand(K, generator=g) * span_w * 0.8
    y1 = torch.rand(K, generator=g) * span_h * 0.8
    w = torch.rand(K, generator=g) * span_w * 0.6 + 0.5
    h = torch.rand(K, generator=g) * span_h * 0.6 + 0.5
    if wild:                      # boxes sticking out of the image, tiny and degenerate boxes
        x1 -= span_w * 0.3
        y1 -= span_h * 0.3
        w[::7] = 0.01
        h[::5] = 0.0
    rois = torch.stack([b, x1, y1, x1 + w, y1 + h], 1)
    return feat, rois


@pytest.mark.parametrize("aligned", [True, False])
@pytest.mark.parametrize("sr", [0, 2])
@pytest.mark.parametrize("wild", [False, True])
def test_roi_align_fwd_bwd_matches_torchvision(aligned, sr, wild):
    feat, rois = make_case(3 + sr, wild=wild)
    feat.requires_grad_(True)
    ref = tv_roi_align(feat, rois, (7, 7), 0.25, sr, aligned)
    out = roi_oracle.roi_align_fwd(feat.detach().numpy(), rois.numpy(), 7, 7, 0.25, sr, aligned)
    np.testing.assert_allclose(out, ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(9))
    ref.backward(g)
    gf = roi_oracle.roi_align_bwd(g.numpy(), rois.numpy(), feat.shape, 0.25, sr, aligned)
    np.testing.assert_allclose(gf, feat.grad.numpy(), rtol=1e-5, atol=1e-5)


def test_roi_align_non_square_output():
    feat, rois = make_case(11)
    ref = tv_roi_align(feat, rois, (3, 5), 0.25, 0, True)
    out = roi_oracle.roi_align_fwd(feat.numpy(), rois.numpy(), 3, 5, 0.25, 0, True)
    np.testing.assert_allclose(out, ref.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("wild", [False, True])
def test_roi_pool_matches_torchvision(wild):
    feat, rois = make_case(5, wild=wild)
    feat.requires_grad_(True)
    ref = tv_roi_pool(feat, rois, (7, 7), 0.25)
    out, arg = roi_oracle.roi_pool_fwd(feat.detach().numpy(), rois.numpy(), 7, 7, 0.25)
    np.testing.assert_array_equal(out, ref.detach().numpy())
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
    ref.backward(g)
    gf = roi_oracle.roi_pool_bwd(g.numpy(), arg, rois.numpy(), feat.shape)
    np.testing.assert_allclose(gf, feat.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_roi_pool_mmcv_variant_hand_computed_and_where_it_meets_the_legacy_bins():
    """mmcv 1.x's RoIPool bins (what lib/ops/__init__.py:6 imports; mmcv is absent, so there is nothing to run it
    against: the restatement is pinned by a hand-computed case and by the case where both conventions must agree).
    ROI (1.3, 2.6, 5.2, 6.9) at scale 0.5 -> x in [0.65, 3.1], y in [1.3, 3.95]; 2 x 2 bins of 1.225 x 1.325:
    columns [0, 2) and [1, 4), rows [1, 3) and [2, 4); on feat = arange(64) the maxima are the bins' last elements."""
    feat = np.arange(64, dtype=np.float32).reshape(1, 1, 8, 8)
    rois = np.array([[0, 1.3, 2.6, 5.2, 6.9]], np.float32)
    out, arg = roi_oracle.roi_pool_fwd(feat, rois, 2, 2, 0.5, variant="mmcv")
    assert out.reshape(-1).tolist() == [17, 19, 25, 27] and arg.reshape(-1).tolist() == [17, 19, 25, 27]
    legacy, _ = roi_oracle.roi_pool_fwd(feat, rois, 2, 2, 0.5)
    assert legacy.reshape(-1).tolist() != out.reshape(-1).tolist()          # rounded corners: other bins
    # degenerate ROI (x2 + 1 <= x1): mmcv pools nothing, the legacy kernel forces a 1 x 1 ROI
    deg = np.array([[0, 6, 2, 4.5, 5]], np.float32)
    o, a = roi_oracle.roi_pool_fwd(feat, deg, 2, 2, 0.5, variant="mmcv")
    assert not o.any() and (a == -1).all()
    # integer corners at scale 1: (x2 + 1) - x1 == x2 - x1 + 1 and floor(p * bin + y1) == floor(p * bin) + y1
    rng = np.random.RandomState(0)
    f = rng.randn(2, 3, 20, 24).astype(np.float32)
    x1, y1 = rng.randint(-3, 20, 40), rng.randint(-3, 16, 40)
    r = np.stack([rng.randint(0, 2, 40), x1, y1, x1 + rng.randint(0, 12, 40), y1 + rng.randint(0, 12, 40)], 1).astype(np.float32)
    om, am = roi_oracle.roi_pool_fwd(f, r, 7, 7, 1.0, variant="mmcv")
    ol, al = roi_oracle.roi_pool_fwd(f, r, 7, 7, 1.0)
    np.testing.assert_array_equal(om, ol)
    np.testing.assert_array_equal(am, al)
    np.testing.assert_array_equal(om, tv_roi_pool(torch.from_numpy(f), torch.from_numpy(r), (7, 7), 1.0).numpy())


def test_roi_align_empty():
    feat, _ = make_case(1)
    out = roi_oracle.roi_align_fwd(feat.numpy(), np.zeros((0, 5), np.float32), 7, 7, 0.25)
    assert out.shape == (0, 5, 7, 7)
