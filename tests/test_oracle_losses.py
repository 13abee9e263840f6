"""The loss-block restatement (oracle/loss_oracle.py) against the fixtures produced by the reference's own
heads.cls_iou_loss / heads.mil_bag_loss with autograd (tests/golden/head_losses.npz)."""
import os

import numpy as np
import pytest

from oracle import loss_oracle
from conftest import GOLDEN, cim_case_names

NPZ = np.load(os.path.join(GOLDEN, "head_losses.npz"))
NAMES = cim_case_names(NPZ)


def case(name):
    g = lambda k: NPZ[f"{name}/{k}"]
    return dict(scores=g("scores"), pl=g("pseudo_labels"), pi=g("pseudo_iou_u16").view(np.float16), lw=g("loss_weights"),
                valid=g("valid"), labels=g("labels"), losses=g("losses"), grad=g("grad"))


@pytest.mark.parametrize("name", NAMES)
def test_losses_and_gradients_match_reference(name):
    c = case(name)
    losses, grad = loss_oracle.head_losses(c["scores"], c["pl"], c["pi"], c["lw"], c["valid"], c["labels"], 3)
    np.testing.assert_array_equal(np.isnan(losses), np.isnan(c["losses"]))
    np.testing.assert_allclose(np.nan_to_num(losses), np.nan_to_num(c["losses"]), rtol=1e-5, atol=1e-8)
    np.testing.assert_array_equal(np.isnan(grad), np.isnan(c["grad"]))
    fin = np.isfinite(c["grad"])
    assert np.abs(grad - c["grad"])[fin].max() <= 1e-5 * max(np.abs(c["grad"][fin]).max(), 1e-3)


def test_skipped_layers_contribute_nothing():
    c = case("voc_r64_none")
    assert not c["valid"].any()
    assert (c["losses"][:, :3] == 0).all() and c["losses"][0, 3, 2] > 0
    assert (c["grad"][2:] == 0).all() and (c["grad"][:2] != 0).any()


PCL = np.load(os.path.join(GOLDEN, "pcl_loss.npz"))


@pytest.mark.parametrize("name", cim_case_names(PCL))
def test_pcl_loss_matches_reference(name):
    """tests/golden/pcl_loss.npz: the reference's own heads.PCL_loss + autograd."""
    loss, grad = loss_oracle.pcl_losses(PCL[f"{name}/predict_cls"], PCL[f"{name}/mat"][None])
    assert abs(loss[0] - float(PCL[f"{name}/loss"])) <= 1e-5 * max(abs(float(PCL[f"{name}/loss"])), 1e-3)
    want = PCL[f"{name}/grad"]
    assert np.abs(grad - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-3)
