"""oracle/mask_oracle.py against the fixture produced by the reference's own mask_utils.py."""
import os

import numpy as np
import pytest

from oracle import mask_oracle, roi_oracle
from conftest import GOLDEN


def unpack(golden_masks):
    bits = golden_masks["masks_bits"]
    hw = int(golden_masks["hw"])
    return np.unpackbits(bits, axis=1, bitorder="little")[:, :hw]


def test_restatement_matches_reference_fixture(golden_masks):
    masks = unpack(golden_masks)
    iou, asy = mask_oracle.mask_overlap_maps(masks)
    np.testing.assert_array_equal(iou.view(np.uint16), golden_masks["iou_u16"])
    np.testing.assert_array_equal(asy.view(np.uint16), golden_masks["asy_u16"])


def test_literal_loops_match_fixture(golden_masks):
    masks = unpack(golden_masks)[:16]
    iou, asy = mask_oracle.mask_overlap_maps_literal(masks)
    np.testing.assert_array_equal(iou.view(np.uint16), golden_masks["iou_u16"][:16, :16])
    np.testing.assert_array_equal(asy.view(np.uint16), golden_masks["asy_u16"][:16, :16])


def test_c_counts_match_numpy(golden_masks):
    masks = unpack(golden_masks)
    inter, area = mask_oracle.overlap_counts(masks)
    ci, ca = roi_oracle.mask_counts(masks)
    np.testing.assert_array_equal(ci, inter)
    np.testing.assert_array_equal(ca, area)


def test_single_fp32_division_equals_reference_rounding_chain():
    """int/int in float64 -> float32 (the reference) == one IEEE float32 division (the kernel),
    exhaustively for all count pairs up to 1200 and on random large counts."""
    i = np.arange(0, 1201, dtype=np.int64)
    a, b = np.meshgrid(i, i[1:], indexing="ij")
    ref = (a.astype(np.float64) / b.astype(np.float64)).astype(np.float32)
    one = a.astype(np.float32) / b.astype(np.float32)
    np.testing.assert_array_equal(ref.view(np.uint32), one.view(np.uint32))
    rng = np.random.RandomState(0)
    a = rng.randint(0, 1 << 18, 2_000_000).astype(np.int64)
    b = a + rng.randint(0, 1 << 18, 2_000_000).astype(np.int64) + 1
    ref = (a.astype(np.float64) / b.astype(np.float64)).astype(np.float32)
    one = a.astype(np.float32) / b.astype(np.float32)
    np.testing.assert_array_equal(ref.view(np.uint32), one.view(np.uint32))


def test_known_answers():
    m = np.zeros((4, 8, 8), np.uint8)
    m[0, :4, :4] = 1            # 16 px
    m[1, :2, :2] = 1            # 4 px, nested in 0
    m[2, 4:, 4:] = 1            # disjoint from 0 and 1
    m[3] = m[0]                 # identical to 0
    iou, asy = mask_oracle.mask_overlap_maps(m)
    assert iou[0, 3] == 1 and iou[0, 2] == 0 and iou[0, 1] == np.float16(0.25)
    assert asy[0, 1] == 1 and asy[1, 0] == np.float16(0.25)      # asy[i,j] = |i&j| / |j|


@pytest.mark.parametrize("name", ["n12x3", "n9x1"])
@pytest.mark.parametrize("mode", ["iou", "asymmetric", "inside", "outside"])
def test_pair_ratio_matches_reference_mask_utils(name, mode):
    """tests/golden/mask_pair.npz: outputs of the reference's own lib/utils/mask_utils.py."""
    z = np.load(os.path.join(GOLDEN, "mask_pair.npz"))
    got = mask_oracle.pair_ratio(z[f"{name}/a"], z[f"{name}/b"], mode)
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got.view(np.uint32) * ~np.isnan(got), z[f"{name}/{mode}"].view(np.uint32) * ~np.isnan(got))
    np.testing.assert_array_equal(np.isnan(got), np.isnan(z[f"{name}/{mode}"]))
