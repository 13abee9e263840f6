"""The window plan of the RoIAlign window-tile path (cim_b200/csrc/roi_window.cuh: maps larger than the shared-memory
tile -- VGG-16 64 x 64 .. 150 x 150, ResNet-50 43 .. 75 cells) evaluated on the CPU through cim_debug_roi_window_plan,
the same __host__ __device__ code the prep kernels run.  The forward is rebuilt from the sub-ROI descriptors alone
(window origin + slot weights + slot -> bin maps), exactly the arithmetic the tile kernels apply to them, and compared
with the oracle (oracle/roi_oracle.c, the reference's roi_align_kernel.cu:16-121 + mmcv's `aligned`); the derived tables
the backward kernel relies on (row-pair lists, owner masks, the strictly-increasing flag) are checked for consistency.
No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from cim_b200 import _lib
from oracle import roi_oracle

DESC_WORDS = 168
D_ROI, D_FLAGY, D_FLAGX, D_TX, D_Y0, D_Y1, D_XINC, D_OWN, D_YLO, D_YMAP, D_YN, D_FLAGS, D_XLO, D_XMAP, D_PHR, D_WY, D_WX = \
    0, 1, 2, 3, 4, 5, 6, 7, 8, 15, 16, 23, 24, 31, 32, 40, 112


def plan(rois, B, H, W, scale, sr=0, aligned=True, cap_per_roi=49):
    K = len(rois)
    rois = np.ascontiguousarray(rois, np.float32)
    desc = np.zeros((K * cap_per_roi, DESC_WORDS), np.int32)
    keys = np.zeros(K * cap_per_roi, np.int32)
    flags = np.zeros(K, np.int32)
    geom = np.zeros(6, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = _lib.lib().cim_debug_roi_window_plan(p(rois), K, B, H, W, float(scale), sr, int(aligned), p(desc), p(keys),
                                             K * cap_per_roi, p(flags), p(geom))
    assert n >= 0, n
    return desc[:n], keys[:n], flags, geom


def rebuild_forward(feat, rois, desc, keys, geom):
    """out [K, C, 7, 7] from the sub-ROI descriptors; also returns how often each (roi, bin) was written."""
    Hp, wh, ww, nwy, nwx, G = (int(v) for v in geom)
    B, Cc, H, W = feat.shape
    K = len(rois)
    out = np.zeros((K, Cc, 7, 7), np.float64)
    written = np.zeros((K, 7, 7), np.int32)
    fpad = np.zeros((B, Cc, Hp, W), np.float64)
    fpad[:, :, :H] = feat
    f32 = lambda a: a.view(np.float32)
    for d, key in zip(desc, keys):
        roi = int(d[D_ROI])
        b = int(rois[roi, 0])
        wy, wx = divmod(int(key), nwx)
        oy, ox = min(wy * G, Hp - wh), min(wx * G, W - ww)
        assert oy % 2 == 0
        win = fpad[b, :, oy:oy + wh, ox:ox + ww]
        ym, xm = int(np.uint32(d[D_YMAP])), int(np.uint32(d[D_XMAP]))
        ny, nx, T = (ym >> 28) & 7, (xm >> 28) & 7, int(d[D_TX])
        assert d[D_FLAGY] == 0 and d[D_FLAGX] == 0 and T in (2, 3, 4, 6, 8)
        virt = np.zeros((7, 7, Cc))
        rows, cols = [], []
        for s in range(7):
            ylo, yn = int(d[D_YLO + s]), int(d[D_YN + s])
            wyv = f32(d[D_WY + s * 10:D_WY + s * 10 + 10])
            assert wyv[0] == 0 and wyv[9] == 0 and 0 <= ylo and ylo + yn <= wh and yn <= 8
            assert not wyv[1 + yn:9].any()
            rows.append((ylo, yn, wyv[1:1 + yn].astype(np.float64)))
            xlo = int(d[D_XLO + s])
            wxv = f32(d[D_WX + s * 8:D_WX + s * 8 + 8])
            assert 0 <= xlo and xlo + T <= ww and not wxv[T:].any()           # the forward reads exactly T taps
            cols.append((xlo, wxv[:T].astype(np.float64)))
            if s >= ny:
                assert yn == 0 and ((ym >> (4 * s)) & 15) == 15
            if s >= nx:
                assert not wxv.any() and ((xm >> (4 * s)) & 15) == 15
        for ys in range(ny):
            ylo, yn, wyv = rows[ys]
            if yn == 0:
                continue
            rowsum = np.tensordot(wyv, win[:, ylo:ylo + yn, :], axes=(0, 1))              # [C, ww]
            for xs in range(nx):
                xlo, wxv = cols[xs]
                virt[ys, xs] = rowsum[:, xlo:xlo + T] @ wxv
        simple = bool(d[D_FLAGS] & 1)
        if simple:
            assert ny == 7 and nx == 7 and [(ym >> (4 * s)) & 15 for s in range(7)] == list(range(7)) \
                and [(xm >> (4 * s)) & 15 for s in range(7)] == list(range(7))
        seen = set()
        for ys in range(ny):
            for xs in range(nx):
                by, bx = (ym >> (4 * ys)) & 15, (xm >> (4 * xs)) & 15
                out[roi, :, by, bx] += virt[ys, xs]
                seen.add((by, bx))
        for by, bx in seen:
            written[roi, by, bx] += 1
        # slots of one bin are consecutive and bins ascend (the epilogue's run-length sum relies on it)
        for m, n_ in ((ym, ny), (xm, nx)):
            bins = [(m >> (4 * s)) & 15 for s in range(n_)]
            assert bins == sorted(bins)
        # ---- tables of the backward kernel
        y0, y1 = int(d[D_Y0]), int(d[D_Y1])
        used = [(lo, n) for lo, n, _ in rows[:ny] if n > 0]
        if used:
            assert y0 == min(lo for lo, n in used) and y1 == max(lo + n for lo, n in used)
        own = 0
        for y in range(y0, y1):
            own |= 1 << ((y >> 1) & 15)
        assert int(d[D_OWN]) == own
        phr = d[D_PHR:D_PHR + 8].view(np.uint8)
        for j in range((y1 + 1) // 2 - y0 // 2):
            pr = y0 // 2 + j
            meet = [s for s in range(ny) if rows[s][1] > 0 and rows[s][0] <= 2 * pr + 1 and 2 * pr < rows[s][0] + rows[s][1]]
            first, cnt = int(phr[j]) & 15, int(phr[j]) >> 4
            assert meet == list(range(first, first + cnt)), (meet, first, cnt)
        xl = [cols[s][0] for s in range(nx)]
        assert int(d[D_XINC]) == int(all(a < b for a, b in zip(xl, xl[1:])))
    return out, written


CASES = [
    # B, C, H, W, scale, K, roi size range (fraction of the map), wild
    (2, 3, 64, 64, 1 / 8, 60, (0.05, 1.0), False),      # VGG-16 at 512 px (cfg3)
    (1, 2, 75, 75, 1 / 16, 50, (0.05, 1.1), True),      # ResNet-50 at scale 1200
    (1, 2, 43, 57, 1 / 16, 50, (0.02, 1.0), False),     # odd sizes, H odd -> padded row
    (1, 2, 150, 112, 1 / 8, 40, (0.1, 1.0), False),     # VGG-16 at scale 1200: bins of up to 23 taps (3 slots)
    (1, 2, 20, 90, 1 / 8, 40, (0.05, 1.0), False),      # only one axis larger than the window
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("aligned,sr", [(True, 0), (False, 0), (True, 2)])
def test_forward_rebuilt_from_window_descriptors_matches_oracle(case, aligned, sr):
    B, Cc, H, W, scale, K, (lo, hi), wild = case
    rng = np.random.RandomState(H * 1000 + W + int(aligned) + 7 * sr)
    feat = rng.randn(B, Cc, H, W).astype(np.float32)
    sw, sh = W / scale, H / scale
    w = rng.uniform(lo, hi, K) * sw
    h = rng.uniform(lo, hi, K) * sh
    x1 = rng.uniform(-0.1 if wild else 0, 1, K) * np.maximum(sw - w, 1)
    y1 = rng.uniform(-0.1 if wild else 0, 1, K) * np.maximum(sh - h, 1)
    rois = np.stack([rng.randint(0, B, K), x1, y1, x1 + w, y1 + h], 1).astype(np.float32)
    rois[0, 1:] = [0, 0, sw, sh]                                  # the whole map
    rois[1, 1:] = [sw * 0.4, sh * 0.4, sw * 0.4 + 3, sh * 0.4 + 2]  # tiny
    if wild:
        rois[2, 1:] = [-sw, -sh, -sw / 2, -sh / 2]                # entirely outside: all bins empty
        rois[3, 1:] = [sw * 0.9, sh * 0.9, sw * 1.5, sh * 1.4]    # hanging over the far corner
    desc, keys, flags, geom = plan(rois, B, H, W, scale, sr, aligned)
    got, written = rebuild_forward(feat, rois, desc, keys, geom)
    want = roi_oracle.roi_align_fwd(feat, rois, 7, 7, scale, sr, aligned)
    ok = flags == 0
    assert ok.sum() >= K * 0.8, "the window path should take (nearly) every ROI of these maps"
    assert (written[ok] == 1).all(), "every bin of a planned ROI is owned by exactly one sub-ROI"
    assert (written[~ok] == 0).all()
    scale_ = np.abs(want).max()
    assert np.abs(got[ok] - want[ok]).max() <= 1e-5 * scale_
    # the whole-map ROI needs several sub-ROIs on maps larger than one window, the tiny one exactly one
    per_roi = np.bincount(desc[:, D_ROI], minlength=K)
    assert per_roi[1] == 1
    if H > 32 or W > 32:
        assert per_roi[0] > 1


def test_plan_is_one_simple_descriptor_per_roi_when_every_roi_fits_one_window():
    """Small ROIs on a large map: one sub-ROI each, flagged SIMPLE (identity slot maps -> the kernel's bulk store)."""
    rng = np.random.RandomState(3)
    K, H, W, scale = 64, 64, 64, 1 / 8
    x1, y1 = rng.uniform(0, 400, K), rng.uniform(0, 400, K)
    rois = np.stack([np.zeros(K), x1, y1, x1 + rng.uniform(8, 100, K), y1 + rng.uniform(8, 100, K)], 1).astype(np.float32)
    desc, keys, flags, geom = plan(rois, 1, H, W, scale)
    assert len(desc) == K and not flags.any() and (desc[:, D_FLAGS] & 1).all()
    assert sorted(desc[:, D_ROI].tolist()) == list(range(K))
