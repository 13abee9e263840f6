"""mask_pack / mask_overlap kernels against the reference-generated fixture and the oracle:
integer intersection counts and the fp16 maps are compared bit for bit."""
import os

import numpy as np
import pytest
import torch

from cim_b200 import _lib, mask_ops, synth
from oracle import mask_oracle
from conftest import GOLDEN, assert_f16_bits_equal

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def u16(t):
    return t.cpu().numpy().view(np.uint16)


def check_against_oracle(masks_cpu):
    n = masks_cpu.shape[0]
    packed = mask_ops.mask_pack(masks_cpu.to(DEV))
    iou, asy, inter, area = mask_ops.mask_overlap(packed, return_counts=True)
    o_inter, o_area = mask_oracle.overlap_counts(masks_cpu.numpy())
    np.testing.assert_array_equal(inter.cpu().numpy(), o_inter)
    np.testing.assert_array_equal(area.cpu().numpy(), o_area)
    o_iou, o_asy = mask_oracle.maps_from_counts(o_inter, o_area)
    assert_f16_bits_equal(u16(iou), o_iou.view(np.uint16))
    assert_f16_bits_equal(u16(asy), o_asy.view(np.uint16))
    assert iou.shape == (n, n) and iou.dtype == torch.float16


def test_reference_fixture_bit_exact(golden_masks):
    hw = int(golden_masks["hw"])
    masks = np.unpackbits(golden_masks["masks_bits"], axis=1, bitorder="little")[:, :hw]
    t = torch.from_numpy(masks).view(-1, 64, 64)
    packed = mask_ops.mask_pack(t.to(DEV), layout="flat")
    # the flat packed layout is little-endian bit order: identical bytes to the fixture's packbits
    np.testing.assert_array_equal(packed.cpu().numpy().view(np.uint8).reshape(len(masks), -1),
                                  golden_masks["masks_bits"])
    for lay in ("flat", "tiled"):
        iou, asy = mask_ops.mask_overlap(mask_ops.mask_pack(t.to(DEV), layout=lay))
        assert_f16_bits_equal(u16(iou), golden_masks["iou_u16"])       # includes NaN entries
        assert_f16_bits_equal(u16(asy), golden_masks["asy_u16"])


@pytest.mark.parametrize("n,h,w", [(40, 64, 64), (9, 8, 16), (70, 24, 80)])
def test_tiled_layout_is_the_documented_permutation(n, h, w):
    """cim_mask_pack_tiled: pixel (y, x) -> q = ((y>>3)*(W>>4) + (x>>4))*128 + (y&7)*16 + (x&15) (cimhead.h)."""
    masks = (torch.rand(n, h, w, generator=torch.Generator().manual_seed(n)) < 0.5).to(torch.uint8)
    flat = mask_ops.mask_pack(masks.to(DEV), layout="flat").cpu().numpy()
    tiled = mask_ops.mask_pack(masks.to(DEV), layout="tiled").cpu().numpy()
    bits = np.unpackbits(flat.view(np.uint8), axis=1, bitorder="little")[:, :h * w]
    y, x = np.divmod(np.arange(h * w), w)
    q = ((y >> 3) * (w >> 4) + (x >> 4)) * 128 + (y & 7) * 16 + (x & 15)
    want = np.zeros_like(bits)
    want[:, q] = bits
    got = np.unpackbits(tiled.view(np.uint8), axis=1, bitorder="little")[:, :h * w]
    np.testing.assert_array_equal(got, want)
    with pytest.raises(ValueError):
        mask_ops.mask_pack(torch.zeros(2, 12, 16, dtype=torch.uint8, device=DEV), layout="tiled")


@pytest.mark.parametrize("n,side,n_img", [(300, 64, 2), (700, 128, 1)])
def test_flat_and_tiled_layouts_give_identical_maps(n, side, n_img):
    imgs = torch.stack([synth.rasterize(synth.proposal_params(n, side, 40 + b)) for b in range(n_img)]).to(DEV)
    outs = {}
    for lay in ("flat", "tiled"):
        for algo in ("tensor", "popc"):
            outs[lay, algo] = mask_ops.mask_overlap(mask_ops.mask_pack(imgs, layout=lay), return_counts=True, algo=algo)
    ref = outs["flat", "popc"]
    for k, o in outs.items():
        assert torch.equal(o[2], ref[2]) and torch.equal(o[3], ref[3]), k
        assert_f16_bits_equal(u16(o[0]), u16(ref[0]))
        assert_f16_bits_equal(u16(o[1]), u16(ref[1]))
    # the diagnostic counter: visited K-blocks, at most tiles x K-blocks per mask
    nrb, ncb = (n + 127) // 128, (n + 255) // 256
    tiles = sum(ncb - (i >> 1) for i in range(nrb)) * n_img
    for lay in ("flat", "tiled"):
        v = mask_ops.mask_overlap(mask_ops.mask_pack(imgs, layout=lay), algo="tensor", return_visited=True)[-1]
        assert 0 < v <= tiles * (side * side // 128)


@pytest.mark.parametrize("n,h,w", [(130, 64, 64), (65, 37, 50), (1, 8, 8), (200, 96, 96), (64, 5, 5)])
def test_random_proposals_and_ragged_shapes(n, h, w):
    if h == w and h in (64, 96):      # nested synthetic proposals
        masks = synth.rasterize(synth.proposal_params(n, h, n))
    else:                              # unstructured masks, hw not a multiple of 32
        masks = (torch.rand(n, h, w, generator=torch.Generator().manual_seed(n)) < 0.4).to(torch.uint8)
    masks = masks.contiguous()
    if n > 8:
        masks[3] = 0                 # empty mask -> NaN
        masks[4] = 1                 # full mask
        masks[6] = masks[5]          # duplicates -> IoU exactly 1
        masks[7] = masks[7] * 255    # any non-zero byte counts as inside
    check_against_oracle(masks)


def test_bool_input_and_batched_images():
    m = [synth.rasterize(synth.proposal_params(70, 64, 100 + b)) for b in range(3)]
    batch = torch.stack(m).to(DEV)
    packed = mask_ops.mask_pack(batch.bool())
    assert packed.shape == (3, 70, 128)
    iou, asy = mask_ops.mask_overlap(packed)
    for b in range(3):
        o_iou, o_asy = mask_oracle.mask_overlap_maps(m[b].numpy())
        assert_f16_bits_equal(u16(iou[b]), o_iou.view(np.uint16))
        assert_f16_bits_equal(u16(asy[b]), o_asy.view(np.uint16))


def test_full_size_cfg2_properties():
    """One image of BASELINE.json configs[1]: 2000 proposals, 512x512 masks."""
    R = 2000
    params = synth.proposal_params(R, 512, 1234)
    masks = synth.rasterize(params, device=DEV)
    packed = mask_ops.mask_pack(masks)
    iou, asy, inter, area = mask_ops.mask_overlap(packed, return_counts=True)
    assert torch.equal(area.long(), masks.view(R, -1).sum(1))
    assert torch.equal(inter, inter.t())                                  # symmetric
    assert torch.equal(inter.diagonal(), area)                             # |m & m| = |m|
    assert bool((inter <= torch.minimum(area[:, None], area[None, :])).all())
    assert bool((iou.diagonal() == 1).all())
    # 24 random rows against exact integer counts computed independently with torch ops
    rows = torch.randperm(R, generator=torch.Generator().manual_seed(0))[:24].to(DEV)
    sub = masks[rows].view(24, -1).float()
    want = torch.zeros(24, R, device=DEV)
    flat = masks.view(R, -1)
    for s in range(0, R, 250):
        want[:, s:s + 250] = sub @ flat[s:s + 250].float().t()            # exact: counts < 2^24
    assert torch.equal(inter[rows].float(), want)
    o_iou, o_asy = mask_oracle.maps_from_counts(inter[rows].cpu().numpy()[:, rows.cpu().numpy()],
                                                area[rows].cpu().numpy())
    assert_f16_bits_equal(u16(iou[rows][:, rows]), o_iou.view(np.uint16))
    assert_f16_bits_equal(u16(asy[rows][:, rows]), o_asy.view(np.uint16))
    # containment really occurs in the synthetic hierarchy (heads.py:386 needs asy > 0.85)
    assert int((asy > 0.85).sum().item()) > R


@pytest.mark.parametrize("n,side,n_img", [(300, 64, 1), (513, 64, 2), (256, 128, 1), (777, 32, 1), (64, 64, 3)])
def test_tensor_core_path_equals_popc_and_oracle(n, side, n_img):
    """tcgen05.mma.kind::i8 kernel vs the AND+POPC kernel vs the oracle: integer counts and both
    fp16 maps bit-exact, including partial tiles, empty / full / duplicate masks."""
    imgs = []
    for b in range(n_img):
        m = synth.rasterize(synth.proposal_params(n, side, 900 + 7 * b + n))
        m[1] = 0
        m[2] = 1
        m[n - 1] = m[0]
        imgs.append(m)
    packed = mask_ops.mask_pack(torch.stack(imgs).to(DEV))
    t_iou, t_asy, t_inter, t_area = mask_ops.mask_overlap(packed, return_counts=True, algo="tensor")
    p_iou, p_asy, p_inter, p_area = mask_ops.mask_overlap(packed, return_counts=True, algo="popc")
    assert torch.equal(t_inter, p_inter) and torch.equal(t_area, p_area)
    assert_f16_bits_equal(u16(t_iou), u16(p_iou))
    assert_f16_bits_equal(u16(t_asy), u16(p_asy))
    for b in range(n_img):
        o_inter, o_area = mask_oracle.overlap_counts(imgs[b].numpy())
        np.testing.assert_array_equal(t_inter[b].cpu().numpy(), o_inter)
        o_iou, o_asy = mask_oracle.maps_from_counts(o_inter, o_area)
        assert_f16_bits_equal(u16(t_iou[b]), o_iou.view(np.uint16))
        assert_f16_bits_equal(u16(t_asy[b]), o_asy.view(np.uint16))


@pytest.mark.parametrize("variant", [_lib.DBG_OVERLAP_LOADER_WARP, 0])
def test_tensor_core_kernel_repeated_launches(variant):
    """The tensor-core kernel (tcgen05.mma.kind::mxf4, every expander thread prefetching its own row) is bit-identical
    to the popcount kernel, also over repeated launches.  CIM_DBG_OVERLAP_LOADER_WARP selected the loader-warp
    pipeline of the int8 kernel; that variant is gone (the K-block list now holds masks up to 2 Mpixel) and the flag must be harmless."""
    n, side = 600, 128
    m = synth.rasterize(synth.proposal_params(n, side, 4242))
    m[5] = 0
    m[n - 1] = m[3]
    packed = mask_ops.mask_pack(m[None].to(DEV))
    want = mask_ops.mask_overlap(packed, return_counts=True, algo="popc")
    with _lib.debug_flags(variant):
        for _ in range(3):                   # repeated launches: barrier phases, TMEM alloc / dealloc, scale columns
            got = mask_ops.mask_overlap(packed, return_counts=True, algo="tensor")
    assert torch.equal(got[2], want[2]) and torch.equal(got[3], want[3])
    assert_f16_bits_equal(u16(got[0]), u16(want[0]))
    assert_f16_bits_equal(u16(got[1]), u16(want[1]))


@pytest.mark.parametrize("layout", ["tiled", "flat"])
def test_overlap_with_precomputed_mask_meta_is_bit_identical(layout):
    """cim_mask_meta + cim_mask_overlap_meta (areas / K-block bitmaps produced with the masks) == the self-contained
    call, for 3 images at once (the per-image tile ranking is interleaved over the images), both pixel orders."""
    n, side = 520, 128
    ms = []
    for b in range(3):
        m = synth.rasterize(synth.proposal_params(n, side, 900 + b))
        m[b + 1] = 0
        ms.append(m)
    packed = torch.stack([mask_ops.mask_pack(m.to(DEV), layout=layout) for m in ms])
    kb = side // 16 if layout == "tiled" else 0
    want = mask_ops.mask_overlap(packed, return_counts=True, algo="tensor", kb_per_row=kb)
    meta = mask_ops.mask_meta(packed, kb)
    got = mask_ops.mask_overlap(packed, return_counts=True, algo="tensor", kb_per_row=kb, meta=meta)
    pop = mask_ops.mask_overlap(packed, return_counts=True, algo="popc")
    for a, b_, c in zip(got, want, pop):
        v = lambda t: t.view(torch.int16) if t.dtype == torch.float16 else t
        assert torch.equal(v(a)[~torch.isnan(a.float())], v(b_)[~torch.isnan(b_.float())])
        assert torch.equal(torch.isnan(a.float()), torch.isnan(c.float()))
        assert torch.equal(v(a)[~torch.isnan(a.float())], v(c)[~torch.isnan(c.float())])


def test_fused_crop_unpack_and_meta_equals_the_two_separate_calls():
    """cim_mask_unpack_crops_tiled_meta == cim_mask_unpack_crops_tiled (after a memset) + cim_mask_meta, bit for bit,
    incl. empty masks, crops touching the borders and garbage in the destination beforehand."""
    import ctypes as C
    from cim_b200 import _lib
    n_img, n, h, w = 2, 150, 64, 96
    g = torch.Generator().manual_seed(7)
    masks = torch.zeros(n_img * n, h, w, dtype=torch.uint8)
    for i in range(n_img * n):
        if i % 37 == 5:
            continue                                                     # an empty mask
        y0, x0 = int(torch.randint(0, h - 1, (1,), generator=g)), int(torch.randint(0, w - 1, (1,), generator=g))
        y1, x1 = int(torch.randint(y0 + 1, h + 1, (1,), generator=g)), int(torch.randint(x0 + 1, w + 1, (1,), generator=g))
        masks[i, y0:y1, x0:x1] = (torch.rand(y1 - y0, x1 - x0, generator=g) < 0.6).to(torch.uint8)
    c = mask_ops.pack_crops_host(masks)
    crops = mask_ops.MaskCrops(c.words.to(DEV), c.meta.to(DEV), c.off.to(DEV), c.height, c.width)
    want = mask_ops.unpack_crops(crops, layout="tiled").view(n_img, n, -1)
    words = want.shape[-1]
    want_meta = mask_ops.mask_meta(want, w // 16)
    L = _lib.lib()
    got = torch.full_like(want, -1)
    meta = torch.full_like(want_meta, 0x5A)
    _lib.check(L.cim_mask_unpack_crops_tiled_meta(_lib.ptr(crops.words), _lib.ptr(crops.meta), _lib.ptr(crops.off),
                                                  _lib.ptr(got), _lib.ptr(meta), meta.numel(), n_img, n, h, w, words,
                                                  _lib.stream_ptr(torch.device(DEV))), "fused unpack")
    assert torch.equal(got, want)
    nm = n_img * n
    bw = ((words // 4) + 31) // 32
    up = lambda v: (v + 255) & ~255
    a0, k0, b0 = 0, up(nm * 4), up(nm * 4) + up(nm * 16)
    assert torch.equal(meta[a0:a0 + nm * 4], want_meta[a0:a0 + nm * 4])                    # areas
    assert torch.equal(meta[k0:k0 + nm * 16], want_meta[k0:k0 + nm * 16])                  # sort keys
    assert torch.equal(meta[b0:b0 + nm * bw * 4], want_meta[b0:b0 + nm * bw * 4])          # K-block bitmaps
    iou1, asy1 = mask_ops.mask_overlap(got, algo="popc")
    iou2, asy2 = mask_ops.mask_overlap(got, algo="popc", meta=meta, kb_per_row=w // 16)
    assert_f16_bits_equal(u16(iou1), u16(iou2))


def test_sparse_crop_unpack_over_successive_batches_equals_the_dense_unpack():
    """cim_mask_unpack_crops_tiled_meta_sparse updates a zero-initialised buffer batch after batch (clearing the rectangle
    a row held before, writing the new one): after every batch masks and metadata are bit-identical to the dense fused
    unpack of that batch alone -- incl. rows that become empty, rows that were empty, and shrinking / moving crops."""
    from cim_b200 import _lib
    n_img, n, h, w = 2, 120, 64, 96
    L = _lib.lib()
    st = _lib.stream_ptr(torch.device(DEV))
    words = h * w // 32
    packed = torch.zeros(n_img, n, words, dtype=torch.int32, device=DEV)
    prev = torch.zeros(n_img * n, 4, dtype=torch.int32, device=DEV)
    nbytes = L.cim_mask_meta_bytes(n_img, n, words)
    meta = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
    for batch in range(4):
        g = torch.Generator().manual_seed(100 + batch)
        masks = torch.zeros(n_img * n, h, w, dtype=torch.uint8)
        for i in range(n_img * n):
            if (i + batch) % 11 == 3:
                continue                                                 # empty in this batch
            y0, x0 = int(torch.randint(0, h - 1, (1,), generator=g)), int(torch.randint(0, w - 1, (1,), generator=g))
            y1, x1 = int(torch.randint(y0 + 1, h + 1, (1,), generator=g)), int(torch.randint(x0 + 1, w + 1, (1,), generator=g))
            masks[i, y0:y1, x0:x1] = (torch.rand(y1 - y0, x1 - x0, generator=g) < 0.5).to(torch.uint8)
        c = mask_ops.pack_crops_host(masks)
        crops = mask_ops.MaskCrops(c.words.to(DEV), c.meta.to(DEV), c.off.to(DEV), c.height, c.width)
        want = torch.full_like(packed, -1)
        want_meta = torch.full_like(meta, 0x33)
        _lib.check(L.cim_mask_unpack_crops_tiled_meta(_lib.ptr(crops.words), _lib.ptr(crops.meta), _lib.ptr(crops.off),
                                                      _lib.ptr(want), _lib.ptr(want_meta), want_meta.numel(), n_img, n, h,
                                                      w, words, st), "dense unpack")
        _lib.check(L.cim_mask_unpack_crops_tiled_meta_sparse(_lib.ptr(crops.words), _lib.ptr(crops.meta),
                                                             _lib.ptr(crops.off), _lib.ptr(packed), _lib.ptr(prev),
                                                             _lib.ptr(meta), meta.numel(), n_img, n, h, w, words, st),
                   "sparse unpack")
        assert torch.equal(packed, want), f"batch {batch}: masks differ"
        nm = n_img * n
        bw = ((words // 4) + 31) // 32
        up = lambda v: (v + 255) & ~255
        a0, k0, b0 = 0, up(nm * 4), up(nm * 4) + up(nm * 16)
        assert torch.equal(meta[a0:a0 + nm * 4], want_meta[a0:a0 + nm * 4])                    # areas
        assert torch.equal(meta[k0:k0 + nm * 16], want_meta[k0:k0 + nm * 16])                  # sort keys
        assert torch.equal(meta[b0:b0 + nm * bw * 4], want_meta[b0:b0 + nm * bw * 4])          # K-block bitmaps
        assert torch.equal(prev.cpu(), c.meta.view(-1, 4))


def test_tensor_path_rejects_what_it_cannot_take():
    packed = mask_ops.mask_pack(torch.ones(70, 5, 5, dtype=torch.uint8, device=DEV))      # 1 word per mask
    with pytest.raises(RuntimeError, match="shape"):
        mask_ops.mask_overlap(packed, algo="tensor")
    iou, _ = mask_ops.mask_overlap(packed, algo="auto")                                    # falls back to popc
    assert bool((iou == 1).all())


@pytest.mark.parametrize("n,h,w", [(60, 64, 64), (33, 37, 50), (20, 96, 160)])
def test_crop_wire_format_round_trip(n, h, w):
    """bbox-cropped wire format -> device unpack == packing the full byte masks."""
    g = torch.Generator().manual_seed(n)
    masks = torch.zeros(n, h, w, dtype=torch.uint8)
    for i in range(n):                         # random boxes with random content, some touching borders
        y0, x0 = int(torch.randint(0, h - 1, (1,), generator=g)), int(torch.randint(0, w - 1, (1,), generator=g))
        y1, x1 = int(torch.randint(y0 + 1, h + 1, (1,), generator=g)), int(torch.randint(x0 + 1, w + 1, (1,), generator=g))
        masks[i, y0:y1, x0:x1] = (torch.rand(y1 - y0, x1 - x0, generator=g) < 0.6).to(torch.uint8)
    masks[2] = 0                               # empty mask: empty crop
    masks[3] = 1                               # full mask: crop = whole image
    crops = mask_ops.pack_crops_host(masks.numpy())
    dev_crops = mask_ops.MaskCrops(crops.words.to(DEV), crops.meta.to(DEV), crops.off.to(DEV), h, w)
    got = mask_ops.unpack_crops(dev_crops, layout="flat")
    want = mask_ops.mask_pack(masks.to(DEV), layout="flat")
    assert torch.equal(got, want)
    if mask_ops.tiled_ok(h, w):
        assert torch.equal(mask_ops.unpack_crops(dev_crops, layout="tiled"), mask_ops.mask_pack(masks.to(DEV), layout="tiled"))
        assert mask_ops.unpack_crops(dev_crops).cim_kb_per_row == w // 16
    if w % 32 == 0:
        again = mask_ops.crops_from_packed_host(want.cpu(), h, w)
        assert torch.equal(again.words, crops.words) and torch.equal(again.meta, crops.meta)
    assert crops.nbytes < masks.numel()        # smaller than byte masks by construction


# ------------------------------------------------------------- rectangular ratios + on-disk formats (8f-4)
@pytest.mark.parametrize("name", ["n12x3", "n9x1"])
def test_pair_ratio_reference_fixture_bit_exact(name):
    z = np.load(os.path.join(GOLDEN, "mask_pair.npz"))
    a, b = torch.from_numpy(z[f"{name}/a"]).to(DEV), torch.from_numpy(z[f"{name}/b"]).to(DEV)
    for mode, fn in (("iou", mask_ops.mask_iou), ("asymmetric", mask_ops.mask_asymmetric_iou),
                     ("inside", mask_ops.mask_inside), ("outside", mask_ops.mask_outside)):
        got, want = fn(a, b).cpu().numpy(), z[f"{name}/{mode}"]
        np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
        np.testing.assert_array_equal(np.nan_to_num(got).view(np.uint32), np.nan_to_num(want).view(np.uint32))


def test_pair_ratio_n_by_one_like_agpl_label_assign():
    """tools/pre/AGPL_label_assign.py:82-86: 2000 proposals against one averaged peak mask, full resolution."""
    params = synth.proposal_params(2000, 512, 77)
    masks = synth.rasterize(params, device=DEV)
    avg = (masks[masks[:, 200, 260] > 0].float().mean(0) > 0.7).to(torch.uint8)[None]
    got = mask_ops.mask_iou(masks, avg).cpu().numpy()
    want = mask_oracle.pair_ratio(masks.cpu().numpy(), avg.cpu().numpy(), "iou")
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_array_equal(np.nan_to_num(got).view(np.uint32), np.nan_to_num(want).view(np.uint32))
    # and the square case agrees with the training-path maps (fp32 here, fp16 there)
    sq = mask_ops.mask_iou(masks[:300], masks[:300])
    iou16, _ = mask_ops.mask_overlap(mask_ops.mask_pack(masks[:300]))
    assert torch.equal(sq.to(torch.float16), iou16)


def test_map_pickle_and_packed_store_round_trip(tmp_path):
    """The reference's cob_iou pickle format (create_cob_iou.py:48-49, read at model_builder.py:148-156) and the
    bit-packed mask store."""
    import pickle
    masks = synth.rasterize(synth.proposal_params(150, 128, 5), device=DEV)
    packed = mask_ops.mask_pack(masks)
    iou, asy = mask_ops.mask_overlap(packed)
    p = tmp_path / "2007_000032.pkl"
    mask_ops.save_map_pickle(p, iou)
    raw = pickle.load(open(p, "rb"))                      # what the reference would read
    assert raw.dtype == np.float16 and raw.shape == (150, 150)
    assert torch.equal(mask_ops.load_map_pickle(p, DEV).view(torch.int16), iou.view(torch.int16))
    idx = torch.arange(149, -1, -1)
    assert torch.equal(mask_ops.load_map_pickle(p, DEV, idx).view(torch.int16),
                       iou[idx.to(DEV)][:, idx.to(DEV)].contiguous().view(torch.int16))
    q = tmp_path / "masks.npz"
    mask_ops.save_packed_masks(q, packed, 128, 128, "tiled" if mask_ops.tiled_ok(128, 128) else "flat")
    back, h, w, layout = mask_ops.load_packed_masks(q, DEV)
    assert (h, w) == (128, 128) and torch.equal(back, packed)
    iou2, asy2 = mask_ops.mask_overlap(back)
    assert torch.equal(iou2.view(torch.int16), iou.view(torch.int16)) and torch.equal(asy2.view(torch.int16), asy.view(torch.int16))


def test_benchmarked_overlap_kernel_256_sampled_rows_against_oracle_counts():
    """The production instance of the tensor kernel (Cfg<4,2,4>: 2048 K-blocks per row at 512 x 512, tiled pixel order,
    precomputed metadata) on one cfg2 image: 256 sampled rows x all 2000 columns against the oracle's integer counts
    (mask_oracle.overlap_counts on the byte masks), fp16 maps bit-exact from those counts."""
    R = 2000
    params = synth.proposal_params(R, 512, 4321)
    masks = synth.rasterize(params, device=DEV)
    packed = mask_ops.mask_pack(masks)[None]
    meta = mask_ops.mask_meta(packed)
    iou, asy, inter, area = mask_ops.mask_overlap(packed, return_counts=True, algo="tensor", meta=meta)
    rows = np.sort(np.random.RandomState(1).choice(R, 256, replace=False))
    m_np = masks.cpu().numpy().reshape(R, -1)
    o_area = m_np.sum(1, dtype=np.int64)
    np.testing.assert_array_equal(area[0].cpu().numpy(), o_area)
    sub_inter = np.zeros((256, R), np.int64)
    a = m_np[rows].astype(np.float32)
    for s in range(0, R, 500):                                 # exact in fp32: counts < 2^24
        sub_inter[:, s:s + 500] = a @ m_np[s:s + 500].astype(np.float32).T
    small_i, small_a = mask_oracle.overlap_counts(m_np[rows[:16]].reshape(16, 512, 512))   # the oracle's own counting
    np.testing.assert_array_equal(sub_inter[:16][:, rows[:16]], small_i)
    np.testing.assert_array_equal(inter[0].cpu().numpy()[rows], sub_inter)
    with np.errstate(divide="ignore", invalid="ignore"):
        want_iou = (sub_inter.astype(np.float32) / (o_area[rows, None] + o_area[None, :] - sub_inter).astype(np.float32)).astype(np.float16)
        want_asy = (sub_inter.astype(np.float32) / o_area[None, :].astype(np.float32)).astype(np.float16)
    assert_f16_bits_equal(u16(iou[0])[rows], want_iou.view(np.uint16))
    assert_f16_bits_equal(u16(asy[0])[rows], want_asy.view(np.uint16))
