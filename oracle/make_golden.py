"""oracle/make_golden.py -- pin the oracle: run the REFERENCE's own code on seeded inputs
and store inputs + outputs as small fixtures under tests/golden/.

Run here (the build container, where /root/reference is mounted):
    python oracle/make_golden.py
It imports, UNMODIFIED, from /root/reference:
    lib/modeling/heads.py   (cls_iou_model, CIM_layer)            -- runs on CPU tensors
    lib/utils/mask_utils.py (mask_iou, mask_asymmetric_iou)       -- behind a 3-line shim that
        makes `chainer.backends.cuda.get_array_module` return numpy (chainer/cupy are absent)
    lib/utils/cython_nms.pyx (nms)                                -- compiled into oracle/_ref by
        oracle/build_ref_nms.py (two numpy type names respelled for numpy 2)
The fixtures travel to the GPU box; /root/reference does not.  While generating, every
fixture is also replayed through the oracle restatement and any mismatch aborts.
"""
import io
import os
import sys
import types
import contextlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("CIM_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from cim_b200 import synth                      # noqa: E402
from oracle import heads_oracle, mask_oracle    # noqa: E402


def load_reference():
    chainer = types.ModuleType("chainer")
    backends = types.ModuleType("chainer.backends")
    cuda = types.ModuleType("chainer.backends.cuda")
    cuda.get_array_module = lambda *a: np
    backends.cuda = cuda
    chainer.backends = backends
    sys.modules.update({"chainer": chainer, "chainer.backends": backends,
                        "chainer.backends.cuda": cuda})
    sys.path.insert(0, os.path.join(REF, "lib"))
    import importlib.util

    def imp(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    return imp("ref_heads", "lib/modeling/heads.py"), imp("ref_mask_utils", "lib/utils/mask_utils.py")


def u16(a):
    return np.ascontiguousarray(a).view(np.uint16)


def reference_maps(ref_mu, masks):
    """tools/pre/create_cob_iou.py:43-48 / create_cob_asy_iou.py:43-51 with numpy for cupy."""
    n = len(masks)
    iou_cols, asy_cols = [], []
    with np.errstate(divide="ignore", invalid="ignore"):
        for j in range(n):
            iou_cols.append(ref_mu.mask_iou(masks, np.expand_dims(masks[j], axis=0)))
            asy_cols.append(ref_mu.mask_asymmetric_iou(masks, np.expand_dims(masks[j], axis=0)))
    return (np.concatenate(iou_cols, axis=1).astype(np.float16),
            np.concatenate(asy_cols, axis=1).astype(np.float16))


def make_mask_fixture(ref_mu):
    params = synth.proposal_params(40, size=64, seed=11)
    masks = synth.rasterize(params).numpy()
    masks[5] = masks[4]                # identical pair -> IoU exactly 1
    masks[7] = 0                       # empty mask -> 0/0 = NaN row/column
    masks[9] = 1                       # full mask contains everything
    iou, asy = reference_maps(ref_mu, masks)
    o_iou, o_asy = mask_oracle.mask_overlap_maps(masks)
    assert np.array_equal(u16(iou), u16(o_iou)) and np.array_equal(u16(asy), u16(o_asy))
    l_iou, l_asy = mask_oracle.mask_overlap_maps_literal(masks)
    assert np.array_equal(u16(iou), u16(l_iou)) and np.array_equal(u16(asy), u16(l_asy))
    np.savez_compressed(os.path.join(GOLD, "mask_overlap.npz"),
                        masks_bits=np.packbits(masks.reshape(40, -1), axis=1, bitorder="little"),
                        hw=np.int64(64 * 64), iou_u16=u16(iou), asy_u16=u16(asy))
    print("mask_overlap.npz: N=40, NaN entries:", int(np.isnan(iou.astype(np.float32)).sum()))


def scores_for(ref_heads, r, c1, seed, sharp):
    """Plausible head outputs: the reference's own cls_iou_model on random features."""
    torch.manual_seed(seed)
    model = ref_heads.cls_iou_model(64, c1, 3)
    x = torch.randn(r, 64) * sharp
    with torch.no_grad():
        return model, x, model(x)


def heads_cases():
    cases = [
        # name, R, C, mask px, seed, layer thresholds (cls_thr, iou_thr), anti-noise, using_CIM, score sharpness
        ("voc_r300_l0", 300, 20, 64, 21, (0.25, 0.5), True, True, 4.0),
        ("voc_r300_l1", 300, 20, 64, 22, (0.35, 0.6), True, True, 4.0),
        ("voc_r300_l2_nosample", 300, 20, 64, 23, (0.45, 0.7), False, True, 4.0),
        ("coco_r257", 257, 80, 48, 24, (0.25, 0.5), True, True, 6.0),
        ("voc_r120_mist", 120, 20, 48, 25, (0.25, 0.5), True, False, 4.0),
        ("voc_r96_dups", 96, 20, 32, 26, (0.35, 0.6), True, True, 4.0),
        ("voc_r64_none", 64, 20, 32, 27, (0.25, 0.5), True, True, 4.0),
        ("voc_r128_nan", 128, 20, 32, 28, (0.25, 0.5), True, True, 4.0),
        ("voc_r160_edges", 160, 20, 48, 29, (0.35, 0.6), True, True, 4.0),
    ]
    return cases


def poison_maps(name, iou, asy, cls_thr, iou_thr, seed):
    """Adversarial map entries: NaNs, and values sitting exactly on / one ulp around the
    float16 images of the thresholds the heads compare against."""
    rng = np.random.RandomState(seed)
    iou, asy = iou.copy(), asy.copy()
    if "nan" in name:
        for m in (iou, asy):
            m[rng.rand(*m.shape) < 0.02] = np.float16("nan")
    if "edges" in name:
        for m, thrs in ((iou, (cls_thr, iou_thr)), (asy, (0.85,))):
            for t in thrs:
                t16 = np.float16(t)
                for v in (t16, np.nextafter(t16, np.float16(0)), np.nextafter(t16, np.float16(1))):
                    m[rng.rand(*m.shape) < 0.03] = v
    return iou, asy


def make_heads_fixtures(ref_heads, ref_mu, cases=None, save=True):
    cases = cases or heads_cases()
    out = {}
    for name, r, c, px, seed, (cls_thr, iou_thr), anti, using_cim, sharp in cases:
        params = synth.proposal_params(r, size=px, seed=seed)
        masks = synth.rasterize(params).numpy()
        if "dups" in name:                       # duplicated proposals -> tied IoU columns, IoU == 1
            masks[10:20] = masks[0:10]
            masks[40] = 0                        # an empty proposal -> NaN row/col in both maps
        iou, asy = mask_oracle.mask_overlap_maps(masks)
        if "none" in name:                       # nothing can be mined: nobody contains anybody
            asy = np.zeros_like(asy)
        iou, asy = poison_maps(name, iou, asy, cls_thr, iou_thr, seed)
        labels = synth.image_labels(c, 2 if c == 20 else 4, seed)
        _, _, (p_cls, p_det, r_cls, r_iou) = scores_for(ref_heads, r, c + 1, seed, sharp)
        # layer 0 consumes (predict_cls, predict_det); later layers (ref_cls, ref_iou)
        if name.endswith("l1") or name.endswith("l2_nosample"):
            s_cls, s_det = r_cls[0], r_iou[0]
        else:
            s_cls, s_det = p_cls, p_det
        rois = synth.rois_from_params(params)
        with contextlib.redirect_stdout(io.StringIO()):
            layer = ref_heads.CIM_layer(p_seed=0.1, cls_thr=cls_thr, iou_thr=iou_thr,
                                        Anti_noise_sampling=anti)
        t_iou, t_asy = torch.from_numpy(iou), torch.from_numpy(asy)
        np.random.seed(3)
        ref_out = layer(s_cls, s_det, rois, labels, t_iou, t_asy, using_CIM=using_cim)
        np.random.seed(3)
        ora_out = heads_oracle.cim_layer_forward(
            s_cls.numpy(), s_det.numpy(), labels.numpy(), iou, asy, p_seed=0.1, cls_thr=cls_thr,
            iou_thr=iou_thr, con_thr=0.85, anti_noise_sampling=anti, using_cim=using_cim)
        pre = name + "/"
        out[pre + "cls"] = s_cls.numpy()
        out[pre + "det"] = s_det.numpy()
        out[pre + "labels"] = labels.numpy()
        out[pre + "iou_u16"] = u16(iou)
        out[pre + "asy_u16"] = u16(asy)
        out[pre + "params"] = np.array([0.1, cls_thr, iou_thr, 0.85, float(anti), float(using_cim), 3.0])
        if ref_out[0] is None:
            assert ora_out[0] is None, name
            out[pre + "none"] = np.int64(1)
            print(f"{name}: reference returned (None, None, None)")
            continue
        pl, pi, lw = (t.numpy() for t in ref_out)
        assert pi.dtype == np.float16
        assert np.array_equal(pl, ora_out[0]), name
        assert np.array_equal(u16(pi), u16(ora_out[1])), name
        assert np.array_equal(lw.view(np.uint32), ora_out[2].view(np.uint32)), name
        # also pin the mining stage on its own (CIM_label / MIST_label outputs)
        if using_cim:
            _, g_lab, g_w, g_idx, flag = layer.CIM_label(s_cls, s_det, rois[:, 1:], labels, t_iou, t_asy)
            o_cls, o_w, o_flag, _ = heads_oracle.cim_label(s_cls.numpy(), s_det.numpy(), labels.numpy(),
                                                           iou, asy, 0.1, cls_thr, 0.85)
            assert np.array_equal(flag.numpy().reshape(-1), o_flag), name
            out[pre + "asy_iou_flag"] = flag.numpy().reshape(-1)
        else:
            _, g_lab, g_w, g_idx = layer.MIST_label(s_cls * s_det, rois[:, 1:], labels, t_iou)
            o_cls, o_w = heads_oracle.mist_label((s_cls * s_det).numpy(), labels.numpy(), iou, 0.1, cls_thr)
        assert np.array_equal(g_idx.numpy(), o_cls >= 0), name
        assert np.array_equal(g_lab.numpy().argmax(1) - 1, o_cls[o_cls >= 0]), name
        assert np.array_equal(g_w.numpy(), o_w[o_cls >= 0]), name
        if not save:
            continue
        out[pre + "gt_idxs"] = g_idx.numpy()
        out[pre + "gt_class"] = (g_lab.numpy().argmax(1) - 1).astype(np.int64)
        out[pre + "gt_weights"] = g_w.numpy()
        out[pre + "pseudo_labels"] = pl
        out[pre + "pseudo_iou_u16"] = u16(pi)
        out[pre + "loss_weights"] = lw
        print(f"{name}: mined {int(g_idx.sum())} pseudo GT, fg rows {int((pl[:, 1:].sum(1) > 0).sum())}, "
              f"bg rows {int(pl[:, 0].sum())}, ignored {int((pl.sum(1) == 0).sum())}, "
              f"NaN iou labels {int(np.isnan(pi.astype(np.float32)).sum())}")
    if save:
        np.savez_compressed(os.path.join(GOLD, "cim_layer.npz"), **out)


def fuzz(ref_heads, ref_mu, n):
    """Oracle-vs-reference on n random cases that are NOT stored (extra pinning)."""
    rng = np.random.RandomState(99)
    kinds = ["plain", "nan", "edges", "dups"]
    for i in range(n):
        r = int(rng.randint(40, 260))
        c = 20 if rng.rand() < 0.7 else 80
        thr = [(0.25, 0.5), (0.35, 0.6), (0.45, 0.7)][i % 3]
        name = f"fuzz{i}_{kinds[i % 4]}"
        case = (name, r, c, int(rng.choice([32, 48])), 1000 + i, thr, bool(i % 2 == 0),
                bool(i % 5 != 4), float(rng.choice([1.0, 4.0, 8.0])))
        with contextlib.redirect_stdout(io.StringIO()):
            make_heads_fixtures(ref_heads, ref_mu, [case], save=False)
    print(f"fuzz: {n} random cases, oracle == reference bit for bit")


def make_scoring_fixture(ref_heads):
    out = {}
    for name, r, c1, seed in [("voc", 77, 21, 5), ("coco", 33, 81, 6)]:
        model, x, (p_cls, p_det, r_cls, r_iou) = scores_for(ref_heads, r, c1, seed, 3.0)
        names = ["classifier", "detector"] + [f"refine_cls.{k}" for k in range(3)] + \
                [f"refine_iou.{k}" for k in range(3)]
        sd = model.state_dict()
        w = np.stack([sd[n + ".weight"].numpy() for n in names])
        b = np.stack([sd[n + ".bias"].numpy() for n in names])
        ref = np.stack([p_cls.numpy(), p_det.numpy()] + [t.numpy() for t in r_cls] + [t.numpy() for t in r_iou])
        o = heads_oracle.score_heads(x.numpy(), list(w), list(b))
        ora = np.stack([o[0], o[1]] + o[2] + o[3])
        err = np.abs(ora - ref).max() / np.abs(ref).max()
        assert np.allclose(ora, ref, rtol=1e-5, atol=1e-8), err
        out[name + "/x"] = x.numpy()
        out[name + "/w"] = w
        out[name + "/b"] = b
        out[name + "/scores"] = ref
        # backward through the reference's model (torch autograd, what loss.backward() runs, tools/train.py:436)
        xg = x.clone().requires_grad_(True)
        model.zero_grad()
        o2 = model(xg)
        flat = [o2[0], o2[1]] + list(o2[2]) + list(o2[3])
        gen = torch.Generator().manual_seed(seed + 100)
        gs = [torch.randn(t.shape, generator=gen) for t in flat]
        sum((t * g).sum() for t, g in zip(flat, gs)).backward()
        gsd = dict(model.named_parameters())
        gw_ref = np.stack([gsd[n + ".weight"].grad.numpy() for n in names])
        gb_ref = np.stack([gsd[n + ".bias"].grad.numpy() for n in names])
        gx_o, gw_o, gb_o = heads_oracle.score_heads_bwd(x.numpy(), list(w), list(b), [g.numpy() for g in gs])
        for got, want, what in ((gx_o, xg.grad.numpy(), "grad_x"), (np.stack(gw_o), gw_ref, "grad_w"),
                                (np.stack(gb_o), gb_ref, "grad_b")):
            e = np.abs(got - want).max() / max(np.abs(want).max(), 1e-3)
            assert e < 1e-5, (what, e)
            print(f"scoring bwd {name}: oracle vs reference {what} max err / max {e:.2e}")
        out[name + "/grads"] = np.stack([g.numpy() for g in gs])
        out[name + "/grad_x"] = xg.grad.numpy()
        out[name + "/grad_w"] = gw_ref
        out[name + "/grad_b"] = gb_ref
        print(f"scoring {name}: oracle vs reference max rel err {err:.2e}")
    np.savez_compressed(os.path.join(GOLD, "score_heads.npz"), **out)


def make_nms_fixture():
    """Box NMS: the reference's own Cython `nms` (lib/utils/cython_nms.pyx, compiled unmodified apart from two
    numpy type names by oracle/build_ref_nms.py) driven per class as mask_eval_utils.py:63-79 does."""
    sys.path.insert(0, os.path.join(HERE, "_ref"))
    import ref_cython_nms
    from oracle import nms_oracle
    out, rs = {}, np.random.RandomState(11)

    def add(name, boxes, scores, score_thr, nms_thr):
        boxes, scores = boxes.astype(np.float32), scores.astype(np.float32)
        n, c = scores.shape
        keep = np.zeros((c, n), np.uint8)
        for j in range(c):
            inds = np.where(scores[:, j] > score_thr)[0]
            dets = np.hstack((boxes[inds], scores[inds, j][:, None])).astype(np.float32, copy=False)
            if len(dets):
                keep[j, inds[ref_cython_nms.nms(dets, np.float32(nms_thr))]] = 1
        assert np.array_equal(keep, nms_oracle.class_keep(boxes, scores, score_thr, nms_thr)), name
        out[name + "/boxes"], out[name + "/scores"], out[name + "/keep"] = boxes, scores, keep
        out[name + "/params"] = np.array([score_thr, nms_thr], np.float64)
        print(f"box nms {name}: n={n} classes={c} kept={int(keep.sum())}; oracle == reference")

    def tie_free(n, c):                        # every column a permutation: no equal scores inside a class
        return np.stack([(rs.permutation(n) + 1.0) / (n + 1) for _ in range(c)], 1) * rs.uniform(0.2, 1.0, (1, c))

    for name, n, c, seed in [("voc_r300", 300, 20, 21), ("coco_r500", 500, 80, 22)]:
        rois = synth.rois_from_params(synth.proposal_params(n, 512, seed)).numpy()[:, 1:]
        sc = tie_free(n, c)
        sc[rs.rand(n, c) < 0.3] = 0.0                                      # below SCORE_THRESH: not candidates
        add(name, rois, sc, 1e-5, 0.3)
    # integer boxes whose overlap lands exactly on the threshold (ovr >= thresh suppresses): 10x10 boxes
    # shifted by 5 px: inter 50, union 150 -> 1/3; thresholds 1/3 (as float32) and just above it
    grid = np.array([[x, y, x + 9, y + 9] for y in range(0, 40, 5) for x in range(0, 40, 5)], np.float32)
    add("exact_third", grid, tie_free(len(grid), 3), 1e-5, float(np.float32(50.0) / np.float32(150.0)))
    add("above_third", grid, tie_free(len(grid), 3), 1e-5, float(np.nextafter(np.float32(50.0) / np.float32(150.0), np.float32(1))))
    dup = np.repeat(grid[:6], 3, axis=0)                                   # identical boxes: ovr == 1
    add("duplicates", dup, tie_free(len(dup), 2), 1e-5, 0.3)
    deg = grid[:12].copy()
    deg[::3, 2] = deg[::3, 0] - 1                                          # x2 = x1 - 1: zero "+1" width
    add("degenerate", deg, tie_free(len(deg), 2), 1e-5, 0.3)
    add("single", grid[:1], np.array([[0.7, 0.0]]), 1e-5, 0.3)
    add("none", grid[:9], np.zeros((9, 4)), 1e-5, 0.3)
    np.savez_compressed(os.path.join(GOLD, "box_nms.npz"), **out)


def make_pair_fixture(ref_mu):
    """Rectangular ratios: the reference's own mask_iou / mask_asymmetric_iou / mask_inside / mask_outside
    (lib/utils/mask_utils.py through the numpy shim) on two small mask sets, incl. an empty mask (0/0 -> NaN)."""
    rs = np.random.RandomState(31)
    out = {}
    for name, na, nb, h, w in [("n12x3", 12, 3, 24, 20), ("n9x1", 9, 1, 17, 23)]:
        a = (rs.rand(na, h, w) < rs.uniform(0.1, 0.8, (na, 1, 1))).astype(np.uint8)
        b = (rs.rand(nb, h, w) < rs.uniform(0.2, 0.7, (nb, 1, 1))).astype(np.uint8)
        a[0] = 0                                   # empty mask
        a[1] = b[0]                                # identical masks: iou 1
        out[name + "/a"], out[name + "/b"] = a, b
        with np.errstate(divide="ignore", invalid="ignore"), contextlib.redirect_stderr(io.StringIO()):
            for mode, fn in (("iou", ref_mu.mask_iou), ("asymmetric", ref_mu.mask_asymmetric_iou),
                             ("inside", ref_mu.mask_inside), ("outside", ref_mu.mask_outside)):
                ref = fn(a, b)
                ora = mask_oracle.pair_ratio(a, b, mode)
                assert ref.dtype == np.float32 and np.array_equal(ref, ora, equal_nan=True), (name, mode)
                out[f"{name}/{mode}"] = ref
        print(f"mask pair {name}: oracle == reference for iou / asymmetric / inside / outside")
    np.savez_compressed(os.path.join(GOLD, "mask_pair.npz"), **out)


def synth_cluster_mat(R, c1, n_fg, seed, with_bg=True):
    """A cluster matrix shaped like tools/pre/AGPL_label_assign.py:60-96 writes it: n_fg foreground clusters, each in
    one class column for a random subset of rows (later clusters overwrite earlier ones on shared rows), then the
    background cluster's id in column 0 of some of the untouched rows."""
    rs = np.random.RandomState(seed)
    mat = np.zeros((R, c1), np.float32)
    k = 1
    for _ in range(n_fg):
        rows = rs.rand(R) < rs.uniform(0.03, 0.15)
        mat[rows, :] = 0
        mat[rows, 1 + rs.randint(c1 - 1)] = k
        k += 1
    if with_bg:
        free = (mat.sum(1) == 0) & (rs.rand(R) < 0.5)
        mat[free, 0] = k
    return mat


def make_pcl_fixture(ref_heads):
    """PCL_loss: the reference's own heads.PCL_loss (imported unmodified; its hard `.cuda()` at heads.py:11 is
    neutralised by making Tensor.cuda the identity for the duration of the call) with autograd."""
    from oracle import loss_oracle
    out = {}
    real_cuda = torch.Tensor.cuda
    for name, R, c1, n_fg, seed, bg in [("voc_r300", 300, 21, 5, 41, True), ("coco_r257", 257, 81, 9, 42, True),
                                        ("voc_nobg", 120, 21, 3, 43, False), ("voc_empty", 64, 21, 0, 44, False)]:
        torch.manual_seed(seed)
        model = ref_heads.cls_iou_model(64, c1, 3)
        p = model(torch.randn(R, 64) * 3)[0].detach().clone()
        p[0, 0], p[1, 1] = 0.0, 1.0                                          # outside the clamp range
        mat = synth_cluster_mat(R, c1, n_fg, seed, bg)
        leaf = p.clone().requires_grad_(True)
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            loss = ref_heads.PCL_loss(leaf, torch.from_numpy(mat), torch.zeros(1, c1 - 1))
        finally:
            torch.Tensor.cuda = real_cuda
        if loss.requires_grad:                       # no cluster at all: the loss is the constant 0
            loss.backward()
        grad = leaf.grad.numpy() if leaf.grad is not None else np.zeros_like(p.numpy())
        o_loss, o_grad = loss_oracle.pcl_losses(p.numpy(), mat[None])
        e_l = abs(o_loss[0] - float(loss)) / max(abs(float(loss)), 1e-3)
        e_g = np.abs(o_grad - grad).max() / max(np.abs(grad).max(), 1e-3)
        assert e_l < 1e-5 and e_g < 1e-5, (name, e_l, e_g)
        print(f"pcl {name}: loss {float(loss):.5f}, oracle vs reference loss err {e_l:.1e}, grad err {e_g:.1e}")
        out[name + "/predict_cls"], out[name + "/mat"] = p.numpy(), mat
        out[name + "/loss"], out[name + "/grad"] = np.float32(float(loss)), grad
    np.savez_compressed(os.path.join(GOLD, "pcl_loss.npz"), **out)


def make_loss_fixture(ref_heads):
    """Loss block: the reference's own heads.cls_iou_loss / heads.mil_bag_loss (imported unmodified) with
    autograd, wired as model_builder.py:170-202, on the inputs + CIM_layer outputs of stored cim_layer cases."""
    from oracle import loss_oracle
    cim = np.load(os.path.join(GOLD, "cim_layer.npz"))
    out = {}
    for name in ["voc_r300_l0", "coco_r257", "voc_r128_nan", "voc_r64_none"]:
        R, c1 = cim[f"{name}/cls"].shape
        k = 3
        torch.manual_seed(len(name))
        model = ref_heads.cls_iou_model(64, c1, k)
        x = torch.randn(R, 64) * 3
        p_cls, p_det, r_cls, r_iou = model(x)
        flat = [p_cls, p_det] + list(r_cls) + list(r_iou)
        scores = torch.stack([t.detach() for t in flat]).clone()
        scores[2, 0, 0], scores[2, 1, 1] = 0.0, 1.0                       # values outside the clamp range
        leaf = scores.clone().requires_grad_(True)
        labels = torch.from_numpy(cim[f"{name}/labels"]).float().reshape(1, -1)
        none = f"{name}/none" in cim.files
        # the same pseudo labels feed all three layers (what matters here is the loss arithmetic); layer 1 is
        # marked invalid to cover the `continue` at model_builder.py:189-190
        if none:
            pl = np.zeros((R, c1), np.float32); pi = np.zeros(R, np.float16); lw = np.zeros(R, np.float32)
        else:
            pl, lw = cim[f"{name}/pseudo_labels"], cim[f"{name}/loss_weights"]
            pi = cim[f"{name}/pseudo_iou_u16"].view(np.float16)
        valid = np.array([[0 if none else 1], [0], [0 if none else 1]], np.uint8)
        losses = torch.zeros(1, k + 1, 3)
        total = torch.zeros(())
        for l in range(k):
            if not valid[l, 0]:
                continue
            lmda = 3 if l == 0 else 1
            c, i, g = ref_heads.cls_iou_loss(leaf[2 + l], leaf[2 + k + l], torch.from_numpy(pl),
                                             torch.from_numpy(pi), lmda * torch.from_numpy(lw), labels)
            losses[0, l] = torch.stack([c.detach(), i.detach(), g.detach()])
            total = total + c + 3 * i + g
        mil = ref_heads.mil_bag_loss(leaf[0], leaf[1], labels)
        losses[0, k, 2] = mil.detach()
        (total + mil).backward()
        pl3 = np.stack([pl] * k)[:, None]
        pi3 = np.stack([pi] * k)[:, None]
        lw3 = np.stack([lw] * k)[:, None]
        o_loss, o_grad = loss_oracle.head_losses(scores.numpy(), pl3, pi3, lw3, valid, labels.numpy(), k)
        fin = np.isfinite(leaf.grad.numpy())
        e_l = np.nanmax(np.abs(o_loss - losses.numpy())) / max(np.nanmax(np.abs(losses.numpy())), 1e-3)
        e_g = np.abs(o_grad - leaf.grad.numpy())[fin].max() / max(np.abs(leaf.grad.numpy()[fin]).max(), 1e-3)
        assert e_l < 1e-5 and e_g < 1e-5, (name, e_l, e_g)
        assert np.array_equal(np.isnan(o_grad), np.isnan(leaf.grad.numpy()))
        print(f"losses {name}: oracle vs reference loss err {e_l:.1e}, grad err {e_g:.1e}, "
              f"losses {losses.numpy()[0].sum(0)}")
        for key, v in dict(scores=scores.numpy(), labels=labels.numpy(), pseudo_labels=pl3, pseudo_iou_u16=pi3.view(np.uint16),
                           loss_weights=lw3, valid=valid, losses=losses.numpy(), grad=leaf.grad.numpy()).items():
            out[f"{name}/{key}"] = v
    np.savez_compressed(os.path.join(GOLD, "head_losses.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    if "--only-nms" in sys.argv:
        make_nms_fixture()
        return
    if "--only-pair" in sys.argv:
        make_pair_fixture(load_reference()[1])
        return
    if "--only-pcl" in sys.argv:
        make_pcl_fixture(load_reference()[0])
        return
    if "--only-losses" in sys.argv:
        make_loss_fixture(load_reference()[0])
        return
    torch.set_num_threads(4)
    ref_heads, ref_mu = load_reference()
    make_mask_fixture(ref_mu)
    make_heads_fixtures(ref_heads, ref_mu)
    make_scoring_fixture(ref_heads)
    make_nms_fixture()
    make_loss_fixture(ref_heads)
    make_pcl_fixture(ref_heads)
    make_pair_fixture(ref_mu)
    if "--fuzz" in sys.argv:
        fuzz(ref_heads, ref_mu, int(sys.argv[sys.argv.index("--fuzz") + 1]))
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
