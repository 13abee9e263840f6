#!/usr/bin/env python
"""tests/golden/model_forward.npz: one training forward of the reference's OWN model code, executed unmodified.

Runs /root/reference/lib/modeling/model_builder.py Generalized_RCNN.forward (:121-207) in training mode on the CPU,
with the configuration of configs/resnet50_voc.yaml -- Conv_Body = resnet50.torch_resnet50, Box_Head =
resnet50.MaskFuse (resnet50.py:94-138) calling roi_feature_transform (model_builder.py:215-233) -> ops.RoIAlign,
heads.cls_iou_model, 3 x heads.CIM_layer, heads.cls_iou_loss / mil_bag_loss / PCL_loss -- and records, with hooks, every
tensor that crosses the boundary this repo replaces.  tests/test_gpu_model_forward.py then replays the same inputs
through cim_b200's operators on the B200 (the reference tree does not travel to the GPU box) and compares stage by stage.

What had to be supplied for the reference to import and run here (none of it touches the files under /root/reference):
  * `torch._six` (lib/nn/parallel/scatter_gather.py:8 imports string_classes from it; gone from torch 2.x): a 2-line
    stub module; `yaml.load` given its pre-6.0 default Loader (core/config.py:678);
  * `mmcv.ops` (lib/ops/__init__.py:6; mmcv is absent): RoIAlign / RoIPool stand-ins that call torchvision's CPU
    roi_align(aligned=True) / roi_pool -- the same Detectron2-derived algorithm as mmcv 1.x, see SURVEY.md section 8c;
  * Tensor.cuda made the identity (heads.PCL_loss does a hard `.cuda()`, heads.py:11);
  * LOAD_IMAGENET_PRETRAINED_WEIGHTS off (no network) and cfg.iou_dir / asy_iou_dir pointed at a temp directory holding
    the two fp16 pickles, written with the reference's own mask_utils (create_cob_iou.py:43-49).
Weights: the sub-modules this repo replaces or feeds are re-initialised from fixed seeds AFTER construction
(cls_iou_model: seed 100, N(0, 0.001); Box_Head: seed 101, He init), so the GPU test can rebuild identical weights from the seed
instead of storing 230 M parameters.

    python oracle/make_model_golden.py
"""
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("CIM_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

SEED_HEADS, SEED_BOXHEAD, SEED_RNG = 100, 101, 3
R, SIZE, C = 40, 128, 20
SEL = [0, 1, 31, 32, 500, 777, 1022, 1023]          # sampled channels of the 1024-channel RoI tensors


def reinit_heads(model):
    """N(0, 0.001) weights (logits of order 1 on the O(20) features a random-init body produces) / N(0, 0.1) biases from SEED_HEADS in named_parameters() order (classifier, detector,
    refine_cls.k, refine_iou.k -- the same names in heads.py:168-185 and cim_b200/heads.py)."""
    g = torch.Generator().manual_seed(SEED_HEADS)
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.001 if name.endswith("weight") else 0.1))


def install_stubs():
    six = types.ModuleType("torch._six")
    six.string_classes, six.int_classes = (str, bytes), int
    sys.modules["torch._six"] = six
    import yaml                                     # core/config.py:678 calls yaml.load(f) the PyYAML < 6 way
    _load = yaml.load
    yaml.load = lambda stream, Loader=yaml.SafeLoader: _load(stream, Loader=Loader)
    from torchvision.ops import roi_align as tv_roi_align, roi_pool as tv_roi_pool

    class RoIAlign(nn.Module):
        def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode="avg", aligned=True):
            super().__init__()
            self.a = (output_size, spatial_scale, sampling_ratio, aligned)

        def forward(self, x, rois):
            o, s, sr, al = self.a
            return tv_roi_align(x, rois, o, s, sr, al)

    class RoIPool(nn.Module):
        def __init__(self, output_size, spatial_scale=1.0):
            super().__init__()
            self.a = (output_size, spatial_scale)

        def forward(self, x, rois):
            return tv_roi_pool(x, rois, *self.a)

    mmcv, ops = types.ModuleType("mmcv"), types.ModuleType("mmcv.ops")
    ops.RoIAlign, ops.RoIPool = RoIAlign, RoIPool
    ops.roi_align = ops.roi_pool = ops.nms = ops.soft_nms = None
    mmcv.ops = ops
    sys.modules.update({"mmcv": mmcv, "mmcv.ops": ops})


def main():
    import make_golden
    from cim_b200 import synth
    _, ref_mu = make_golden.load_reference()            # also puts /root/reference/lib first on sys.path
    install_stubs()
    from core.config import cfg, cfg_from_file, assert_and_infer_cfg
    import modeling.model_builder as mb
    import modeling.heads as ref_heads

    tmp = tempfile.mkdtemp()
    cfg_from_file(os.path.join(REF, "configs", "resnet50_voc.yaml"))
    cfg.MODEL.LOAD_IMAGENET_PRETRAINED_WEIGHTS = False
    cfg.MODEL.NUM_CLASSES = C
    cfg.iou_dir, cfg.asy_iou_dir = os.path.join(tmp, "iou"), os.path.join(tmp, "asy")
    os.makedirs(cfg.iou_dir), os.makedirs(cfg.asy_iou_dir)
    assert cfg.FAST_RCNN.ROI_XFORM_METHOD == "RoIAlign" and cfg.FAST_RCNN.ROI_XFORM_RESOLUTION == 7
    assert cfg.REFINE_TIMES == 3 and cfg.Anti_noise_sampling

    torch.manual_seed(0)
    torch.set_num_threads(8)
    model = mb.Generalized_RCNN()
    reinit_heads(model.cls_iou_model)
    torch.manual_seed(SEED_BOXHEAD)
    with torch.no_grad():
        for p in model.Box_Head.parameters():             # He init: activations stay O(1) through conv + 2 x fc
            p.normal_(0, (2.0 / p[0].numel()) ** 0.5 if p.dim() > 1 else 0.01)
    model.train()

    params = synth.proposal_params(R, SIZE, 77)
    rois = synth.rois_from_params(params)                                  # [R, 5]
    masks_full = synth.rasterize(params).numpy()
    iou, asy = make_golden.reference_maps(ref_mu, masks_full)              # the reference's own mask_utils, fp16
    pickle.dump(iou, open(os.path.join(cfg.iou_dir, "img0.pkl"), "wb"))
    pickle.dump(asy, open(os.path.join(cfg.asy_iou_dir, "img0.pkl"), "wb"))
    g = torch.Generator().manual_seed(5)
    data = torch.randn(1, 3, SIZE, SIZE, generator=g)
    masks7 = (torch.rand(R, 7, 7, generator=g) > 0.4).float()
    labels = synth.image_labels(C, 2, 77)
    mat = synth.cluster_mat(R, C, np.nonzero(labels[0].numpy())[0], 4, 77)

    rec = {}
    xform = model.Box_Head.roi_xform
    def rec_xform(*a, **k):
        out = xform(*a, **k)
        rec["box_x"] = out.detach()
        return out
    model.Box_Head.roi_xform = rec_xform
    model.Box_Head.mask_branch.register_forward_pre_hook(lambda m, inp: rec.__setitem__("box_mask_cat", inp[0].detach()))
    model.Box_Head.register_forward_hook(lambda m, inp, out: rec.__setitem__("seg_x", out.detach()))
    model.cls_iou_model.register_forward_hook(lambda m, inp, out: rec.__setitem__("scores", out))
    for i, layer in enumerate(model.CIM_layer_list):
        layer.register_forward_hook(lambda m, inp, out, i=i: rec.__setitem__(f"cim{i}", out))

    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    np.random.seed(SEED_RNG)
    try:
        ret = model(data, rois[None], masks7[None], labels[None], None, mat[None], path="data/VOC/img0.jpg",
                    index=torch.arange(R)[None])
    finally:
        torch.Tensor.cuda = real_cuda

    p_cls, p_det, r_cls, r_iou = rec["scores"]
    out = {
        "blob_conv": ret["blob_conv"].detach().numpy(), "rois": rois.numpy(), "masks7": masks7.numpy(),
        "labels": labels.numpy(), "mat": mat.numpy(), "iou": iou, "asy": asy,
        "sel": np.array(SEL), "box_x_sel": rec["box_x"][:, SEL].numpy(),
        "box_mask_cat_sel": rec["box_mask_cat"][:, SEL + [1024 + c for c in SEL]].numpy(),
        "seg_x": rec["seg_x"].numpy(),
        "scores": np.stack([t.detach().numpy() for t in [p_cls, p_det, *r_cls, *r_iou]]),
        "losses": np.array([float(ret["losses"][k]) for k in ("cls_loss", "iou_loss", "bag_loss", "pcl_loss")], np.float32),
        "seeds": np.array([SEED_HEADS, SEED_BOXHEAD, SEED_RNG]),
    }
    for i in range(3):
        pl, pi, lw = rec[f"cim{i}"]
        out[f"cim{i}/valid"] = np.array(pl is not None)
        if pl is not None:
            out[f"cim{i}/pseudo_labels"], out[f"cim{i}/pseudo_iou"], out[f"cim{i}/loss_weights"] = \
                pl.numpy(), pi.numpy(), lw.numpy()
    np.savez_compressed(os.path.join(GOLD, "model_forward.npz"), **out)
    print("losses (cls, 3*iou, bag, pcl):", out["losses"], " mined layers:", [bool(out[f'cim{i}/valid']) for i in range(3)])
    print("written", os.path.join(GOLD, "model_forward.npz"), os.path.getsize(os.path.join(GOLD, "model_forward.npz")), "bytes")


if __name__ == "__main__":
    main()
