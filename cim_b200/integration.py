"""Drop-in helpers for running the unmodified reference tree on top of cim_b200."""
import os
import sys


def shim_path():
    """Directory to put on sys.path / PYTHONPATH so that `import mmcv.ops` resolves to cim_b200.ops."""
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def install_mmcv_shim():
    p = shim_path()
    if p not in sys.path:
        sys.path.insert(0, p)


def patch_reference_heads(ref_heads_module, pcl_loss=True):
    """Replace cls_iou_model and CIM_layer inside the reference's `modeling.heads` module object
    (lib/modeling/heads.py:168,222) and, with pcl_loss, PCL_loss (heads.py:10-41: same signature, one kernel
    instead of a Python loop over clusters with two host syncs each); cls_iou_loss / mil_bag_loss stay the
    reference's (their fused, batched counterpart is cim_b200.heads.head_losses), so
    lib/modeling/model_builder.py:86-94,143,176-204 runs unchanged."""
    from . import heads
    ref_heads_module.cls_iou_model = heads.cls_iou_model
    ref_heads_module.CIM_layer = heads.CIM_layer
    if pcl_loss:
        ref_heads_module.PCL_loss = heads.PCL_loss
    return ref_heads_module
