from cim_b200.ops import RoIAlign, RoIPool, nms, roi_align, roi_pool, soft_nms

__all__ = ["RoIPool", "RoIAlign", "roi_pool", "roi_align", "nms", "soft_nms"]
