"""Seeded synthetic inputs for the CIM head path.

COB-like mask proposals: 64 random seed ellipses per image, every proposal is one seed
ellipse rescaled about its centre, so proposals nest and containment > 0.85 really occurs
(i.i.d. random masks never nest).  All geometry is integer, so a CPU and a CUDA
rasterisation of the same parameters are identical bit for bit.

ROIs follow tools/pre/generate_7_7_voc.py:36-41 of the reference: the tight box of the mask
as (xmin, ymin, xmax+1, ymax+1), prefixed with the batch index (lib/roi_data/minibatch.py:45-48).
"""
from dataclasses import dataclass

import numpy as np
import torch

N_SEED_ELLIPSES = 64


@dataclass
class ProposalParams:
    cx: np.ndarray   # [R] int64 centre x
    cy: np.ndarray   # [R] int64 centre y
    ax: np.ndarray   # [R] int64 semi-axis x (>=1)
    ay: np.ndarray   # [R] int64 semi-axis y (>=1)
    size: int        # image side in pixels


def proposal_params(num_props, size=512, seed=1234):
    """Ellipse parameters of `num_props` proposals on a size x size image."""
    rng = np.random.RandomState(seed)
    unit = size / 512.0
    scx = rng.randint(0, size, N_SEED_ELLIPSES)
    scy = rng.randint(0, size, N_SEED_ELLIPSES)
    sax = rng.uniform(16 * unit, 160 * unit, N_SEED_ELLIPSES)
    say = rng.uniform(16 * unit, 160 * unit, N_SEED_ELLIPSES)
    which = np.arange(num_props) % N_SEED_ELLIPSES
    scale = rng.uniform(0.4, 1.6, num_props)
    ax = np.maximum(1, np.rint(sax[which] * scale)).astype(np.int64)
    ay = np.maximum(1, np.rint(say[which] * scale)).astype(np.int64)
    return ProposalParams(scx[which].astype(np.int64), scy[which].astype(np.int64), ax, ay, size)


def rasterize(params, device="cpu", out_size=None, chunk=256):
    """uint8 masks [R, S, S] (0/1).  out_size < size samples every (size/out_size)-th pixel
    (nearest down-sampling, cfg4 of BASELINE.json uses 1/4 resolution)."""
    size = params.size
    out_size = out_size or size
    step = size // out_size
    assert step * out_size == size
    dev = torch.device(device)
    coords = torch.arange(0, size, step, device=dev, dtype=torch.int64)
    r = len(params.cx)
    out = torch.empty((r, out_size, out_size), dtype=torch.uint8, device=dev)
    cx = torch.as_tensor(params.cx, device=dev)
    cy = torch.as_tensor(params.cy, device=dev)
    ax = torch.as_tensor(params.ax, device=dev)
    ay = torch.as_tensor(params.ay, device=dev)
    for s in range(0, r, chunk):
        e = min(r, s + chunk)
        dx2 = (coords[None, :] - cx[s:e, None]) ** 2          # [n, S]
        dy2 = (coords[None, :] - cy[s:e, None]) ** 2
        a2 = (ax[s:e] ** 2)[:, None, None]
        b2 = (ay[s:e] ** 2)[:, None, None]
        inside = dx2[:, None, :] * b2 + dy2[:, :, None] * a2 <= a2 * b2
        out[s:e] = inside.to(torch.uint8)
    return out


def rois_from_params(params, batch_index=0, im_scale=1.0):
    """float32 [R,5] = (batch_idx, x1, y1, x2, y2): tight mask box, exclusive max edge."""
    s = params.size
    x1 = np.maximum(params.cx - params.ax, 0)
    y1 = np.maximum(params.cy - params.ay, 0)
    x2 = np.minimum(params.cx + params.ax, s - 1) + 1
    y2 = np.minimum(params.cy + params.ay, s - 1) + 1
    rois = np.stack([np.full_like(x1, batch_index), x1, y1, x2, y2], axis=1).astype(np.float32)
    rois[:, 1:] *= im_scale
    return torch.from_numpy(rois)


def image_labels(num_classes=20, num_present=2, seed=1234):
    """float32 [1, C] 0/1 image-level labels with `num_present` classes set."""
    rng = np.random.RandomState(seed + 7919)
    lab = np.zeros((1, num_classes), dtype=np.float32)
    lab[0, rng.choice(num_classes, num_present, replace=False)] = 1
    return torch.from_numpy(lab)


def cluster_mat(num_props, num_classes=20, present=(0, 1), n_fg=6, seed=1234):
    """float32 [R, C+1] proposal-cluster matrix shaped like tools/pre/AGPL_label_assign.py:60-96 writes it (the `mat`
    blob PCL_loss consumes): n_fg foreground clusters, each in the column of a present class for ~5 % of the
    proposals (a later cluster takes over shared rows), then the background cluster's id in column 0 of about half
    of the untouched rows."""
    rng = np.random.RandomState(seed + 104729)
    mat = np.zeros((num_props, num_classes + 1), dtype=np.float32)
    k = 1
    for _ in range(n_fg):
        rows = rng.rand(num_props) < 0.05
        mat[rows, :] = 0
        mat[rows, 1 + int(present[rng.randint(len(present))])] = k
        k += 1
    free = (mat.sum(1) == 0) & (rng.rand(num_props) < 0.5)
    mat[free, 0] = k
    return torch.from_numpy(mat)


# Feature-map shapes of a 512x512 image for the reference backbones
# (lib/modeling/resnet50.py:42-44, vgg16.py:80-81, HRNet.py:316-318).
BACKBONES = {
    "resnet50": dict(channels=1024, stride=16),
    "vgg16": dict(channels=512, stride=8),
    "hrnet48": dict(channels=2048, stride=32),
}


def feature_shape(backbone, size=512):
    b = BACKBONES[backbone]
    return b["channels"], size // b["stride"], size // b["stride"], 1.0 / b["stride"]
