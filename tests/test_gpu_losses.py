"""cim_head_losses (loss block forward + backward in one launch) against the fixtures of the reference's own
loss functions + autograd, and against the oracle at batch sizes.  Tolerance 1e-5 relative (fp32)."""
import os

import numpy as np
import pytest
import torch

from cim_b200 import heads
from oracle import loss_oracle
from conftest import GOLDEN, cim_case_names

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NPZ = np.load(os.path.join(GOLDEN, "head_losses.npz"))
NAMES = cim_case_names(NPZ)


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def run(scores, pl, pi, lw, valid, labels, k=3, upstream=None):
    s = cuda(scores).requires_grad_(True)
    assigned = dict(pseudo_labels=cuda(pl), pseudo_iou_labels=cuda(pi), loss_weights=cuda(lw), valid=cuda(valid))
    out = heads.head_losses(s, assigned, cuda(labels), k)
    up = torch.ones_like(out["total"]) if upstream is None else cuda(upstream)
    out["total"].backward(up)
    return out, s.grad.cpu().numpy()


def close(got, want, tol=1e-5):
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    fin = np.isfinite(want)
    if fin.any():
        assert np.abs(got - want)[fin].max() <= tol * max(np.abs(want[fin]).max(), 1e-3)


@pytest.mark.parametrize("name", NAMES)
def test_reference_fixtures(name):
    g = lambda k: NPZ[f"{name}/{k}"]
    out, grad = run(g("scores"), g("pseudo_labels"), g("pseudo_iou_u16").view(np.float16), g("loss_weights"),
                    g("valid"), g("labels"))
    close(out["losses"].cpu().numpy(), g("losses"))
    close(grad, g("grad"))
    want = g("losses")
    close(out["total"].detach().cpu().numpy(), want[:, :, 0].sum(1) + 3 * want[:, :, 1].sum(1) + want[:, :, 2].sum(1))


@pytest.mark.parametrize("n_img,R,C", [(3, 500, 20), (2, 2000, 80)])
def test_batched_matches_oracle(n_img, R, C):
    """Several images, full proposal counts, random pseudo labels (fg one-hot / bg / ignored rows), an invalid
    (layer, image) pair and a non-trivial upstream gradient per image."""
    rs = np.random.RandomState(R + C)
    k, c1 = 3, C + 1
    z = rs.randn(2 + 2 * k, n_img * R, c1).astype(np.float32) * 2
    scores = np.empty_like(z)
    e = np.exp(z - z.max(-1, keepdims=True))
    scores[[0, 2, 3, 4]] = (e / e.sum(-1, keepdims=True))[[0, 2, 3, 4]]
    d = z[1].reshape(n_img, R, c1)
    ed = np.exp(d - d.max(1, keepdims=True))
    scores[1] = (ed / ed.sum(1, keepdims=True)).reshape(n_img * R, c1)
    scores[5:] = 1 / (1 + np.exp(-z[5:]))
    labels = np.zeros((n_img, C), np.float32)
    pl = np.zeros((k, n_img, R, c1), np.float32)
    for b in range(n_img):
        present = rs.choice(C, 3, replace=False)
        labels[b, present] = 1
        for l in range(k):
            kind = rs.rand(R)
            fg = kind < 0.3
            pl[l, b, fg, 1 + rs.choice(present, fg.sum())] = 1
            pl[l, b, (kind >= 0.3) & (kind < 0.8), 0] = 1
    pi = (rs.rand(k, n_img, R) > 0.5).astype(np.float16)
    lw = rs.rand(k, n_img, R).astype(np.float32)
    valid = np.ones((k, n_img), np.uint8)
    valid[1, 0] = 0
    up = rs.rand(n_img).astype(np.float32) + 0.5
    out, grad = run(scores, pl, pi, lw, valid, labels, k, upstream=up)
    o_loss, o_grad = loss_oracle.head_losses(scores, pl, pi, lw, valid, labels, k)
    o_grad = (o_grad.reshape(2 + 2 * k, n_img, R, c1) * up[None, :, None, None]).reshape(grad.shape)
    close(out["losses"].cpu().numpy(), o_loss.astype(np.float32))
    close(grad, o_grad.astype(np.float32))


def test_losses_feed_the_scoring_backward():
    """End of the chain: scores -> cim_head_losses -> grad_scores -> cim_score_heads_bwd -> parameter gradients,
    against torch autograd through the oracle restatement of the same chain (fp32 reference of a fp kernel)."""
    torch.manual_seed(3)
    R, D, C, k = 256, 128, 20, 3
    m = heads.cls_iou_model(D, C + 1, k).to(DEV)
    x = torch.randn(R, D, device=DEV)
    rs = np.random.RandomState(0)
    pl = np.zeros((k, 1, R, C + 1), np.float32)
    labels = np.zeros((1, C), np.float32)
    labels[0, [2, 7]] = 1
    for l in range(k):
        kind = rs.rand(R)
        pl[l, 0, kind < 0.3, 1 + rs.choice([2, 7], (kind < 0.3).sum())] = 1
        pl[l, 0, (kind >= 0.3) & (kind < 0.8), 0] = 1
    pi = (rs.rand(k, 1, R) > 0.5).astype(np.float16)
    lw = rs.rand(k, 1, R).astype(np.float32)
    valid = np.ones((k, 1), np.uint8)
    assigned = dict(pseudo_labels=cuda(pl), pseudo_iou_labels=cuda(pi), loss_weights=cuda(lw), valid=cuda(valid))
    s = m.forward_batched(x, 1)
    heads.head_losses(s, assigned, cuda(labels), k)["total"].sum().backward()
    got = {n: p.grad.cpu().numpy() for n, p in m.named_parameters()}
    # reference chain on the CPU in float64
    m64 = heads.cls_iou_model(D, C + 1, k).double()
    m64.load_state_dict({n: v.double().cpu() for n, v in m.state_dict().items()})
    x64 = x.double().cpu()
    layers = [m64.classifier, m64.detector, *m64.refine_cls, *m64.refine_iou]
    sc = [torch.softmax(layers[0](x64), -1), torch.softmax(layers[1](x64), 0)] + \
         [torch.softmax(l(x64), -1) for l in layers[2:2 + k]] + [torch.sigmoid(l(x64)) for l in layers[2 + k:]]
    total = loss_oracle.mil_bag_loss(sc[0], sc[1], torch.from_numpy(labels[0]))
    for l in range(k):
        c, i, g = loss_oracle.cls_iou_loss(sc[2 + l], sc[2 + k + l], torch.from_numpy(pl[l, 0]).double(),
                                           torch.from_numpy(pi[l, 0].astype(np.float32)).double(),
                                           (3.0 if l == 0 else 1.0) * torch.from_numpy(lw[l, 0]).double(),
                                           torch.from_numpy(labels[0]))
        total = total + c + 3 * i + g
    total.backward()
    for n, p in m64.named_parameters():
        want = p.grad.numpy()
        assert np.abs(got[n] - want).max() <= 2e-5 * max(np.abs(want).max(), 1e-2), n


# ----------------------------------------------------------------------------------- PCL_loss
PCL = np.load(os.path.join(GOLDEN, "pcl_loss.npz"))


@pytest.mark.parametrize("name", cim_case_names(PCL))
def test_pcl_loss_reference_fixtures(name):
    p = cuda(PCL[f"{name}/predict_cls"]).requires_grad_(True)
    loss = heads.PCL_loss(p, cuda(PCL[f"{name}/mat"]), None)
    assert loss.dim() == 0
    loss.backward()
    want_l, want_g = float(PCL[f"{name}/loss"]), PCL[f"{name}/grad"]
    assert abs(float(loss.detach()) - want_l) <= 1e-5 * max(abs(want_l), 1e-3)
    close(p.grad.cpu().numpy(), want_g)


def test_pcl_loss_batched_accumulate_and_errors():
    """3 images at cfg2 size in one launch against the oracle; accumulate mode adds on top of an existing
    gradient; ids the reference would choke on give NaN."""
    import ctypes as C
    from cim_b200 import _lib
    n_img, R, c1 = 3, 2000, 21
    rs = np.random.RandomState(8)
    z = rs.randn(n_img * R, c1).astype(np.float32) * 2
    e = np.exp(z - z.max(-1, keepdims=True))
    p = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    mat = np.zeros((n_img, R, c1), np.float32)
    for b in range(n_img):
        k = 1
        for _ in range(6 + b):
            rows = rs.rand(R) < 0.05
            mat[b, rows, :] = 0
            mat[b, rows, 1 + rs.randint(c1 - 1)] = k
            k += 1
        free = (mat[b].sum(1) == 0) & (rs.rand(R) < 0.4)
        mat[b, free, 0] = k
    mat[1, 5, 3], mat[1, 5, 7] = 2, 3                       # a row in two clusters
    up = np.array([0.5, 2.0, 1.0], np.float32)
    pt = cuda(p).requires_grad_(True)
    loss = heads.PCL_loss(pt, cuda(mat), None, n_img=n_img)
    (loss * cuda(up)).sum().backward()
    o_loss, o_grad = loss_oracle.pcl_losses(p, mat)
    o_grad = (o_grad.reshape(n_img, R, c1) * up[:, None, None]).reshape(n_img * R, c1)
    close(loss.detach().cpu().numpy(), o_loss.astype(np.float32))
    close(pt.grad.cpu().numpy(), o_grad.astype(np.float32))
    # accumulate on top of ones
    L = _lib.lib()
    g = torch.ones(n_img * R, c1, device=DEV)
    lo = torch.empty(n_img, device=DEV)
    dp, dm = cuda(p), cuda(mat)                              # keep the device inputs alive across the raw call
    _lib.check(L.cim_pcl_loss(_lib.ptr(dp), _lib.ptr(dm), _lib.ptr(lo), _lib.ptr(g), n_img, R, c1, 128, 1.0,
                              1, _lib.stream_ptr(torch.device(DEV))), "cim_pcl_loss")
    torch.cuda.synchronize()
    o_loss1, o_grad1 = loss_oracle.pcl_losses(p, mat)
    close(g.cpu().numpy() - 1.0, o_grad1.astype(np.float32), tol=1e-4)      # 1 + x - 1 costs a few ulps of 1
    # two background ids (heads.py:20 asserts) and a non-integer id -> NaN
    bad = mat.copy()
    bad[0, 0, 0], bad[0, 1, 0] = 40, 41
    bad[2, 0, 4] = 1.5
    lo2 = heads.PCL_loss(cuda(p), cuda(bad), None, n_img=n_img).cpu().numpy()
    assert np.isnan(lo2[0]) and np.isnan(lo2[2]) and abs(lo2[1] - o_loss[1]) < 1e-4
