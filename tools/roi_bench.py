"""RoIAlign fwd / bwd micro-benchmark at the cfg2 shape (8 images x 2000 proposals, R-50 features),
through the C ABI.  Prints device times (CUDA events, median of N) and the algorithmic HBM rate.
Optional check against torchvision's CUDA roi_align on a subset (tool only; tests use the oracle).

    python tools/roi_bench.py [--images 8] [--props 2000] [--backbone resnet50] [--iters 20] [--check]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cim_b200 import _lib, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--props", type=int, default=2000)
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--pool", action="store_true", help="also time RoIPool forward / backward (both bin conventions)")
    ap.add_argument("--size", type=int, default=512, help="image side in pixels (map side = size / stride): the "
                    "reference trains at scales 480 .. 1200 (configs/resnet50_voc.yaml:34)")
    ap.add_argument("--debug-flags", type=int, default=0, help="cim_set_debug_flags (8 = global-pairs path instead "
                    "of window tiles, 1 = shared-memory gradient tile)")
    ap.add_argument("--profile-fused-bwd", action="store_true", help="run only the fused backward 3 times (for ncu)")
    ap.add_argument("--fused", action="store_true", help="also time the fused MaskFuse prologue and the unfused "
                    "torch expression it replaces (box * mask, concat)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    Cf, H, W, scale = synth.feature_shape(a.backbone, a.size)
    B, R = a.images, a.props
    _lib.lib().cim_set_debug_flags(a.debug_flags)
    rois = torch.cat([synth.rois_from_params(synth.proposal_params(R, a.size, 1234 + b), b) for b in range(B)]).to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    feat = torch.randn(B, Cf, H, W, device=dev, generator=g)
    K = rois.size(0)
    out = torch.empty(K, Cf, 7, 7, device=dev)
    gout = torch.randn(K, Cf, 7, 7, device=dev, generator=g)
    gfeat = torch.empty_like(feat)
    L = _lib.lib()
    ws = torch.empty(L.cim_roi_align_workspace_bytes_ex(B, Cf, H, W, K, 7, 7), dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr(dev)

    def fwd():
        _lib.check(L.cim_roi_align_fwd(_lib.ptr(feat), _lib.ptr(rois), _lib.ptr(out), B, Cf, H, W, K, 7, 7, scale, 0, 1,
                                       _lib.ptr(ws), ws.numel(), st), "fwd")

    def bwd():
        _lib.check(L.cim_roi_align_bwd(_lib.ptr(gout), _lib.ptr(rois), _lib.ptr(gfeat), B, Cf, H, W, K, 7, 7, scale, 0,
                                       1, _lib.ptr(ws), ws.numel(), st), "bwd")

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    if a.profile_fused_bwd:
        masks = (torch.rand(K, 7, 7, device=dev, generator=g) > 0.5).float()
        gout2 = torch.randn(K, 2 * Cf, 7, 7, device=dev, generator=g)
        for _ in range(3):
            _lib.check(L.cim_roi_align_maskfuse_bwd(_lib.ptr(gout2), _lib.ptr(rois), _lib.ptr(masks), _lib.ptr(gfeat),
                                                    B, Cf, H, W, K, 7, 7, scale, 0, 1, _lib.ptr(ws), ws.numel(), st),
                       "fbwd")
        torch.cuda.synchronize()
        return
    mb = (K * Cf * 49 * 4 + B * Cf * H * W * 4 + 20 * K) / 1e6
    tf, tb = timeit(fwd), timeit(bwd)
    print(f"{a.backbone} {H}x{W}x{Cf} flags={a.debug_flags}: roi_align fwd {tf:.3f} ms  {mb / tf:.1f} GB/s   bwd {tb:.3f} ms  {mb / tb:.1f} GB/s   ({mb:.0f} MB each)")

    if a.pool:
        argmax = torch.empty(K, Cf, 7, 7, dtype=torch.int32, device=dev)
        for variant, name in ((1, "mmcv"), (0, "legacy")):
            def pfwd():
                _lib.check(L.cim_roi_pool_fwd_ex(_lib.ptr(feat), _lib.ptr(rois), _lib.ptr(out), _lib.ptr(argmax), B, Cf,
                                                 H, W, K, 7, 7, scale, variant, st), "pool fwd")

            def pbwd():
                _lib.check(L.cim_roi_pool_bwd(_lib.ptr(gout), _lib.ptr(argmax), _lib.ptr(rois), _lib.ptr(gfeat), B, Cf,
                                              H, W, K, 7, 7, st), "pool bwd")
            mbp = (2 * K * Cf * 49 * 4 + B * Cf * H * W * 4 + 20 * K) / 1e6          # values + argmax + map
            tpf, tpb = timeit(pfwd), timeit(pbwd)
            print(f"roi_pool[{name}] fwd {tpf:.3f} ms  {mbp / tpf:.1f} GB/s   bwd {tpb:.3f} ms  {mbp / tpb:.1f} GB/s   "
                  f"({mbp:.0f} MB each: pooled values + int32 argmax + the map)")

    if a.fused:
        masks = (torch.rand(K, 7, 7, device=dev, generator=g) > 0.5).float()
        out2 = torch.empty(K, 2 * Cf, 7, 7, device=dev)
        gout2 = torch.randn(K, 2 * Cf, 7, 7, device=dev, generator=g)

        def ffwd():
            _lib.check(L.cim_roi_align_maskfuse_fwd(_lib.ptr(feat), _lib.ptr(rois), _lib.ptr(masks), _lib.ptr(out2), B,
                                                    Cf, H, W, K, 7, 7, scale, 0, 1, _lib.ptr(ws), ws.numel(), st), "ffwd")

        def fbwd():
            _lib.check(L.cim_roi_align_maskfuse_bwd(_lib.ptr(gout2), _lib.ptr(rois), _lib.ptr(masks), _lib.ptr(gfeat),
                                                    B, Cf, H, W, K, 7, 7, scale, 0, 1, _lib.ptr(ws), ws.numel(), st),
                       "fbwd")

        def unfused_fwd():                       # lib/modeling/resnet50.py:131-134 after the plain kernel
            fwd()
            torch.cat((out, out * masks[:, None]), 1, out=out2)

        def unfused_bwd():
            torch.add(gout2[:, :Cf], gout2[:, Cf:] * masks[:, None], out=gout)
            bwd()

        mb2 = (2 * K * Cf * 49 * 4 + B * Cf * H * W * 4 + 20 * K + K * 196) / 1e6
        t1, t2, t3, t4 = timeit(ffwd), timeit(fbwd), timeit(unfused_fwd), timeit(unfused_bwd)
        print(f"maskfuse fused  fwd {t1:.3f} ms  {mb2 / t1:.1f} GB/s   bwd {t2:.3f} ms  {mb2 / t2:.1f} GB/s   ({mb2:.0f} MB each)")
        print(f"maskfuse unfused (kernel + torch mul/cat)  fwd {t3:.3f} ms   bwd {t4:.3f} ms")

    if a.check:
        from torchvision.ops import roi_align as tv
        sel = torch.arange(0, K, 7, device=dev)
        want = tv(feat, rois[sel], 7, scale, 0, True)
        err = (out[sel] - want).abs().max().item() / want.abs().max().item()
        f2 = feat.clone().requires_grad_(True)
        tv(f2, rois, 7, scale, 0, True).backward(gout)
        errb = (gfeat - f2.grad).abs().max().item() / f2.grad.abs().max().item()
        print(f"vs torchvision CUDA: fwd rel err {err:.2e}   bwd rel err {errb:.2e}")


if __name__ == "__main__":
    main()
