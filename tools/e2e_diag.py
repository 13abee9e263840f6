#!/usr/bin/env python
"""Where does the end-to-end step (host inputs, results read back) lose time against the device-only step?
Times run() and run_host() at cfg2 with pieces of the host path switched off (diagnostic, not a bench)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cim_b200 import mask_ops
from cim_b200.step import CIMHeadStep

cfg = bench.WORKLOADS["cfg2_r50_voc_8x2000"]
dev = torch.device("cuda:0")
inp = bench.build_inputs(cfg, dev, 1234)
Cf, H, W, scale = inp["shape"]
step = CIMHeadStep(cfg["n_img"], cfg["R"], cfg["C"], Cf, H, W, scale, inp["packed"].shape[-1],
                   device=dev, mask_kb_per_row=inp["kb_per_row"], head_grads=True, rng="stream")
crops = mask_ops.crops_from_packed_host(inp.pop("packed_flat").view(cfg["n_img"] * cfg["R"], -1), cfg["mask"], cfg["mask"])
step.alloc_host_io(mask_hw=(cfg["mask"], cfg["mask"]), crop_capacity_words=int(crops.words.numel() * 1.25) + 1024)
step.hi_rois.copy_(inp["rois"]); step.hi_labels.copy_(inp["labels"]); step.set_host_crops(crops)
mat = inp["mat"]

def timed(fn, n=10, flush=None):
    for _ in range(3):
        fn()
    if flush: flush()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    if flush: flush()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

np.random.seed(3)
run = lambda: step.run(inp["feat"], inp["rois"], inp["grad_out"], inp["packed"], inp["seg_x"], inp["weight"], inp["bias"],
                       inp["labels"], mat=mat)
hp = torch.cuda.Stream(device=dev, priority=-1)
print(f"run() default stream          {timed(run):.3f} ms")
with torch.cuda.stream(hp):
    print(f"run() high-priority stream    {timed(run):.3f} ms")
    host = lambda **kw: step.run_host(inp["feat"], inp["grad_out"], inp["seg_x"], inp["weight"], inp["bias"], mat=mat, **kw)
    print(f"run_host lag_results          {timed(lambda: host(lag_results=True), flush=step.flush_results):.3f} ms")
    print(f"run_host sync every step      {timed(lambda: host(lag_results=False)):.3f} ms")
    print(f"run_host lag, no prefetch     {timed(lambda: host(lag_results=True, prefetch_next=False), flush=step.flush_results):.3f} ms")
    # no crop unpack: full packed masks over PCIe instead
    step2 = CIMHeadStep(cfg["n_img"], cfg["R"], cfg["C"], Cf, H, W, scale, inp["packed"].shape[-1],
                        device=dev, mask_kb_per_row=inp["kb_per_row"], head_grads=True)
    step2.alloc_host_io()
    step2.hi_rois.copy_(inp["rois"]); step2.hi_labels.copy_(inp["labels"]); step2.hi_masks.copy_(inp["packed"].cpu())
    host2 = lambda: step2.run_host(inp["feat"], inp["grad_out"], inp["seg_x"], inp["weight"], inp["bias"], mat=mat, lag_results=True)
    print(f"run_host lag, full masks H2D  {timed(host2, flush=step2.flush_results):.3f} ms  ({step2.h2d_bytes / 1e6:.0f} MB / step)")
    # ---- pieces of the host path switched off one at a time
    real_stage = step.stage_host_inputs
    step.stage_host_inputs = lambda defer_kernels=False, direct=False: ((lambda after=None: None) if defer_kernels else step)   # no H2D, no unpack
    print(f"run_host lag, NO staging      {timed(lambda: host(lag_results=True), flush=step.flush_results):.3f} ms   (host loop + result read-back only)")
    step.stage_host_inputs = real_stage

    def stage_only():
        real_stage()
        step._slot ^= 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(step.copy_stream)
    for _ in range(5):
        real_stage()
    e1.record(step.copy_stream)
    torch.cuda.synchronize()
    print(f"staging alone (copy stream)   {e0.elapsed_time(e1) / 5:.3f} ms   (H2D of the crops + unpack/meta kernel, nothing else running)")
    # ---- which half of the staging costs the 0.3 ms: the H2D copies or the unpack / metadata kernel?
    def stage_copies_only(defer_kernels=False, direct=False):
        k = real_stage(defer_kernels=True, direct=direct)
        buf = step.di[step._slot ^ 1] if step._staged else step.di[step._slot]
        with torch.cuda.stream(step.copy_stream):
            buf["ready"].record(step.copy_stream)
        return (lambda after=None: None) if defer_kernels else step
    step.stage_host_inputs = stage_copies_only
    print(f"run_host lag, H2D copies only {timed(lambda: host(lag_results=True), flush=step.flush_results):.3f} ms")

    def stage_kernels_only(defer_kernels=False, direct=False):
        buf = step.di[step._slot ^ 1] if step._staged else step.di[step._slot]
        def kernels(after=None):
            with torch.cuda.stream(step.copy_stream):
                if after is not None:
                    step.copy_stream.wait_event(after)
                step.copy_stream.wait_event(buf["free"])
                rc = step.L.cim_mask_unpack_crops_tiled_meta_sparse(
                    step.di[0]["crop_words"].data_ptr(), step.di[0]["crop_meta"].data_ptr(), step.di[0]["crop_off"].data_ptr(),
                    buf["masks"].data_ptr(), buf["prev_rects"].data_ptr(), buf["meta"].data_ptr(), buf["meta"].numel(),
                    step.n_img, step.R, step.mask_hw[0], step.mask_hw[1], step.words, step.copy_stream.cuda_stream)
                assert rc == 0
                buf["ready"].record(step.copy_stream)
        if defer_kernels:
            return kernels
        kernels()
        return step
    step.stage_host_inputs = stage_kernels_only
    print(f"run_host lag, kernels only    {timed(lambda: host(lag_results=True), flush=step.flush_results):.3f} ms")
    step.stage_host_inputs = real_stage
