#!/usr/bin/env python
"""bench.py -- CIM-head images/s on B200 (BASELINE.json metric).

A step = one pass of the hot path over one batch of synthetic images:
RoIAlign fwd + RoIAlign bwd + mask IoU/containment + scoring heads fwd + 3 x (mining + assignment) + the loss
block incl. PCL_loss (fwd + bwd) + scoring heads bwd (head gradients, averaged over the ranks with one NCCL allreduce when N > 1).
Workload = BASELINE.json configs[1]: ResNet-50 VOC, 8 images x 2000 mask proposals per GPU
(512x512 images -> 1024x32x32 features, 512x512 bit-packed proposal masks, 20 classes).

    python bench.py [--gpus N] [--steps K] [--warmup W]          one JSON line on rank 0
    python bench.py --impl reference ...                          the CPU restatement, host cores

Weak scaling: every rank owns its own 8 images (the path shards per image, no data-path
collective).  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, MAX over ranks.  The per-step working set (>3 GB of RoI gradients
alone) is far larger than the 126 MB L2, so no explicit flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The CPU arm uses every host core whatever the launcher exported: torchrun sets OMP_NUM_THREADS=1 for its workers,
# which made the N > 1 reference lines 3.5x slower than the N = 1 line in round 1.  BLAS / OpenMP read these when
# numpy / torch load, so they are set before either is imported.
if "reference" in sys.argv[1:] and "--impl" in sys.argv[1:] or any(a.startswith("--impl=reference") for a in sys.argv[1:]):
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CIM-head images/s (ROIAlign fwd+bwd + mask IoU + scoring fwd+bwd + mining + losses incl. PCL)"
WORKLOADS = {
    # name: backbone, images per GPU, proposals, classes, present classes, mask side
    "cfg2_r50_voc_8x2000": dict(backbone="resnet50", n_img=8, R=2000, C=20, present=2, mask=512),
    "cfg3_vgg16_voc_8x2000": dict(backbone="vgg16", n_img=8, R=2000, C=20, present=2, mask=512),
    "cfg4_r50_coco_8x2000_q": dict(backbone="resnet50", n_img=8, R=2000, C=80, present=4, mask=128),
    "cfg1_r50_voc_1x300": dict(backbone="resnet50", n_img=1, R=300, C=20, present=2, mask=512),
    "tiny": dict(backbone="resnet50", n_img=2, R=200, C=20, present=2, mask=128),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def algorithmic_bytes(cfg, Cf, H, W):
    """SURVEY.md section 8d, bytes per image and stage."""
    R, C, hwm = cfg["R"], cfg["C"], cfg["mask"] ** 2
    F, O = Cf * H * W * 4, R * Cf * 49 * 4
    return {
        "roi_align_fwd": F + 20 * R + O,
        "roi_align_bwd": O + 20 * R + F,
        "mask_overlap": R * hwm // 8 + 2 * R * R * 2,
        "score_heads": R * 4096 * 4 + 8 * (C + 1) * 4097 * 4 + 8 * R * (C + 1) * 4,
        "mine_assign": 3 * (2 * R * R * 2 + R * (C + 1) * 4 + R * 6),
        # scoring backward: x read, grad_x written, scores and grad_scores read, weights read and head
        # gradients written
        "score_heads_bwd": 2 * R * 4096 * 4 + 2 * 8 * R * (C + 1) * 4 + 2 * 8 * (C + 1) * 4097 * 4,
        # loss block: scores read, pseudo labels / weights read, grad_scores written
        "head_losses": 2 * 8 * R * (C + 1) * 4 + 3 * R * (C + 1) * 4 + 3 * R * 6 + R * (C + 1) * 4,
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, smax, reasons = [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=smax, reasons=sorted(reasons),
                    samples=len(sm))


# ------------------------------------------------------------------------------------ CPU arm
def literal_mask_utils(cfg):
    """The reference's LITERAL mask_utils double loop (one Python-level pair at a time, create_cob_iou.py:43-48)
    extrapolated to this workload by its pair count from the us/pair measured in the build container with the
    unmodified reference file (tools/literal_mask_utils_timing.py -> profiles/literal_mask_utils.json; the reference
    tree does not travel to the GPU box).  The CPU arm itself times the vectorised restatement (M . M^T)."""
    try:
        with open(os.path.join(ROOT, "profiles", "literal_mask_utils.json")) as f:
            t = json.load(f)
        us = t["sizes"][str(cfg["mask"])]["us_per_pair_both_maps"]
    except (OSError, ValueError, KeyError):
        return None
    return {"s_per_image": us * 1e-6 * cfg["R"] ** 2, "us_per_pair_iou_plus_asy": us,
            "measured_on": f"build container, 1 of {t['host_cores']} cores (the loop is single-threaded numpy), "
                           f"{t['n_masks']} masks of {cfg['mask']}^2", "pairs_per_image": cfg["R"] ** 2}


def cpu_reference(cfg, steps, warmup, sample_rois=128, head_grads=True):
    """The reference's algorithm for the path, restated for CPU (oracle/), timed on the host cores
    on a BOUNDED sample of the workload: one image; RoIAlign fwd+bwd and the mask overlap on
    `sample_rois` of its R proposals (cost scaled by R / sample_rois: both are linear in the
    number of rows processed); scoring and the 3 mining layers at full R."""
    import torch
    from cim_b200 import synth
    from oracle import heads_oracle, loss_oracle, roi_oracle
    ncpu = os.cpu_count() or 1
    torch.set_num_threads(ncpu)
    R, C, S = cfg["R"], cfg["C"], min(sample_rois, cfg["R"])
    Cf, H, W, scale = synth.feature_shape(cfg["backbone"])
    params = synth.proposal_params(R, 512, 1234)
    rois = synth.rois_from_params(params).numpy()
    feat = np.random.RandomState(0).randn(1, Cf, H, W).astype(np.float32)
    g_out = np.random.RandomState(1).randn(S, Cf, 7, 7).astype(np.float32)
    masks = synth.rasterize(params, out_size=cfg["mask"]).numpy().reshape(R, -1)
    mf = masks.astype(np.float32)
    torch.manual_seed(0)
    w = [np.random.RandomState(10 + i).uniform(-1 / 64, 1 / 64, (C + 1, 4096)).astype(np.float32) for i in range(8)]
    b = [np.zeros(C + 1, np.float32) for _ in range(8)]
    x = np.random.RandomState(2).randn(R, 4096).astype(np.float32)
    labels = synth.image_labels(C, cfg["present"], 1234).numpy()
    cmat = synth.cluster_mat(R, C, np.nonzero(labels[0])[0], 6, 1234).numpy()
    from oracle import mask_oracle
    iou16, asy16 = mask_oracle.mask_overlap_maps(masks[:, ::max(1, masks.shape[1] // 4096)])   # setup only

    def one_step():
        t0 = time.perf_counter()
        roi_oracle.roi_align_fwd(feat, rois[:S], 7, 7, scale, 0, True)
        roi_oracle.roi_align_bwd(g_out, rois[:S], feat.shape, scale, 0, True)
        t_roi = time.perf_counter() - t0
        t0 = time.perf_counter()
        inter = mf[:S] @ mf.T                                       # S x R intersection counts
        area = mf.sum(1)
        with np.errstate(divide="ignore", invalid="ignore"):
            (inter / (area[:S, None] + area[None, :] - inter)).astype(np.float16)
            (inter / area[None, :]).astype(np.float16)
        t_mask = time.perf_counter() - t0
        t0 = time.perf_counter()
        p_cls, p_det, r_cls, r_iou = heads_oracle.score_heads(x, w, b)
        t_score = time.perf_counter() - t0
        t0 = time.perf_counter()
        cls_l, det_l = [p_cls, r_cls[0], r_cls[1]], [p_det, r_iou[0], r_iou[1]]
        assigned = []
        for l in range(3):
            assigned.append(heads_oracle.cim_layer_forward(cls_l[l], det_l[l], labels, iou16, asy16, 0.1,
                                                           0.25 + 0.1 * l, 0.5 + 0.1 * l, 0.85, True))
        t_mine = time.perf_counter() - t0
        if head_grads:                                   # loss block fwd+bwd, then the scoring backward
            t0 = time.perf_counter()
            ok = [a[0] is not None for a in assigned]
            pl = np.stack([a[0] if o else np.zeros((R, C + 1), np.float32) for a, o in zip(assigned, ok)])[:, None]
            pi = np.stack([a[1] if o else np.zeros(R, np.float16) for a, o in zip(assigned, ok)])[:, None]
            lw = np.stack([a[2] if o else np.zeros(R, np.float32) for a, o in zip(assigned, ok)])[:, None]
            sc = np.stack([p_cls, p_det] + r_cls + r_iou)
            _, g = loss_oracle.head_losses(sc, pl, pi, lw, np.array(ok, np.uint8)[:, None], labels.reshape(1, -1), 3,
                                           dtype=torch.float32)
            _, g_pcl = loss_oracle.pcl_losses(p_cls, cmat[None], dtype=torch.float32)
            g[0] += g_pcl
            heads_oracle.score_heads_bwd(x, w, b, list(g))
            t_score += time.perf_counter() - t0
        return (t_roi + t_mask) * (R / S) + t_score + t_mine, dict(roi=t_roi * R / S, mask=t_mask * R / S,
                                                                   score=t_score, mine=t_mine)

    np.random.seed(3)
    for _ in range(warmup):
        one_step()
    per_image, parts = [], None
    t_wall = time.perf_counter()
    for _ in range(steps):
        t, parts = one_step()
        per_image.append(t)
    wall = time.perf_counter() - t_wall
    sec = float(np.mean(per_image))
    return dict(images_per_s=1.0 / sec, sec_per_image=sec, cores=ncpu, wall_s=wall, parts=parts,
                literal=literal_mask_utils(cfg),
                sample=f"1 image of {cfg['R']} proposals; RoIAlign fwd+bwd and mask overlap on {S} proposal rows "
                       f"(x{R / S:.1f}), scoring {'fwd+bwd + loss block' if head_grads else 'fwd'} + 3 mining layers at full size; "
                       f"numpy/OpenMP on all host cores")


# ------------------------------------------------------------------------------------ GPU arm
def build_inputs(cfg, dev, seed_base):
    import torch
    from cim_b200 import heads, mask_ops, synth
    n_img, R, C = cfg["n_img"], cfg["R"], cfg["C"]
    Cf, H, W, scale = synth.feature_shape(cfg["backbone"])
    gen = torch.Generator(device=dev).manual_seed(seed_base)
    feat = torch.randn(n_img, Cf, H, W, device=dev, generator=gen)
    grad_out = torch.randn(n_img * R, Cf, 7, 7, device=dev, generator=gen)
    seg_x = torch.randn(n_img * R, 4096, device=dev, generator=gen)
    rois, packed, packed_flat, labels, mats = [], [], [], [], []
    kb_per_row = cfg["mask"] // 16 if mask_ops.tiled_ok(cfg["mask"], cfg["mask"]) else 0
    for b in range(n_img):
        params = synth.proposal_params(R, 512, seed_base + b)
        rois.append(synth.rois_from_params(params, b))
        m = synth.rasterize(params, device=dev, out_size=cfg["mask"])
        packed.append(mask_ops.mask_pack(m))                         # tiled 8 x 16 patches when the size allows
        packed_flat.append(mask_ops.mask_pack(m, layout="flat").cpu())   # only to derive the host wire format
        labels.append(synth.image_labels(C, cfg["present"], seed_base + b))
        mats.append(synth.cluster_mat(R, C, np.nonzero(labels[-1][0].numpy())[0], 6, seed_base + b))
    torch.manual_seed(0)
    model = heads.cls_iou_model(4096, C + 1, 3).to(dev)
    weight, bias = (t.detach().contiguous() for t in model._stacked())
    labels = torch.cat(labels)
    return dict(feat=feat, rois=torch.cat(rois).to(dev), grad_out=grad_out, packed=torch.stack(packed),
                packed_flat=torch.stack(packed_flat), kb_per_row=kb_per_row, mat=torch.stack(mats).to(dev),
                seg_x=seg_x, weight=weight, bias=bias, labels=labels.to(dev), labels_host=labels.numpy(),
                shape=(Cf, H, W, scale))


def measured_traffic(workload, stage):
    """DRAM bytes per launch of the stage's main kernel from the committed ncu --set full capture
    (profiles/traffic.json), or None when there is no capture for this workload / stage."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if t.get("workload") != workload or stage not in t:
            return None
        return int(t[stage]["dram_read"] + t[stage]["dram_write"])
    except (OSError, ValueError, KeyError):
        return None


def visited_kblocks(step, cfg):
    """Number of 128-pixel K-blocks the tensor-core overlap kernel visited in its last launch (it leaves the
    count in the first 8 bytes of its workspace, include/cimhead.h) and the dense total: 128 x 256 tiles on or
    right of the diagonal x all K-blocks."""
    import torch
    torch.cuda.synchronize()
    visited = int(step.overlap_ws[:8].view(torch.int64).item())
    n, kblocks = cfg["R"], step.words // 4
    nrb, ncb = (n + 127) // 128, (n + 255) // 256
    tiles = sum(ncb - (i >> 1) for i in range(nrb))
    return visited, tiles * cfg["n_img"] * kblocks


def time_stages(step, inp, iters=5):
    """Per-stage device time (ms) with CUDA events, each stage launched back to back `iters` times."""
    import ctypes as C
    import torch
    from cim_b200 import _lib
    L, P, p = step.L, _lib.ptr, step.p
    st = _lib.stream_ptr(step.dev)
    n_img, R = step.n_img, step.R
    calls = {
        "roi_align_fwd": lambda: L.cim_roi_align_fwd(P(inp["feat"]), P(inp["rois"]), P(step.roi_out), n_img, step.Cf,
                                                     step.H, step.W, n_img * R, 7, 7, step.scale, 0, 1,
                                                     P(step.roi_ws), step.roi_ws.numel(), st),
        "roi_align_bwd": lambda: L.cim_roi_align_bwd(P(inp["grad_out"]), P(inp["rois"]), P(step.grad_feat), n_img,
                                                     step.Cf, step.H, step.W, n_img * R, 7, 7, step.scale, 0, 1,
                                                     P(step.roi_ws), step.roi_ws.numel(), st),
        "mask_overlap": lambda: L.cim_mask_overlap_ex(P(inp["packed"]), n_img, R, step.words, step.kb_per_row, None,
                                                      P(step.area), P(step.iou), P(step.asy), P(step.overlap_ws),
                                                      step.overlap_ws.numel(), 0, st),
        "score_heads": lambda: L.cim_score_heads(P(inp["seg_x"]), P(inp["weight"]), P(inp["bias"]), P(step.scores),
                                                 n_img, R, step.D, step.C + 1, step.K, P(step.score_ws),
                                                 step.score_ws.numel(), st),
        "score_heads_bwd": lambda: L.cim_score_heads_bwd(P(inp["seg_x"]), P(inp["weight"]), P(step.scores),
                                                         P(step.grad_scores), P(step.grad_seg_x), P(step.grad_weight),
                                                         P(step.grad_bias), n_img, R, step.D, step.C + 1, step.K,
                                                         P(step.score_bwd_ws), step.score_bwd_ws.numel(), st),
        "head_losses": lambda: (L.cim_head_losses(P(step.scores), P(step.pseudo_labels), P(step.pseudo_iou),
                                                  P(step.loss_weights), P(step.valid), P(inp["labels"]), P(step.losses),
                                                  P(step.grad_scores), n_img, R, step.C, step.K, step.K, 3.0, 1.0, 3.0,
                                                  1.0 / n_img, st) or
                                L.cim_pcl_loss(P(step.scores), P(inp["mat"]), P(step.pcl_loss), P(step.grad_scores), n_img,
                                               R, step.C + 1, 127, 1.0 / n_img, 1, st)),
        "mine": lambda: L.cim_mine(C.byref(p), step.cls_ptrs, step.det_ptrs, P(inp["labels"]), P(step.iou),
                                   P(step.asy), P(step.gt_count), P(step.gt_rows), P(step.gt_class),
                                   P(step.gt_weight), P(step.asy_flag), P(step.mine_ws), step.mine_ws.numel(), st),
        "assign": lambda: L.cim_assign(C.byref(p), P(step.iou), P(step.gt_count), P(step.gt_rows), P(step.gt_class),
                                       P(step.gt_weight), None, P(step.pseudo_labels), P(step.pseudo_iou),
                                       P(step.loss_weights), P(step.valid), st),
    }
    if not step.head_grads:
        del calls["score_heads_bwd"], calls["head_losses"]
    out = {}
    for name, fn in calls.items():
        _lib.check(fn(), name)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            _lib.check(fn(), name)
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / iters
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_r50_voc_8x2000", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-anti-noise", action="store_true")
    ap.add_argument("--no-head-grads", action="store_true",
                    help="leave the loss block, the scoring backward and the allreduce of the head gradients out of the step")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: everything else a library may write to fd 1 (NCCL prints its
    # version banner there on the first collective) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    cfg = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = dict(workload=args.workload, images_per_gpu=cfg["n_img"], proposals_per_image=cfg["R"],
                  classes=cfg["C"], backbone=cfg["backbone"], mask_px=cfg["mask"], sharding="per image, weak",
                  l2="per-step working set > 3 GB >> 126 MB L2, no explicit flush")

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, min(args.steps, 3)), 1          # bounded; one untimed pass warms BLAS / pages
        r = cpu_reference(cfg, steps, warmup, head_grads=not args.no_head_grads)
        emit(({
            "impl": "reference", "metric": METRIC, "value": r["images_per_s"], "unit": "images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * r["sec_per_image"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": r["images_per_s"], "unit": "images/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"], "parts_s_per_image": r["parts"],
                             "literal_mask_utils_s_per_image": r["literal"]},
            "e2e": {"value": r["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import torch
    from cim_b200 import dist as cdist
    from cim_b200.step import CIMHeadStep, KERNELS_HEAD_GRADS, KERNELS_PCL, KERNELS_PER_STEP
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    rank, world, local = cdist.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    peaks = load_peaks()

    inp = build_inputs(cfg, dev, 1234 + 1000 * rank)
    Cf, H, W, scale = inp["shape"]
    words = inp["packed"].shape[-1]
    step = CIMHeadStep(cfg["n_img"], cfg["R"], cfg["C"], Cf, H, W, scale, words, anti_noise_sampling=not args.no_anti_noise,
                       max_present=max(4, 2 * cfg["present"]), device=dev, mask_kb_per_row=inp["kb_per_row"],
                       head_grads=not args.no_head_grads)
    mat = None if args.no_head_grads else inp["mat"]
    run = lambda: step.run(inp["feat"], inp["rois"], inp["grad_out"], inp["packed"], inp["seg_x"], inp["weight"],
                           inp["bias"], inp["labels"], inp["labels_host"], mat=mat)
    np.random.seed(3)
    for _ in range(max(args.warmup, 3)):
        run()
    torch.cuda.synchronize()
    cdist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    cdist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = cdist.max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
    n_images = cfg["n_img"] * world
    value = n_images / (ms_step * 1e-3)

    # end to end through the public call with HOST inputs (rois, labels, bit-packed masks) and
    # results read back to the host every step
    # host wire format of the proposal masks: bounding-box crops, bit-packed (mask_ops.MaskCrops)
    from cim_b200 import mask_ops
    crops = mask_ops.crops_from_packed_host(inp.pop("packed_flat").view(cfg["n_img"] * cfg["R"], -1), cfg["mask"],
                                            cfg["mask"])
    step.alloc_host_io(mask_hw=(cfg["mask"], cfg["mask"]), crop_capacity_words=int(crops.words.numel() * 1.25) + 1024)
    step.hi_rois.copy_(inp["rois"])
    step.hi_labels.copy_(inp["labels"])
    step.set_host_crops(crops)
    run_host = lambda: step.run_host(inp["feat"], inp["grad_out"], inp["seg_x"], inp["weight"], inp["bias"], mat=mat,
                                     lag_results=True)
    # the step runs on a HIGH-priority stream: the prefetch of the next step's inputs (H2D + the crop-unpack kernel on
    # the step's low-priority copy stream) then only takes SMs the step's own kernels are not waiting for
    hp = torch.cuda.Stream(device=dev, priority=-1)
    hp.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(hp):
        for _ in range(3):
            run_host()
        step.flush_results()
        cdist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            run_host()                 # reads the previous step's results on the host while this step computes
        step.flush_results()           # ... and the last step's: every timed step's H2D, D2H and host wait are inside
        e1.record()
        torch.cuda.synchronize()
    ms_e2e = cdist.max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
    e2e_value = n_images / (ms_e2e * 1e-3)

    if step.trace is not None and rank == 0:                 # host timeline of the last e2e steps (CIM_STEP_TRACE=1)
        ev = step.trace[-5 * 3:]
        for name, t in ev:
            print(f"trace {name:16s} {(t - ev[0][1]) * 1e3:8.3f} ms", file=sys.stderr)
    stages = time_stages(step, inp) if rank == 0 else None
    cdist.barrier()
    cdist.shutdown()
    if rank != 0:
        return

    bytes_img = algorithmic_bytes(cfg, Cf, H, W)
    if args.no_head_grads:
        del bytes_img["score_heads_bwd"], bytes_img["head_losses"]
    stage_ms = dict(stages)
    stage_ms["mine_assign"] = stage_ms.pop("mine") + stage_ms.pop("assign")
    table = {}
    for name, ms in stage_ms.items():
        gbs = bytes_img[name] * cfg["n_img"] / (ms * 1e-3) / 1e9
        table[name] = dict(ms=round(ms, 4), algorithmic_mb=round(bytes_img[name] * cfg["n_img"] / 1e6, 2),
                           gb_per_s=round(gbs, 1), hbm_frac=round(gbs / peaks["hbm_gbs"], 4))
    dominant = max(stage_ms, key=stage_ms.get)
    total_bytes = sum(bytes_img.values())
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": table[dominant]["gb_per_s"], "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": table[dominant]["hbm_frac"], "traffic": measured_traffic(args.workload, dominant),
                "peak_source": peaks["source"],
                "share_of_step": round(stage_ms[dominant] / sum(stage_ms.values()), 3)}
    if dominant == "mask_overlap":
        # a dense R x HW x R contraction of 0/1 operands: tensor-pipe bound.  Algorithmic
        # work = the symmetric half, R^2 * HW MAC-flops per image.  The kernel sorts the masks by
        # position and skips K-blocks where one operand block is all zero, so `achieved` counts the
        # MMA flops actually EXECUTED (visited K-blocks x 2*128*256*128); the algorithmic-equivalent
        # rate is given next to it.  Peak: int8 runs at twice the bf16 rate on sm_100;
        # MEASURED_PEAKS.json only has bf16, so peak = 2 x measured bf16 (burst: kernel timed alone).
        alg = float(cfg["R"]) ** 2 * cfg["mask"] ** 2 * cfg["n_img"]
        visited, total = visited_kblocks(step, cfg)
        executed = visited * 2.0 * 128 * 256 * 128
        secs = stage_ms[dominant] * 1e-3
        roofline.update({"bound": "tensor", "achieved": round(executed / secs / 1e12, 1),
                         "peak": round(2 * peaks["bf16_tflops"], 1), "unit": "TFLOP/s",
                         "frac": round(executed / secs / 1e12 / (2 * peaks["bf16_tflops"]), 4),
                         "peak_note": "int8 = 2 x measured bf16 burst",
                         "executed_kblock_fraction": round(visited / total, 4),
                         "algorithmic_equivalent_tflops": round(alg / secs / 1e12, 1),
                         "algorithmic": "R^2*HW MACs per image counted as flops (upper triangle only)"})
    result = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(step.h2d_bytes + step.last_mask_h2d_bytes),
                "d2h_bytes_per_step": int(step.d2h_bytes),
                "host_inputs": "rois, labels, bbox-cropped bit-packed proposal masks (unpacked on the device); "
                               "features/seg_x/grad_out are device-produced",
                "host_outputs": "per-image losses [n_img, K+1, 3], valid flags, two checksums of the RoIAlign outputs, "
                                "plus the pseudo-GT lists of the sampling hop" if not args.no_head_grads else
                                "pseudo labels / IoU labels / loss weights, valid flags, checksums, sampling-hop lists",
                "pipelining": "H2D of step i+1 on a copy stream overlaps the kernels of step i; the results of step i "
                              "are copied D2H at its end and read by the host (one event wait per step) while step "
                              "i+1 runs its RoIAlign forward; the last step's wait is inside the timed region"},
        "gpu_launches": (KERNELS_PER_STEP + (0 if args.no_head_grads else KERNELS_HEAD_GRADS + KERNELS_PCL)) * args.steps,
        "collective": ("none (single process)" if world == 1 else
                       f"NCCL allreduce (avg) of the {step.head_bucket.numel() * 4} B head-gradient bucket per step, "
                       "inside the timed region, overlapped with the RoIAlign kernels") if not args.no_head_grads
        else "none (head gradients excluded)",
        "roofline": roofline,
        "step_roofline": {"algorithmic_mb_per_image": round(total_bytes / 1e6, 1),
                          "hbm_frac": round(value / world * total_bytes / 1e9 / peaks["hbm_gbs"], 4)},
        "stages": table, "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(cfg, 1, 0, head_grads=not args.no_head_grads)
        result["cpu_baseline"] = {"value": r["images_per_s"], "unit": "images/s", "cores": r["cores"], "kind": "port",
                                  "sample": r["sample"], "parts_s_per_image": r["parts"],
                                  "literal_mask_utils_s_per_image": r["literal"]}
    emit(result)


if __name__ == "__main__":
    main()
