#!/usr/bin/env python
"""Post-render crop & resize: 32 full-frame renders (channels-last, as the renderer writes them) ->
256x256 colour crops and 64x64 normal crops.  Ours (dibr_roi_align_*) vs torchvision's CUDA roi_align -- the op behind
detectron2's ROIAlign -- fed the way the reference feeds it (permuted view, made contiguous by the op).  CUDA events,
median of 20 after 5 warm-ups, L2 flushed between repetitions.  Algorithmic bytes: the forward reads the roi's pixels and
writes the crops, the backward reads the crop gradients and writes the dense frame gradient."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchvision
from self6dpp_b200.zoom_utils import batch_crop_resize

DEV = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
B, H, W = 32, 480, 640
frames = torch.rand(B, H, W, 3, generator=g).to(DEV)
cc = torch.stack([torch.rand(B, generator=g) * 400 + 120, torch.rand(B, generator=g) * 280 + 100], dim=1)
hh = torch.rand(B, 1, generator=g) * 60 + 60                       # 120..240 px boxes
rois = torch.cat([torch.arange(B).float().view(-1, 1), cc - hh, cc + hh], dim=1).to(DEV)


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


for res in (256, 64):
    go = torch.randn(B, 3, res, res, generator=g).to(DEV)
    x = frames.clone().requires_grad_(True)
    xv = x.permute(0, 3, 1, 2)
    row = {"config": f"batch_crop_resize {B} x [3,{H},{W}] -> [3,{res},{res}]"}
    for name, op in (("ours", lambda: batch_crop_resize(xv, rois, res, res)),
                     ("torchvision", lambda: torchvision.ops.roi_align(xv, rois, (res, res), 1.0, 0, True))):
        row[name + "_fwd_ms"] = timed(op)
        y = op()
        row[name + "_bwd_ms"] = timed(lambda: torch.autograd.grad(y, x, go, retain_graph=True))
    roi_px = float(((2 * hh) ** 2).sum())
    fwd_bytes = 4 * 3 * (roi_px + B * res * res)
    bwd_bytes = 4 * 3 * (B * res * res + B * H * W)
    row["fwd_GBps"] = fwd_bytes / row["ours_fwd_ms"] / 1e6
    row["bwd_GBps"] = bwd_bytes / row["ours_bwd_ms"] / 1e6
    print(json.dumps(row))
