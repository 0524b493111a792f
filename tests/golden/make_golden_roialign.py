"""Golden vectors for the crop & resize op (tests/golden/ref_roialign.npz), made in the build container with
torchvision's CPU roi_align -- the op behind detectron2's ROIAlign, which the reference's batch_crop_resize
(core/utils/zoom_utils.py:80-95) instantiates as ROIAlign(output_size, 1.0, 0, aligned=True).
    python tests/golden/make_golden_roialign.py
"""
import os

import numpy as np
import torch
import torchvision
from torchvision.ops import roi_align

HERE = os.path.dirname(os.path.abspath(__file__))
out = {"torchvision_version": np.array(torchvision.__version__)}
cases = {
    # tag: (N, C, H, W, out_h, out_w, aligned, sampling_ratio, rois)
    "a": (2, 3, 20, 24, 8, 8, True, 0, [[0, 2.3, 3.1, 17.9, 15.2], [1, -4.0, -3.0, 9.5, 8.0], [0, 5.0, 4.0, 30.0, 26.0],
                                        [1, 10.2, 9.7, 10.9, 10.1], [0, 12.0, 9.0, 6.0, 3.0], [1, 0.0, 0.0, 24.0, 20.0]]),
    "b": (1, 5, 17, 13, 5, 7, False, 0, [[0, 1.2, 2.2, 11.0, 15.5], [0, 3.3, 3.3, 3.6, 3.5], [0, -2.0, 1.0, 20.0, 9.0]]),
    "c": (2, 2, 12, 12, 4, 4, True, 2, [[1, 0.5, 0.5, 11.5, 11.5], [0, 3.0, 2.0, 9.0, 10.0]]),
}
g = torch.Generator().manual_seed(11)
for tag, (N, C, H, W, oh, ow, aligned, sr, rois) in cases.items():
    x = torch.randn(N, C, H, W, generator=g, requires_grad=True)
    r = torch.tensor(rois, dtype=torch.float32)
    go = torch.randn(len(rois), C, oh, ow, generator=g)
    y = roi_align(x, r, (oh, ow), 1.0, sr, aligned)
    (y * go).sum().backward()
    for k, v in (("x", x.detach()), ("rois", r), ("go", go), ("y", y.detach()), ("gx", x.grad)):
        out[f"{tag}_{k}"] = v.numpy()
    out[f"{tag}_cfg"] = np.array([oh, ow, int(aligned), sr])
np.savez_compressed(os.path.join(HERE, "ref_roialign.npz"), **out)
print("wrote", {k: v.shape for k, v in out.items()})
