"""One launch shape of the reprojection-loss kernel for ncu (B=64 full-resolution maps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import esac_b200.api as api  # noqa: E402
from esac_b200.synth import make_scene  # noqa: E402

sc = make_scene(E=1, H=480, W=640, M=8, sub=1, seed=5)
B = 64
pred = torch.from_numpy(sc.coords[0]).cuda().unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
gts = torch.from_numpy(sc.gt_pose).unsqueeze(0).repeat(B, 1, 1).contiguous()
grads = torch.empty_like(pred)
for _ in range(3):
    api.reproj_loss(pred, gts, 525.0, 1, -1, 10.0, 1, outGradients=grads)
torch.cuda.synchronize()
