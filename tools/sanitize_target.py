"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck): forward (two-lane sampling forced on),
backward, backward_batch (worker streams), hypothesis assignment, reprojection loss (vector and scalar paths)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import esac_b200.api as api  # noqa: E402
from esac_b200.synth import make_scene  # noqa: E402

ctx = api.context()
sc = make_scene(E=3, H=24, W=32, M=1024, sub=8, seed=2, active_only=False)   # M >= 1024: two sampling lanes
coords = torch.from_numpy(sc.coords).cuda()
assign = torch.from_numpy(sc.assign).cuda()
out = torch.zeros(4, 4, device="cuda")
print("forward", api.forward(coords, assign, out, *sc.params), ctx.stats()["kernel_launches"])
# multi-CTA refinement groups (LL exchange, shared-memory inlier lists) and an uncached share (global inlier list)
big = make_scene(E=1, H=120, W=160, M=16, sub=4, seed=5)
cb, ab = torch.from_numpy(big.coords).cuda(), torch.from_numpy(big.assign).cuda()
for grp in (0, 2):
    ctx.set_option("refine_group", grp)
    print("forward 120x160 group", grp, api.forward(cb, ab, out, *big.params), ctx.stats()["refine_group"])
ctx.set_option("refine_group", 0)
sc = make_scene(E=2, H=24, W=32, M=48, sub=8, seed=3)
coords = torch.from_numpy(sc.coords).cuda()
grads = torch.zeros_like(coords)
print("backward", api.backward(coords, grads, torch.from_numpy(sc.assign).cuda(), torch.from_numpy(sc.gt_pose), 1.0, 100.0, 100.0,
                               *sc.params))
B = 3
cs = torch.stack([coords] * B)
gs = torch.zeros_like(cs)
print("backward_batch", api.backward_batch(cs, gs, torch.stack([torch.from_numpy(sc.assign).cuda()] * B),
                                           torch.stack([torch.from_numpy(sc.gt_pose)] * B).cuda(), 1.0, 100.0, 100.0, [0, 1, -1], 0,
                                           *sc.params[2:]))
w = torch.rand(4, 9, device="cuda")
a, h = api.assign_hypotheses(w, 300, 5, maxExperts=4)
print("assign", h.sum().item())
for (H, W) in [(24, 32), (23, 31)]:
    p = make_scene(E=1, H=H, W=W, M=8, sub=8, seed=4)
    pred = torch.from_numpy(p.coords).cuda()
    g = torch.empty_like(pred)
    print("reproj", api.reproj_loss(pred, torch.from_numpy(p.gt_pose)[None], 525.0, 1, 2, 10.0, 8, outGradients=g))
torch.cuda.synchronize()
