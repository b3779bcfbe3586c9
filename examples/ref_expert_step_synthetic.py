"""The expert-refinement training step (code/ref_expert.py:95-160) with its loss block replaced by the fused kernel.

ref_expert.py builds the reprojection loss from ~15 torch ops per step (projection, clamps, norm, two masked sums) and lets
autograd differentiate them; `esac_b200.autograd.reproj_loss` is the same loss and gradient as ONE kernel
(esac_b200/csrc/reproj.cu).  Dataset and network are stand-ins (esac_b200.compat, a 1x1-conv "expert"); the loop itself is the
reference's.  With --check every step is also evaluated with the original op sequence (oracle/reproj_loss_oracle.py).

    python examples/ref_expert_step_synthetic.py --iterations 5 --check
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import torch
import torch.nn as nn

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from esac_b200.autograd import reproj_loss  # noqa: E402
from esac_b200.compat import OUTPUT_SUBSAMPLE, SyntheticRoomDataset, random_shift  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=5)
    ap.add_argument("--learningrate", "-lr", type=float, default=0.0001)   # ref_expert.py:22
    ap.add_argument("--cutloss", "-cl", type=float, default=10)            # :37
    ap.add_argument("--check", action="store_true")
    opt = ap.parse_args()
    dev = torch.device("cuda")
    trainset = SyntheticRoomDataset(num_experts=1, length=max(opt.iterations, 1), seed=11, noise=0.05)
    trainset_loader = torch.utils.data.DataLoader(trainset, shuffle=False, num_workers=0)
    model = nn.Conv2d(3, 3, 1).to(dev)                                      # stand-in for Expert
    nn.init.eye_(model.weight.view(3, 3))
    nn.init.zeros_(model.bias)
    optimizer = torch.optim.Adam(model.parameters(), lr=opt.learningrate)   # :66
    out = []
    for iteration, (idx, image, focallength, gt_pose, gt_coords, gt_expert) in enumerate(trainset_loader):   # :95
        start_time = time.time()
        image = image.to(dev)
        padX, padY, image = random_shift(image, OUTPUT_SUBSAMPLE / 2)       # :100
        prediction = model(trainset.prediction_for(int(idx)).to(dev))       # :102 (the stand-in expert sees its prior)
        focallength = float(focallength[0])                                 # :114
        # :103-148 in one call; the principal point is the image centre (:118-119)
        robust_loss = reproj_loss(prediction, gt_pose.to(dev), focallength, padX, padY, opt.cutloss, OUTPUT_SUBSAMPLE,
                                  image.size(3) / 2, image.size(2) / 2)
        robust_loss.backward()                                              # :150
        loss_value = robust_loss.item()
        if opt.check:
            from oracle.reproj_loss_oracle import reproj_loss as original
            ref = original(prediction.detach().cpu(), gt_pose[0], focallength, padX, padY, opt.cutloss, OUTPUT_SUBSAMPLE,
                           image.size(3), image.size(2))
            assert abs(float(ref) - loss_value) < 1e-4 * max(1.0, float(ref)), (float(ref), loss_value)
        optimizer.step()                                                    # :151
        optimizer.zero_grad()                                               # :153
        print("Iteration: %6d, Loss: %.1f, Time: %.2fs" % (iteration, loss_value, time.time() - start_time), flush=True)  # :155
        out.append(loss_value)
    return out


if __name__ == "__main__":
    main()
