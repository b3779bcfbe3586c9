"""The reference's end-to-end training step (code/train_esac.py:105-183) driven through this repository's `esac` module.

The reference's datasets and CNNs are out of scope (no data, no network access, scikit-image missing),
so this script supplies stand-ins with the same interfaces: esac_b200.compat.SyntheticRoomDataset instead of RoomDataset, and a tiny
gating CNN + per-expert 1x1-conv "experts" that start from the synthetic ground-truth coordinates plus noise instead of
ExpertEnsemble.  Everything from the gating draw to ensemble.update() follows the trainer line by line, and esac.backward is
called with the trainer's exact positional arguments -- on CUDA tensors, so the `.cpu()` / `.cuda()` copies of
train_esac.py:152,180 disappear.

    python examples/train_step_synthetic.py --iterations 5
"""
from __future__ import annotations

import argparse
import math
import sys
import time
from pathlib import Path

import torch
import torch.nn as nn

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import esac  # noqa: E402  (this repository's drop-in module)
import esac_b200.api as esac_api  # noqa: E402
from esac_b200.compat import OUTPUT_SUBSAMPLE, SyntheticRoomDataset, random_shift  # noqa: E402


class TinyExpert(nn.Module):
    """Stand-in for Expert: predicts scene coordinates [1,3,H/8,W/8]; here a learnable 1x1 conv on the coordinate prior the
    synthetic dataset attaches to the image (set with `see`)."""

    def __init__(self):
        super().__init__()
        self.adjust = nn.Conv2d(3, 3, 1)
        nn.init.eye_(self.adjust.weight.view(3, 3))
        nn.init.zeros_(self.adjust.bias)
        self.prior = None

    def see(self, prior: torch.Tensor):
        self.prior = prior[None]

    def forward(self, image):
        return self.adjust(self.prior)


class TinyGating(nn.Module):
    """Stand-in for Gating (code/gating.py:19-58): image -> log-probabilities over experts."""

    def __init__(self, n):
        super().__init__()
        self.conv = nn.Conv2d(1, 8, 3, 2, 1)
        self.fc = nn.Linear(8, n)

    def forward(self, image):
        x = torch.relu(self.conv(image)).mean(dim=(2, 3))
        return torch.log_softmax(self.fc(x), dim=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=5)
    ap.add_argument("--experts", type=int, default=4)
    ap.add_argument("--hypotheses", "-hyps", type=int, default=256)     # train_esac.py:29
    ap.add_argument("--threshold", type=float, default=10)              # :32
    ap.add_argument("--inlieralpha", type=float, default=100)           # :35
    ap.add_argument("--inlierbeta", type=float, default=0.5)            # :38
    ap.add_argument("--maxreprojection", type=float, default=100)       # :41
    ap.add_argument("--weightrot", type=float, default=1.0)
    ap.add_argument("--weighttrans", type=float, default=100.0)
    ap.add_argument("--losscut", type=float, default=100.0)
    ap.add_argument("--device-assignment", action="store_true", help="draw e_hyps on the GPU (esacb200_assign_hypotheses)")
    opt = ap.parse_args()
    dev = torch.device("cuda")
    trainset = SyntheticRoomDataset(num_experts=opt.experts, length=max(opt.iterations, 1), hypotheses=opt.hypotheses, seed=3)
    trainset_loader = torch.utils.data.DataLoader(trainset, shuffle=False, num_workers=0)        # train_esac.py:77 (batch 1)
    experts = [TinyExpert().to(dev) for _ in range(opt.experts)]
    gating = TinyGating(opt.experts).to(dev)
    opt_e = [torch.optim.Adam(m.parameters(), lr=1e-5) for m in experts]  # one optimiser per expert (expert_ensemble.py:9-37)
    opt_g = torch.optim.Adam(gating.parameters(), lr=1e-4)
    losses = []
    for it, (idx, image, focallength, gt_pose, gt_coords, gt_expert) in enumerate(trainset_loader):  # :96
        t0 = time.time()
        pp_x = float(image.size(3) / 2)                                                          # :110-112
        pp_y = float(image.size(2) / 2)
        focallength = float(focallength[0])
        gt_pose = gt_pose[0]                                                                     # :114
        pred_w, pred_h = math.ceil(image.size(3) / OUTPUT_SUBSAMPLE), math.ceil(image.size(2) / OUTPUT_SUBSAMPLE)  # :117-118
        image = image.to(dev)
        padX, padY, image = random_shift(image, OUTPUT_SUBSAMPLE / 2)                            # :125 (compat: int bound)
        prior = trainset.prediction_for(int(idx)).to(dev)
        for e in range(opt.experts):
            experts[e].see(prior[e] + 0.01 * torch.randn_like(prior[e]))
        gating_log_probs = gating(image)                                                         # :128
        gating_probs = torch.exp(gating_log_probs)
        if opt.device_assignment:
            # clamp_probs + multinomial + histc in one kernel, no .cpu() (esacb200_assign_hypotheses)
            e_hyps, e_hyps_hist = esac_api.assign_hypotheses(gating_probs.detach(), opt.hypotheses, seed=1305 + it)
            e_hyps, e_hyps_hist = e_hyps[0], e_hyps_hist[0].cpu()
        else:
            gating_probs = gating_probs.cpu()
            e_hyps = torch.multinomial(gating_probs[0], opt.hypotheses, replacement=True)            # :138
            e_hyps_hist = torch.histc(e_hyps.float(), bins=opt.experts, min=0, max=opt.experts - 1)  # :141
        preds = []
        for e, count in enumerate(e_hyps_hist):                                                  # :143-145
            preds.append(experts[e](image)[0] if count > 0 else torch.zeros(3, pred_h, pred_w, device=dev))
        prediction = torch.stack(preds)
        prediction_gradients = torch.zeros_like(prediction)                                      # :148 (stays on the GPU)
        loss = esac.backward(prediction.detach(), prediction_gradients, e_hyps.to(dev), gt_pose,  # :151-168, no .cpu()
                             opt.weightrot, opt.weighttrans, opt.losscut, padX, padY, focallength, pp_x, pp_y,
                             opt.threshold, opt.inlieralpha, opt.inlierbeta, opt.maxreprojection, OUTPUT_SUBSAMPLE)
        gating_log_prob_gradients = (loss * e_hyps_hist).unsqueeze(0)                            # :171-176
        for o in opt_e + [opt_g]:
            o.zero_grad()
        torch.autograd.backward((prediction, gating_log_probs), (prediction_gradients, gating_log_prob_gradients.to(dev)))  # :178-180
        for e, count in enumerate(e_hyps_hist):                                                  # ensemble.update (expert_ensemble.py:58-68)
            if count > 0:
                opt_e[e].step()
        opt_g.step()
        losses.append(loss)
        print("Iteration: %6d, Loss: %.2f, Time: %.2fs" % (it, loss, time.time() - t0), flush=True)  # :185
    return losses


if __name__ == "__main__":
    main()
