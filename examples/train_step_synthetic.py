"""The reference's end-to-end training step (code/train_esac.py:105-183) driven through this repository's `esac` module.

The reference's datasets and CNNs are out of scope (no data, no network access, scikit-image missing),
so this script supplies stand-ins with the same interfaces: a synthetic sample generator instead of RoomDataset, and a tiny
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
from esac_b200.synth import make_scene  # noqa: E402

OUTPUT_SUBSAMPLE = 8  # code/expert.py:13


class TinyExpert(nn.Module):
    """Stand-in for Expert: predicts scene coordinates [1,3,H/8,W/8]; here a learnable 1x1 conv on a coordinate prior."""

    def __init__(self, prior: torch.Tensor):
        super().__init__()
        self.register_buffer("prior", prior[None])
        self.adjust = nn.Conv2d(3, 3, 1)
        nn.init.eye_(self.adjust.weight.view(3, 3))
        nn.init.zeros_(self.adjust.bias)

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
    opt = ap.parse_args()
    dev = torch.device("cuda")
    sc = make_scene(E=opt.experts, H=60, W=80, M=opt.hypotheses, sub=OUTPUT_SUBSAMPLE, seed=3, active_only=False)
    coords = torch.from_numpy(sc.coords)
    experts = [TinyExpert(coords[e] + 0.01 * torch.randn_like(coords[e])).to(dev) for e in range(opt.experts)]
    gating = TinyGating(opt.experts).to(dev)
    opt_e = [torch.optim.Adam(m.parameters(), lr=1e-5) for m in experts]  # one optimiser per expert (expert_ensemble.py:9-37)
    opt_g = torch.optim.Adam(gating.parameters(), lr=1e-4)
    image = torch.rand(1, 1, 480, 640, device=dev)
    gt_pose = torch.from_numpy(sc.gt_pose)
    losses = []
    for it in range(opt.iterations):
        t0 = time.time()
        pred_w, pred_h = math.ceil(640 / OUTPUT_SUBSAMPLE), math.ceil(480 / OUTPUT_SUBSAMPLE)
        prediction = torch.zeros((opt.experts, 3, pred_h, pred_w), device=dev)                     # train_esac.py:121
        padX, padY = 0, 0                                                                        # util.random_shift
        gating_log_probs = gating(image)                                                         # :128
        gating_probs = torch.exp(gating_log_probs).cpu()
        e_hyps = torch.multinomial(gating_probs[0], opt.hypotheses, replacement=True)            # :138
        e_hyps_hist = torch.histc(e_hyps.float(), bins=opt.experts, min=0, max=opt.experts - 1)  # :141
        preds = []
        for e, count in enumerate(e_hyps_hist):                                                  # :143-145
            preds.append(experts[e](image)[0] if count > 0 else torch.zeros(3, pred_h, pred_w, device=dev))
        prediction = torch.stack(preds)
        prediction_gradients = torch.zeros_like(prediction)                                      # :148 (stays on the GPU)
        loss = esac.backward(prediction.detach(), prediction_gradients, e_hyps.to(dev), gt_pose,  # :151-168, no .cpu()
                             opt.weightrot, opt.weighttrans, opt.losscut, padX, padY, sc.f, sc.ppx, sc.ppy,
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
