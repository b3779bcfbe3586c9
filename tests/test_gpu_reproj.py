"""GPU parity of the fused reprojection loss (ref_expert.py:103-150) against the torch restatement (run with `-m gpu`).

Tolerance: the original computes in float32, and float32 itself sits ~1e-4 (relative to the largest gradient entry) away
from the exact value: pixel coordinates of ~1e3 carry ~6e-5 px of rounding while inlier errors are ~1 px, so the unit
vector (px - target)/err is ill-conditioned in exactly the cells that matter.  The yardstick is therefore the float64
evaluation of the same op sequence, and the bar is "at least as accurate as torch's own float32 evaluation": RMS
deviation from float64 no larger than torch-float32's (1.5x slack), worst cell within 4x of torch-float32's worst
cell, loss within 1e-5 relative.  The loss has two kinks -- at err = cutloss the gradient halves, at err = 100 px it
drops to zero -- and a cell whose error lies within float32 rounding of a kink may land on either side in ANY float32
evaluation; cells within 1e-3 px of a kink (by the float64 evaluation) are therefore left out of the gradient comparison
(and counted: they must stay a handful)."""
import numpy as np
import pytest

from esac_b200.synth import make_scene

pytestmark = pytest.mark.gpu


def _case(B, H, W, seed, **kw):
    scenes = [make_scene(E=1, H=H, W=W, M=8, sub=8, seed=seed + b, **kw) for b in range(B)]
    pred = np.stack([s.coords[0] for s in scenes])
    gts = np.stack([s.gt_pose for s in scenes])
    return pred, gts


@pytest.mark.parametrize("B,H,W,kw", [(1, 60, 80, {}), (3, 60, 80, {"outlier_frac": 0.6}), (2, 33, 47, {}),     # odd size: scalar path
                                      (2, 80, 60, {"noise": 0.5}), (1, 480, 640, {})])
@pytest.mark.parametrize("kind", ["cpu", "cuda"])
def test_reproj_loss_and_gradient_match_torch(B, H, W, kw, kind):
    import torch
    import esac_b200.api as api
    from oracle.reproj_loss_oracle import reproj_errors, reproj_loss_and_grad
    pred, gts = _case(B, H, W, 300 + H, **kw)
    pred[0, :, 0, 0] = [0.0, 0.0, -50.0]       # behind the camera -> depth clamp
    pred[0, :, 1, 1] = [1e4, -1e4, 3.0]        # error far beyond 100 px -> zero gradient
    padx = [2, -3, 0][:B]
    pady = [-1, 4, 0][:B]
    f, cut = 525.0, 10.0
    dev = "cuda" if kind == "cuda" else "cpu"
    tp = torch.from_numpy(pred).to(dev)
    tg = torch.full(pred.shape, 7.0, device=dev)      # overwritten, not accumulated
    losses = api.reproj_loss(tp, torch.from_numpy(gts).to(dev), f, padx, pady, cut, 8, outGradients=tg)
    g = tg.cpu().numpy()
    for b in range(B):
        l32, g32 = reproj_loss_and_grad(pred[b], gts[b], f, padx[b], pady[b], cut)
        l64, g64 = reproj_loss_and_grad(pred[b], gts[b], f, padx[b], pady[b], cut, dtype=torch.float64)
        assert abs(losses[b] - l64) <= 1e-5 * max(1.0, abs(l64)), (losses[b], l32, l64)
        e64 = reproj_errors(torch.from_numpy(pred[b]), torch.from_numpy(gts[b]), f, padx[b], pady[b], dtype=torch.float64)
        e64 = e64.numpy().reshape(H, W)
        kink = (np.abs(e64 - cut) < 1e-3) | (np.abs(e64 - 100.0) < 1e-3) & (e64 < 100.0)
        assert kink.sum() <= 2e-4 * H * W + 2
        keep = ~kink[None]
        d32 = (g32.double() - g64).numpy() * keep
        dk = (g[b] - g64.numpy()) * keep
        scale = g64.abs().max().item()
        assert np.sqrt((dk ** 2).mean()) <= 1.5 * np.sqrt((d32 ** 2).mean()) + 1e-7 * scale, (np.sqrt((dk ** 2).mean()), np.sqrt((d32 ** 2).mean()))
        assert np.abs(dk).max() <= 4 * np.abs(d32).max() + 1e-6 * scale, (np.abs(dk).max(), np.abs(d32).max(), scale)
    # loss-only call leaves no gradient behind and returns the same numbers
    again = api.reproj_loss(tp, torch.from_numpy(gts).to(dev), f, padx, pady, cut, 8)
    assert again == losses


def test_reproj_loss_autograd_node_trains_like_the_original():
    import torch
    from esac_b200.autograd import reproj_loss
    from oracle.reproj_loss_oracle import reproj_loss as ref_loss
    pred, gts = _case(2, 24, 32, 900)
    p = torch.from_numpy(pred).cuda().requires_grad_(True)
    loss = reproj_loss(p, torch.from_numpy(gts).cuda(), 525.0, [1, -2], [0, 3], 10.0)
    (loss * 3.0).backward()
    q = torch.from_numpy(pred).double().requires_grad_(True)
    ref = sum(ref_loss(q[b], torch.from_numpy(gts[b]), 525.0, [1, -2][b], [0, 3][b], 10.0, dtype=torch.float64) for b in range(2)) / 2
    (ref * 3.0).backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    assert (p.grad.cpu().double() - q.grad).abs().max().item() < 5e-4 * q.grad.abs().max().item()


def test_reproj_loss_rejects_bad_input():
    import torch
    import esac_b200.api as api
    with pytest.raises(RuntimeError):
        api.reproj_loss(torch.zeros(1, 3, 8, 8, dtype=torch.float64), torch.eye(4).unsqueeze(0), 525., 0, 0, 10.)
    with pytest.raises(RuntimeError):
        api.reproj_loss(torch.zeros(2, 3, 8, 8), torch.eye(4).unsqueeze(0), 525., 0, 0, 10.)
    with pytest.raises(RuntimeError, match="singular"):
        api.reproj_loss(torch.zeros(1, 3, 8, 8), torch.zeros(1, 4, 4), 525., 0, 0, 10.)
