"""HBM roofline of the fused reprojection-loss kernel (esac_b200/csrc/reproj.cu): 12 B read + 12 B written per cell.
Prints achieved GB/s (CUDA-event time of the kernel alone, from the library's own stage timers) for a few batch shapes
next to a plain torch implementation of the same loss + autograd (what ref_expert.py:103-150 runs)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import esac_b200.api as api  # noqa: E402
from esac_b200.synth import make_scene  # noqa: E402


def torch_ref(pred, pose_inv, cam, grid, cut):
    """The original op sequence on CUDA tensors (batch of 1, as the reference trains)."""
    p = pred.detach().clone().requires_grad_(True)
    ones = torch.ones((1, 1, p.size(2), p.size(3)), device=p.device)
    x = torch.cat((p, ones), 1)[0].view(4, -1)
    px = torch.mm(cam, torch.mm(pose_inv, x))
    px[2].clamp_(min=0.1)
    px = px[0:2] / px[2]
    px = (px - grid).norm(2, 0).clamp(0, 100)
    l1 = px[px <= cut]
    sq = torch.sqrt(cut * px[px > cut])
    loss = (l1.sum() + sq.sum()) / float(px.size(0))
    loss.backward()
    return loss, p.grad


def main():
    peak = None
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print(f"HBM peak (MEASURED_PEAKS.json): {peak} GB/s")
    ctx = api.context()
    for (B, H, W, sub) in [(1, 60, 80, 8), (8, 60, 80, 8), (1, 480, 640, 1), (8, 480, 640, 1), (64, 480, 640, 1), (256, 480, 640, 1)]:
        sc = make_scene(E=1, H=H, W=W, M=8, sub=sub, seed=5)
        pred = torch.from_numpy(sc.coords[0]).cuda().unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
        pred += 0.01 * torch.randn_like(pred)
        gts = torch.from_numpy(sc.gt_pose).unsqueeze(0).repeat(B, 1, 1).contiguous()
        grads = torch.empty_like(pred)
        ms = []
        wall = []
        for it in range(30):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            api.reproj_loss(pred, gts, 525.0, 1, -1, 10.0, sub, outGradients=grads)
            wall.append((time.perf_counter() - t0) * 1e3)
            ms.append(ctx.stats()["ms_score"])
        ms = sorted(ms[5:])[len(ms[5:]) // 2]
        wl = sorted(wall[5:])[len(wall[5:]) // 2]
        nbytes = B * H * W * 24
        line = f"B={B:4d} {H}x{W}: kernel {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.1f} GB/s"
        if peak:
            line += f" ({nbytes/ms/1e6/peak:.2f} of peak)"
        line += f"  call wall {wl*1e3:8.1f} us"
        # torch baseline on one image
        pose_inv = torch.from_numpy(sc.gt_pose).inverse()[0:3, :].cuda()
        cam = torch.eye(3)
        cam[0, 0] = cam[1, 1] = 525.0
        cam[0, 2], cam[1, 2] = W * sub / 2, H * sub / 2
        cam = cam.cuda()
        ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        grid = torch.stack((xs * sub + sub / 2 - 1.0, ys * sub + sub / 2 + 1.0)).float().view(2, -1).cuda()
        for _ in range(3):
            torch_ref(pred[:1], pose_inv, cam, grid, 10.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            torch_ref(pred[:1], pose_inv, cam, grid, 10.0)
        torch.cuda.synchronize()
        tt = (time.perf_counter() - t0) / 10 * 1e3
        line += f"  | torch ops + autograd, 1 image: {tt*1e3:8.1f} us"
        print(line, flush=True)


if __name__ == "__main__":
    main()
