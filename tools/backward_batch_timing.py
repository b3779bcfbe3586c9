"""backward_batch (worker streams) against a Python loop of esac.backward, native 60x80 maps (train_esac.py's shape)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import esac_b200.api as api  # noqa: E402
from esac_b200.synth import make_scene  # noqa: E402


def main():
    ctx = api.context()
    for (E, M, B) in [(1, 64, 8), (7, 256, 8), (20, 1024, 8), (20, 1024, 32)]:
        scenes = [make_scene(E=E, H=60, W=80, M=M, sub=8, seed=100 + b) for b in range(B)]
        coords = torch.from_numpy(np.stack([s.coords for s in scenes])).cuda()
        assign = torch.from_numpy(np.stack([s.assign for s in scenes])).cuda()
        gts = torch.from_numpy(np.stack([s.gt_pose for s in scenes])).cuda()
        p = scenes[0].params
        grads = torch.zeros_like(coords)

        def loop():
            for b in range(B):
                api.backward(coords[b], grads[b], assign[b], gts[b], 1.0, 100.0, 100.0, *p)

        def batch():
            api.backward_batch(coords, grads, assign, gts, 1.0, 100.0, 100.0, 0, 0, *p[2:])

        res = {}
        for name, fn, workers in [("loop", loop, 0), ("batch w=1", batch, 1), ("batch w=2", batch, 2), ("batch w=4", batch, 4),
                                  ("batch w=8", batch, 8)]:
            if workers:
                ctx.set_option("batch_workers", workers)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            res[name] = (time.perf_counter() - t0) / 5 / B * 1e3
        print(f"E={E} M={M} B={B} 60x80: ms per image  " + "  ".join(f"{k} {v:.3f}" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
