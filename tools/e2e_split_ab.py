"""A/B of the split upload (esac.forward on pinned host tensors, bench workload): upload_split 0 / 1 interleaved on one box,
next to the raw pinned H2D rate of the same 25.8 MB."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import esac  # noqa: E402
import esac_b200.api as api  # noqa: E402
from esac_b200.synth import make_scene  # noqa: E402

ctx = api.context()
scenes = [make_scene(E=7, H=480, W=640, M=256, sub=1, seed=i, per_expert=True, active_only=False) for i in range(4)]
hc = [torch.from_numpy(s.coords).pin_memory() for s in scenes]
ha = [torch.from_numpy(s.assign).pin_memory() for s in scenes]
out = torch.zeros(4, 4).pin_memory()
dev = torch.empty_like(hc[0], device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    dev.copy_(hc[i % 4], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print(f"raw pinned H2D: {hc[0].numel() * 4 / dt / 1e9:.1f} GB/s ({dt * 1e3:.3f} ms per image)")
for rep in range(3):
    for split in (0, 1):
        ctx.set_option("upload_split", split)
        for i in range(5):
            esac.forward(hc[i % 4], ha[i % 4], out, *scenes[0].params)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(40):
            esac.forward(hc[i % 4], ha[i % 4], out, *scenes[0].params)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 40
        print(f"upload_split={split}: {dt * 1e3:.3f} ms per forward  ({1792 / dt / 1e6:.3f} M hyp/s)", flush=True)
ctx.set_option("upload_split", 1)
