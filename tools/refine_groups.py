"""Times esac.forward's refinement stage for different CTA group sizes (GPU only): bench workload and native 60x80 shapes."""
import sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
import numpy as np, torch
import esac_b200.api as api
from esac_b200.synth import make_scene

ctx = api.context(); ctx.set_option("fixed_seed", 1)
cases = [("bench 7x256 480x640", dict(E=7, H=480, W=640, M=256, sub=1, seed=0, per_expert=True, active_only=False), (0, 32, 64, 96, 148)),
         ("native 60x80 E=7 M=256", dict(E=7, H=60, W=80, M=256, sub=8, seed=1), (0, 1, 2, 4, 8, 16, 32)),
         ("native 60x80 E=1 M=64", dict(E=1, H=60, W=80, M=64, sub=8, seed=2), (0, 1, 2, 4, 8, 16, 32)),
         ("120x160 E=4 M=256", dict(E=4, H=120, W=160, M=256, sub=4, seed=3), (0, 2, 4, 8, 16, 32, 64)),
         ("240x320 E=4 M=256", dict(E=4, H=240, W=320, M=256, sub=2, seed=4), (0, 8, 16, 32, 64, 96))]
for name, kw, groups in cases:
    sc = make_scene(**kw)
    coords = torch.from_numpy(sc.coords).cuda(); assign = torch.from_numpy(sc.assign).cuda(); out = torch.zeros(4, 4, device='cuda')
    for g in groups:
        ctx.set_option("refine_group", g)
        ts = []
        for _ in range(8):
            api.set_seed(3); api.forward(coords, assign, out, *sc.params); st = ctx.stats(); ts.append(st["ms_refine"])
        print(f"{name:24s} refine_group {g:3d} (used {st['refine_group']:3d}): refine {np.median(ts[2:]):.4f} ms rounds {st['refine_rounds']} total {st['ms_total']:.3f}", flush=True)
ctx.set_option("refine_group", 0)
