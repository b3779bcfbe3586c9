"""Times esac.forward's refinement stage for different CTA group sizes on the bench workload (GPU only)."""
import sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
import numpy as np, torch
import esac_b200.api as api
from esac_b200.synth import make_scene
sc = make_scene(E=7, H=480, W=640, M=256, sub=1, seed=0, per_expert=True, active_only=False)
coords = torch.from_numpy(sc.coords).cuda(); assign = torch.from_numpy(sc.assign).cuda(); out = torch.zeros(4, 4, device='cuda')
ctx = api.context(); ctx.set_option("fixed_seed", 1)
for g in (0, 16, 32, 64, 96, 148):
    ctx.set_option("refine_group", g)
    ts = []
    for _ in range(6):
        api.set_seed(3); api.forward(coords, assign, out, *sc.params); st = ctx.stats(); ts.append(st["ms_refine"])
    print(f"refine_group {g:3d} (used {st['refine_group']}): refine {np.median(ts[2:]):.4f} ms rounds {st['refine_rounds']} sample {st['ms_sample']:.3f} total {st['ms_total']:.3f}", flush=True)
