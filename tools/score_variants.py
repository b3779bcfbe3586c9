"""Times the scoring kernel for different launch shapes on the bench workload (GPU only)."""
import sys, json
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
import numpy as np, torch
import esac_b200.api as api
from esac_b200.synth import make_scene
sc = make_scene(E=7, H=480, W=640, M=256, sub=1, seed=0, per_expert=True, active_only=False)
coords = torch.from_numpy(sc.coords).cuda(); assign = torch.from_numpy(sc.assign).cuda()
rng = np.random.default_rng(0)
poses = np.zeros((len(sc.assign), 6)); poses[:, :3] = rng.normal(0, 0.3, (len(sc.assign), 3)); poses[:, 3:] = rng.normal(0, 2, (len(sc.assign), 3))
ctx = api.context()
for ppt in (8, 4, 2):
    for hc in (64, 32, 16):
        ctx.set_option("score_ppt", ppt); ctx.set_option("score_hc", hc)
        ts = []
        for _ in range(6):
            api.score_poses(coords, assign, poses, *sc.params)
            ts.append(ctx.stats()["ms_score"])
        print(f"ppt {ppt} hc {hc:2d}: score kernel {np.median(ts[2:]):.4f} ms  (grid {ctx.stats()['score_grid']})", flush=True)
