"""Times esac.backward (config-5-like shapes) on CUDA tensors and prints the stage breakdown (GPU only)."""
import sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
import numpy as np, torch
import esac_b200.api as api
from esac_b200.synth import make_scene
ctx = api.context()
for name, kw in [("native 60x80 E=20 M=1024", dict(E=20, H=60, W=80, M=1024, sub=8, seed=1)),
                 ("full-res 480x640 E=7 M=256", dict(E=7, H=480, W=640, M=256, sub=1, seed=2)),
                 ("full-res 480x640 E=20 M=1024 (config 5)", dict(E=20, H=480, W=640, M=1024, sub=1, seed=3))]:
    sc = make_scene(**kw)
    coords = torch.from_numpy(sc.coords).cuda(); assign = torch.from_numpy(sc.assign).cuda()
    gt = torch.from_numpy(sc.gt_pose)
    grads = torch.zeros_like(coords)
    ts = []
    for i in range(4):
        grads.zero_()
        loss = api.backward(coords, grads, assign, gt, 1.0, 100.0, 100.0, *sc.params)
        st = ctx.stats(); ts.append(st["ms_total"])
    print(f"{name}: total {np.median(ts[1:]):.3f} ms  loss {loss:.4f}  n_contrib {st['n_contrib']}  stages "
          f"sample {st['ms_sample']:.3f} score {st['ms_score']:.3f} refine {st['ms_refine']:.3f} backward {st['ms_backward']:.3f}  "
          f"group {st['refine_group']}  grad absmax {float(grads.abs().max()):.4g}", flush=True)
