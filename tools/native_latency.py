"""Latency of esac.forward at the reference's native shapes (60x80 map from a 480x640 image) on CUDA tensors (GPU only)."""
import sys, time
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
import numpy as np, torch
import esac_b200.api as api
from esac_b200.synth import make_scene
ctx = api.context()
for name, kw in [("config 1: E=1 M=64 60x80", dict(E=1, H=60, W=80, M=64, sub=8, seed=1)),
                 ("config 2 native: E=7 M=256 60x80", dict(E=7, H=60, W=80, M=256, sub=8, seed=2)),
                 ("E=19 M=256 60x80", dict(E=19, H=60, W=80, M=256, sub=8, seed=3)),
                 ("E=20 M=1024 60x80", dict(E=20, H=60, W=80, M=1024, sub=8, seed=4))]:
    sc = make_scene(**kw)
    coords = torch.from_numpy(sc.coords).cuda(); assign = torch.from_numpy(sc.assign).cuda(); out = torch.zeros(4, 4, device='cuda')
    hc = torch.from_numpy(sc.coords).pin_memory(); ha = torch.from_numpy(sc.assign); ho = torch.zeros(4, 4)
    for host in (False, True):
        ts, wall = [], []
        for i in range(30):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            if host: api.forward(hc, ha, ho, *sc.params)
            else: api.forward(coords, assign, out, *sc.params)
            wall.append(time.perf_counter() - t0); st = ctx.stats(); ts.append(st["ms_total"])
        print(f"{name} [{'host' if host else 'cuda'} tensors]: wall {1e3*np.median(wall[5:]):.3f} ms, device {np.median(ts[5:]):.3f} ms "
              f"(sample {st['ms_sample']:.3f} score {st['ms_score']:.3f} select {st['ms_select']:.3f} refine {st['ms_refine']:.3f}) "
              f"launches {st['kernel_launches']} -> {kw['M']/np.median(wall[5:]):.0f} hyp/s", flush=True)
