"""Device-level check of the float prefilter: sampling with the prefilter on must pick exactly the tries it picks with the
prefilter off (every try through the fp64 path), over many seeds and scene types.  Prints hypotheses compared / mismatches."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import esac_b200.api as api  # noqa: E402
from esac_b200.synth import make_scene  # noqa: E402

ctx = api.context()
ctx.set_option("fixed_seed", 1)
ctx.set_option("max_tries", 20000)
total = bad = 0
cases = [dict(E=2, H=60, W=80, M=512, sub=8), dict(E=2, H=60, W=80, M=512, sub=8, outdoor=True),
         dict(E=2, H=60, W=80, M=512, sub=8, world_offset=700.0), dict(E=3, H=120, W=160, M=512, sub=4, noise=0.1),
         dict(E=2, H=24, W=32, M=512, sub=8), dict(E=2, H=107, W=60, M=512, sub=8, outdoor=True, outlier_frac=0.7),
         dict(E=4, H=240, W=320, M=1024, sub=2, per_expert=False)]
for ci, kw in enumerate(cases):
    for seed in range(12):
        sc = make_scene(seed=1000 * ci + seed, active_only=False, **kw)
        coords = torch.from_numpy(sc.coords).cuda()
        assign = torch.from_numpy(sc.assign).cuda()
        out = torch.zeros(4, 4, device="cuda")
        res = []
        for pf in (1, 0):
            ctx.set_option("sample_prefilter", pf)
            api.set_seed(4242 + seed)
            api.forward(coords, assign, out, *sc.params)
            hy = api.last_hypotheses()
            res.append((hy["tries"].copy(), hy["cells"].copy()))
        ctx.set_option("sample_prefilter", 1)
        m = (res[0][0] != res[1][0]) | (res[0][1] != res[1][1]).reshape(len(res[0][0]), -1).any(axis=1)
        total += len(m)
        bad += int(m.sum())
    print(f"case {ci} {kw}: cumulative {total} hypotheses, {bad} mismatches", flush=True)
print("RESULT", total, bad)
sys.exit(1 if bad else 0)
