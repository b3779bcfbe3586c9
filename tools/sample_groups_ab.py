"""A/B of the sampling stage: one lane (sample_groups=1) against two interleaved lanes on two streams (sample_groups=2).
Bench workload (7 experts x 256 hypotheses, 480x640) and the native 60x80 shape.  Prints the median stage times."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import esac_b200.api as api  # noqa: E402
from esac_b200.synth import make_scene  # noqa: E402


def run(sc, groups, reps=30):
    ctx = api.context()
    ctx.set_option("sample_groups", groups)
    coords = torch.from_numpy(sc.coords).cuda()
    assign = torch.from_numpy(sc.assign).cuda()
    out = torch.zeros(4, 4, device="cuda")
    rows = []
    for i in range(reps):
        api.forward(coords, assign, out, *sc.params)
        st = ctx.stats()
        rows.append((st["ms_sample"], st["ms_score"], st["ms_refine"], st["ms_total"]))
    ctx.set_option("sample_groups", 2)
    return np.median(np.array(rows[5:]), axis=0)


def main():
    for name, kw in [("bench 7x256 480x640", dict(E=7, H=480, W=640, M=256, sub=1, per_expert=True, seed=1)),
                     ("bench scene 2", dict(E=7, H=480, W=640, M=256, sub=1, per_expert=True, seed=2)),
                     ("native E=7 M=256 60x80", dict(E=7, H=60, W=80, M=256, sub=8, seed=1)),
                     ("native E=20 M=1024 60x80", dict(E=20, H=60, W=80, M=1024, sub=8, seed=1))]:
        sc = make_scene(**kw)
        for g in (1, 2, 1, 2):
            r = run(sc, g)
            print(f"{name:28s} groups={g}: sample {r[0]:.3f} ms  score {r[1]:.3f}  refine {r[2]:.3f}  total {r[3]:.3f}", flush=True)


if __name__ == "__main__":
    main()
