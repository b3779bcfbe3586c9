"""Refinement stage timing (GPU only): forward winner refinement and the backward's refine-all at the bench shape, for the
fp64 / mixed Jacobian variants and a few job-group shapes.  Prints stage times from the library's CUDA-event timers."""
import sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
import numpy as np, torch
import esac_b200.api as api
from esac_b200.synth import make_scene, pose_error

ctx = api.context()
ctx.set_option("fixed_seed", 1)
full = make_scene(E=7, H=480, W=640, M=256, sub=1, seed=0, per_expert=True, active_only=False)
nat = make_scene(E=7, H=60, W=80, M=256, sub=8, seed=0)
c5 = make_scene(E=20, H=480, W=640, M=1024, sub=1, seed=3)
for name, sc in (("full 7x256 480x640", full), ("native 60x80 M=256", nat)):
    coords = torch.from_numpy(sc.coords).cuda(); assign = torch.from_numpy(sc.assign).cuda()
    out = torch.zeros(4, 4, device="cuda")
    ref = None
    for mixed in (0, 1):  # inlier compaction off / on
        ctx.set_option("refine_compact", mixed)
        ts, rs = [], []
        for i in range(6):
            ctx.set_seed(100 + i)
            api.forward(coords, assign, out, *sc.params)
            st = ctx.stats(); ts.append(st["ms_refine"]); rs.append(st["refine_rounds"])
        ctx.set_seed(100)
        api.forward(coords, assign, out, *sc.params)
        p = out.cpu().numpy().copy()
        if ref is None:
            ref = p
        d = pose_error(p, ref)
        print(f"fwd {name}: compact={mixed} refine ms {np.round(ts, 3).tolist()} rounds {rs} group {st['refine_group']} "
              f"total {st['ms_total']:.3f} pose diff vs uncompacted {d[0]:.2e} deg {d[1]:.2e} m", flush=True)
for name, sc in (("full 7x256 480x640", full), ("c5 20E M=1024 480x640", c5), ("native 60x80 M=256", nat)):
    coords = torch.from_numpy(sc.coords).cuda(); assign = torch.from_numpy(sc.assign).cuda()
    gt = torch.from_numpy(sc.gt_pose)
    grads = torch.zeros_like(coords)
    base = None
    for mixed, jpg, grp in ((0, 3, 0), (1, 3, 0), (1, 2, 0), (1, 4, 0), (1, 6, 0), (1, 3, 16), (1, 3, 32)):
        ctx.set_option("refine_compact", mixed)
        ctx.set_option("refine_jobs_per_group", jpg); ctx.set_option("refine_group", grp)
        ts = []
        for i in range(4):
            grads.zero_()
            ctx.set_seed(7)
            loss = api.backward(coords, grads, assign, gt, 1.0, 100.0, 100.0, *sc.params)
            st = ctx.stats(); ts.append(st["ms_refine"])
        g = grads.cpu().numpy().copy()
        if base is None:
            base = (loss, g)
        dg = np.abs(g - base[1]).max() / max(np.abs(base[1]).max(), 1e-30)
        print(f"bwd {name}: compact={mixed} jobs/group={jpg} group_opt={grp} -> group {st['refine_group']} contrib {st['n_contrib']} "
              f"refine ms {np.round(ts[1:], 3).tolist()} total {st['ms_total']:.3f} loss diff {abs(loss - base[0]):.2e} grad rel diff {dg:.2e}", flush=True)
    ctx.set_option("refine_group", 0); ctx.set_option("refine_jobs_per_group", 3)
