"""Sampling-stage diagnostics (GPU only): tries the waves pushed through the float prefilter against the tries the
sequential loop needs (sum of the per-hypothesis try counts), survivors judged by the fp64 path, waves, stage time --
for a few window policies.  Bench scene (7 experts x 256 hypotheses, 480x640)."""
import sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
import numpy as np, torch
import esac_b200.api as api
from esac_b200.synth import make_scene

ctx = api.context()
ctx.set_option("fixed_seed", 1)
scs = [make_scene(E=7, H=480, W=640, M=256, sub=1, seed=s, per_expert=True, active_only=False) for s in (0, 1)]
dev = [(torch.from_numpy(sc.coords).cuda(), torch.from_numpy(sc.assign).cuda(), sc) for sc in scs]
out = torch.zeros(4, 4, device="cuda")
# arguments: 0,span0,window,waves,lanes[,tail_boost]
combos = [(0, 128, 1.25, 7, 2)] + [tuple(float(v) for v in a.split(",")) for a in sys.argv[1:] if not a.startswith("--")]
if "--native" in sys.argv or "--trace" in sys.argv:
    combos = []
for combo in combos:
    mode, span0, window, waves, lanes = combo[:5]
    boost = combo[5] if len(combo) > 5 else 1.0
    ctx.set_option("sample_tail_boost", boost)
    ctx.set_option("sample_groups", lanes)
    ctx.set_option("sample_span0", span0); ctx.set_option("sample_window", window); ctx.set_option("sample_waves", waves)
    ms, tot, rows = [], [], []
    for rep in range(6):
        co, asg, sc = dev[rep % 2]
        ctx.set_seed(100 + rep)
        api.forward(co, asg, out, *sc.params)
        st = ctx.stats(); pr = ctx.sample_profile()
        if rep >= 2:
            need = int(ctx.hypotheses()["tries"].sum())
            ms.append(st["ms_sample"]); tot.append(st["ms_total"])
            rows.append((need, pr["tries_prefiltered"], pr["survivors_judged"], pr["waves"], pr["left_to_tail"]))
    r = np.array(rows, float).mean(0)
    print(f"lanes {int(lanes)} boost {boost} span0 {int(span0)} window {window} waves {int(waves)}: sample ms {np.mean(ms):.3f} (total {np.mean(tot):.3f})  needed {r[0]:.0f} "
          f"prefiltered {r[1]:.0f} (x{r[1] / r[0]:.2f}) survivors {r[2]:.0f} ({100 * r[2] / r[1]:.2f}%) waves/windows {r[3]:.1f} tail hyps {r[4]:.1f}", flush=True)

if "--native" in sys.argv:  # the reference's native shape: 60x80 cells (sub 8), 256 hypotheses in all
    nat = make_scene(E=7, H=60, W=80, M=256, sub=8, seed=0)
    co, asg = torch.from_numpy(nat.coords).cuda(), torch.from_numpy(nat.assign).cuda()
    for mode in (0,):
        ctx.set_option("sample_groups", 2)
        ctx.set_option("sample_span0", 128); ctx.set_option("sample_window", 1.25); ctx.set_option("sample_waves", 7)
        ms, tot = [], []
        for rep in range(8):
            ctx.set_seed(100 + rep)
            api.forward(co, asg, out, *nat.params)
            st = ctx.stats()
            if rep >= 2:
                ms.append(st["ms_sample"]); tot.append(st["ms_total"])
        print(f"native 60x80 M=256 : sample ms {np.mean(ms):.3f} total {np.mean(tot):.3f} launches {st['kernel_launches']}")

if "--trace" in sys.argv:  # timeline of the waves: who runs when
    for lanes, spread in ((1, 0), (2, 0)):
        ctx.set_option("sample_tail_boost", 1.0)
        ctx.set_option("sample_groups", lanes); ctx.set_option("sample_trace", 1)
        ctx.set_option("sample_span0", 128); ctx.set_option("sample_window", 1.25); ctx.set_option("sample_waves", 7)
        co, asg, sc = dev[0]
        for rep in range(3):
            ctx.set_seed(100 + rep)
            api.forward(co, asg, out, *sc.params)
        tr = ctx.sample_trace()
        print(f"--- trace, {lanes} lanes, (us from the first stamp; sample stage {ctx.stats()['ms_sample']:.3f} ms)")
        for g in range(lanes):
            for r in range(8):
                if tr[g, r, 0, 0] < 0:
                    continue
                p0, p1, e0, e1 = tr[g, r, 0, 0] / 1e3, tr[g, r, 0, 1] / 1e3, tr[g, r, 1, 0] / 1e3, tr[g, r, 1, 1] / 1e3
                print(f"lane {g} wave {r}: prefilter {p0:7.1f} -> {p1:7.1f} ({p1 - p0:5.1f})   exact {e0:7.1f} -> {e1:7.1f} ({e1 - e0:5.1f})")
        ctx.set_option("sample_trace", 0)
