"""A few esac.forward calls at the bench shape (GPU only) -- the target of ncu captures of single kernels."""
import sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
import torch
import esac_b200.api as api
from esac_b200.synth import make_scene

ctx = api.context()
ctx.set_option("fixed_seed", 1)
sc = make_scene(E=7, H=480, W=640, M=256, sub=1, seed=0, per_expert=True, active_only=False)
coords = torch.from_numpy(sc.coords).cuda(); assign = torch.from_numpy(sc.assign).cuda()
out = torch.zeros(4, 4, device="cuda")
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ctx.set_seed(100)
    e = api.forward(coords, assign, out, *sc.params)
    st = ctx.stats()
    print(f"expert {e} rounds {st['refine_rounds']} stages ms: sample {st['ms_sample']:.3f} score {st['ms_score']:.3f} refine {st['ms_refine']:.3f} total {st['ms_total']:.3f}")
