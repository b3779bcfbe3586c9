"""Writes the golden fixtures tests/golden/*.npz from the cv2 oracle (oracle/esac_oracle.py).

The reference has no golden vectors and cannot be built here, so these are ORACLE outputs (parity unpinned
against the compiled reference).  They serve two purposes: the CPU suite checks the
oracle still reproduces them, and the GPU suite (tests/test_gpu_golden.py) checks the CUDA path against
them without re-running the oracle.  Run:  python tests/golden/make_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from esac_b200.synth import make_scene  # noqa: E402
from oracle import esac_oracle as O  # noqa: E402

CASES = {
    "c1_single_expert_60x80": dict(E=1, H=60, W=80, M=64, sub=8, seed=31),                 # BASELINE configs[0] shape
    "ensemble3_30x40_shift": dict(E=3, H=30, W=40, M=48, sub=8, seed=32, shiftX=2, shiftY=-3),
    "portrait_40x27": dict(E=2, H=40, W=27, M=32, sub=8, seed=33),
    "world_scale_24x32": dict(E=2, H=24, W=32, M=32, sub=8, seed=34, world_offset=700.0),
}

if __name__ == "__main__":
    here = Path(__file__).resolve().parent
    for name, kw in CASES.items():
        sc = make_scene(**kw)
        seed = 1000 + kw["seed"]
        out = np.zeros((4, 4), np.float32)
        e, tr = O.forward(sc.coords, sc.assign, out, *sc.params, seed=seed, trace=True)
        g = np.zeros_like(sc.coords)
        loss, bt = O.backward(sc.coords, g, sc.assign, sc.gt_pose, 1.0, 100.0, 100.0, *sc.params, seed=seed, trace=True)
        np.savez_compressed(here / f"{name}.npz", coords=sc.coords, assign=sc.assign, gt_pose=sc.gt_pose,
                            params=np.array(sc.params, np.float64), seed=seed, expert=e, pose=out,
                            scores=np.array(tr.scores), winner=tr.winner, rounds=tr.rounds,
                            ref_pose6=np.concatenate([tr.ref_rvec.ravel(), tr.ref_tvec.ravel()]),
                            tries=np.array([h.tries for h in tr.hyps]), loss=loss, grads=g,
                            losses=np.array(bt.losses))
        print(name, "expert", e, "winner", tr.winner, "score", tr.scores[tr.winner], "loss", loss)
