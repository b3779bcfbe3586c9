"""CUDA path against the committed golden fixtures (oracle outputs, tests/golden/make_golden.py)."""
from pathlib import Path

import numpy as np
import pytest

from esac_b200.synth import pose_error

pytestmark = pytest.mark.gpu
GOLD = sorted(p for p in (Path(__file__).resolve().parent / "golden").glob("*.npz") if not p.name.startswith("ref_"))


def _params(z):
    p = z["params"].tolist()
    return (int(p[0]), int(p[1])) + tuple(float(v) for v in p[2:9]) + (int(p[9]),)


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_forward_and_backward_match_golden(path):
    import esac_b200.api as api
    api.context().set_option("fixed_seed", 1)
    z = np.load(path)
    params = _params(z)
    api.set_seed(int(z["seed"]))
    out = np.zeros((4, 4), np.float32)
    e = api.forward(z["coords"], z["assign"], out, *params)
    hy = api.last_hypotheses()
    st = api.last_stats()
    assert hy["tries"].tolist() == z["tries"].tolist()
    assert np.abs(hy["scores"] - z["scores"]).max() < 1e-4
    assert e == int(z["expert"]) and st["winner"] == int(z["winner"]) and st["refine_rounds"] == int(z["rounds"])
    rot, trans = pose_error(out, z["pose"])
    assert rot < 1e-3 and trans < 1e-5
    api.set_seed(int(z["seed"]))
    g = np.zeros_like(z["coords"])
    loss = api.backward(z["coords"], g, z["assign"], z["gt_pose"], 1.0, 100.0, 100.0, *params)
    assert abs(loss - float(z["loss"])) < 1e-6 * max(1.0, abs(float(z["loss"])))
    scale = max(np.abs(z["grads"]).max(), 1e-12)
    assert np.abs(g - z["grads"]).max() / scale < 1e-3
