"""Loads tests/golden/ref_*.npz (outputs of the compiled reference, tests/golden/make_ref_golden.py)."""
import ast
from pathlib import Path

import numpy as np

from esac_b200.synth import make_scene

GOLD_DIR = Path(__file__).resolve().parent / "golden"
REF_GOLD = sorted(GOLD_DIR.glob("ref_*.npz"))


def params_of(z):
    p = z["params"].tolist()
    return (int(p[0]), int(p[1])) + tuple(float(v) for v in p[2:9]) + (int(p[9]),)


def load(path):
    """Returns (z, coords): small fixtures carry the maps, 480x640 ones are regenerated from the stored make_scene arguments
    and verified by their checksum."""
    z = np.load(path)
    if "coords" in z.files:
        return z, z["coords"]
    kw = ast.literal_eval(str(z["scene_kw"]))
    sc = make_scene(**kw)
    assert np.array_equal(sc.assign, z["assign"])
    assert abs(sc.coords.astype(np.float64).sum() - float(z["coords_sum"])) <= 1e-9 * abs(float(z["coords_sum"])), \
        "make_scene no longer regenerates this fixture's maps"
    return z, sc.coords
