"""Multi-process driver of oracle/_ref (the reference's own compiled esac extension) -- BASELINE INFRASTRUCTURE ONLY.

The reference parallelises over hypotheses with `#pragma omp parallel for` (esac.cpp:131, esac_util.h:152).  In oracle/_ref the
OpenCV calls go through the cv2 module and therefore through the GIL, which would serialise those OpenMP threads -- an
artefact of the stand-in, not of the reference.  To time the reference's CPU path at the host's full parallelism, the
hypotheses are dealt to a fork pool instead: every worker runs the UNMODIFIED esac_forward, single-threaded, on its slice of
the hypothesis assignment (sampling, scoring, selection and refinement of its slice's winner), so the aggregate rate is what
an OpenMP build against native OpenCV would approach.  projectPoints without Jacobian uses the shim's native loop
(bit-identical to cv2's, tests/test_ref_pin.py): the Python binding always computes the 2N x 15 Jacobian the C++ reference
never asks for in esac_forward.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

_G: dict = {}


def _work(args):
    lo, hi, seed = args
    import torch
    R = _G["ref"]
    sc = _G["scene"]
    if not _G.get("quiet"):
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)  # the reference narrates its stages on stdout (esac.cpp:105-170)
        _G["quiet"] = True
    torch.set_num_threads(1)
    R.set_num_threads(1)
    R.set_native_project(True)
    R.force_init(int(seed) + lo)
    out = torch.zeros(4, 4)
    e = R.forward(_G["coords"], torch.from_numpy(np.ascontiguousarray(sc["assign"][lo:hi])), out, *sc["params"])
    return hi - lo, int(e)


def available() -> bool:
    from .build_ref import load_ref
    try:
        return load_ref() is not None
    except Exception:
        return False


def forward_ref_parallel(scene, take_idx, seed: int = 1305, workers: int | None = None, per_worker: int = 4):
    """scene: esac_b200.synth.Scene; take_idx: indices of the hypotheses of this sample.  Returns (hypotheses processed,
    seconds, workers used)."""
    import torch
    from .build_ref import load_ref
    workers = workers or os.cpu_count() or 1
    R = load_ref()
    assert R is not None, "oracle/_ref is not built"
    assign = np.ascontiguousarray(scene.assign[take_idx])
    M = len(assign)
    _G["ref"] = R
    _G["scene"] = dict(assign=assign, params=scene.params)
    _G["coords"] = torch.from_numpy(scene.coords)
    step = max(1, min(per_worker, (M + workers - 1) // workers))
    chunks = [(lo, min(M, lo + step), seed) for lo in range(0, M, step)]
    t0 = time.perf_counter()
    if workers > 1:
        ctx = mp.get_context("fork")
        with ctx.Pool(min(workers, len(chunks))) as pool:
            parts = pool.map(_work, chunks, chunksize=1)
    else:
        parts = [_work(c) for c in chunks]
    dt = time.perf_counter() - t0
    return sum(p[0] for p in parts), dt, min(workers, len(chunks))
