"""Multi-process driver of the cv2 oracle -- TEST / BASELINE INFRASTRUCTURE ONLY (see esac_oracle.py).

Mirrors the reference's `#pragma omp parallel for` over hypotheses (esac.cpp:131, esac_util.h:152) with a
fork pool: each worker samples and scores a slice of the hypotheses with single-threaded OpenCV; the
parent then does softMax / draw / refineHyp like esac_forward (esac.cpp:153-187).  Used by bench.py's
cpu_baseline and `--impl reference` legs to time the reference's CPU path on the box's host cores."""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import esac_oracle as O

_G = {}


def _work(args):
    lo, hi, seed = args
    import cv2
    cv2.setNumThreads(1)
    sc = _G["scene"]
    coords, assign = sc["coords"], sc["assign"]
    K = O.cam_mat(sc["f"], sc["ppx"], sc["ppy"])
    H, W = coords.shape[2:]
    sampling = O.create_sampling(W, H, sc["sub"], sc["shiftX"], sc["shiftY"])
    out = []
    for h in range(lo, hi):
        # sample_hypotheses addresses the stream by hypothesis index: run it on a one-element view
        e = int(assign[h])
        hyp = None
        for t in range(O.MAX_SAMPLING_TRIES):
            cells = O.draw_minimal_set(seed, h, t, W, H)
            img = np.array([sampling[y, x] for (x, y) in cells], np.float32)
            obj = np.array([O._cell_obj(coords, e, x, y) for (x, y) in cells], np.float32)
            ok, r, tv = O.safe_solve_pnp(obj, img, K, None, None, False, cv2.SOLVEPNP_P3P)
            hyp = (r, tv)
            if not ok:
                continue
            proj, _ = cv2.projectPoints(obj.reshape(-1, 1, 3), r, tv, K, None)
            d = img - proj.reshape(-1, 2).astype(np.float32)
            if np.all(np.sqrt(d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2) < sc["tau"]):
                break
        errs, _ = O.get_repro_errs(coords, hyp[0], hyp[1], e, sampling, K, sc["max_reproj"])
        score = O.get_hyp_scores([errs], sc["tau"], sc["alpha"], sc["beta"])[0]
        out.append((h, hyp[0], hyp[1], score))
    return out


def forward_parallel(scene: dict, seed: int = 1305, workers: int | None = None):
    """scene: dict(coords, assign, shiftX, shiftY, f, ppx, ppy, tau, alpha, beta, max_reproj, sub).
    Returns (winning expert, camera pose 4x4, seconds, workers used)."""
    workers = workers or os.cpu_count() or 1
    M = len(scene["assign"])
    _G["scene"] = scene
    t0 = time.perf_counter()
    chunks = [(lo, min(M, lo + max(1, (M + 4 * workers - 1) // (4 * workers))), seed)
              for lo in range(0, M, max(1, (M + 4 * workers - 1) // (4 * workers)))]
    if workers > 1:
        ctx = mp.get_context("fork")
        with ctx.Pool(workers) as pool:
            parts = pool.map(_work, chunks)
    else:
        parts = [_work(c) for c in chunks]
    res = sorted([r for p in parts for r in p], key=lambda r: r[0])
    scores = [r[3] for r in res]
    probs = O.softmax(scores)
    w = O.draw(probs, False)
    import cv2
    coords, assign = scene["coords"], scene["assign"]
    K = O.cam_mat(scene["f"], scene["ppx"], scene["ppy"])
    H, W = coords.shape[2:]
    sampling = O.create_sampling(W, H, scene["sub"], scene["shiftX"], scene["shiftY"])
    errs, _ = O.get_repro_errs(coords, res[w][1], res[w][2], int(assign[w]), sampling, K, scene["max_reproj"])
    r, t, _, _ = O.refine_hyp(coords, errs, sampling, K, int(assign[w]), scene["tau"], O.MAX_REF_STEPS,
                              scene["max_reproj"], res[w][1], res[w][2])
    T = O.pose2trans(r, t).astype(np.float32)
    return int(assign[w]), T, time.perf_counter() - t0, workers


def scene_dict(sc, take: int | None = None) -> dict:
    """esac_b200.synth.Scene -> the plain dict forward_parallel wants; `take` keeps the first hypotheses only."""
    assign = sc.assign if take is None else sc.assign[:take]
    return dict(coords=sc.coords, assign=assign, shiftX=sc.shiftX, shiftY=sc.shiftY, f=sc.f, ppx=sc.ppx, ppy=sc.ppy,
                tau=sc.tau, alpha=sc.alpha, beta=sc.beta, max_reproj=sc.max_reproj, sub=sc.sub)
