"""CPU checks of the geometry the CUDA kernels run (esac_b200/csrc/esac_geom.cuh, esac_rng.cuh), exercised
through the host test hooks of libesac_b200.so against OpenCV (cv2) and the oracle.  No GPU needed."""
import ctypes as C

import cv2
import numpy as np
import pytest

from oracle import esac_oracle as O

K = np.array([[525, 0, 320], [0, 525, 240], [0, 0, 1]], np.float32)


def _ptr(a):
    return a.ctypes.data


def test_rodrigues_matches_cv2(lib):
    rng = np.random.default_rng(0)
    vecs = [rng.normal(size=3) * s for s in (1e-12, 1e-3, 0.3, 1.0, 3.0) for _ in range(20)] + [np.zeros(3)]
    for r in vecs:
        r = np.ascontiguousarray(r, np.float64)
        R = np.zeros(9); J = np.zeros(27)
        lib.esacb200_host_rodrigues(_ptr(r), _ptr(R), _ptr(J))
        Rc, Jc = cv2.Rodrigues(r.reshape(3, 1))
        assert np.abs(R.reshape(3, 3) - Rc).max() < 1e-14
        assert np.abs(J.reshape(3, 9) - Jc).max() < 1e-12
        back = np.zeros(3)
        lib.esacb200_host_rodrigues_inv(_ptr(np.ascontiguousarray(Rc.reshape(-1))), _ptr(back))
        rc, _ = cv2.Rodrigues(Rc)
        assert np.abs(back - rc.ravel()).max() < 1e-9


def test_rodrigues_inverse_near_pi(lib):
    rng = np.random.default_rng(1)
    for _ in range(50):
        axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
        r = axis * (np.pi - rng.uniform(0, 1e-7))
        Rc, _ = cv2.Rodrigues(r.reshape(3, 1))
        back = np.zeros(3)
        lib.esacb200_host_rodrigues_inv(_ptr(np.ascontiguousarray(Rc.reshape(-1))), _ptr(back))
        rc, _ = cv2.Rodrigues(Rc)
        assert np.abs(back - rc.ravel()).max() < 1e-6


def _random_case(rng, noise):
    rv = rng.normal(size=3); rv *= rng.uniform(0, 0.8) / np.linalg.norm(rv)
    R, _ = cv2.Rodrigues(rv)
    t = rng.uniform(-1, 1, 3) + np.array([0, 0, 3.0])
    X = rng.uniform(-1, 1, (4, 3)).astype(np.float32)
    xc = (R @ X.T.astype(np.float64)).T + t
    uv = xc[:, :2] / xc[:, 2:] * 525 + np.array([320, 240])
    uv = np.ascontiguousarray(np.round(uv + rng.normal(0, noise, uv.shape)).astype(np.float32))
    return X, uv


@pytest.mark.parametrize("noise", [0.0, 1.0, 4.0])
def test_p3p_matches_cv2_solvepnp(lib, noise):
    """solvePnP(4 points, SOLVEPNP_P3P): same pose, incl. the 4th-point disambiguation and failures."""
    rng = np.random.default_rng(int(noise * 10) + 1)
    n_cmp = 0
    for _ in range(1500):
        X, uv = _random_case(rng, noise)
        ok, r1, t1 = cv2.solvePnP(X.reshape(-1, 1, 3), uv.reshape(-1, 1, 2), K, None, flags=cv2.SOLVEPNP_P3P)
        cv_ok = ok and not np.isnan(t1).any()
        pose = np.zeros(6); gate = C.c_int()
        mine = lib.esacb200_host_p3p_pose(_ptr(X), _ptr(uv), 525.0, 320.0, 240.0, 10.0, _ptr(pose), C.byref(gate))
        assert bool(mine) == bool(cv_ok)
        if not cv_ok:
            continue
        d = max(np.abs(pose[:3] - r1.ravel()).max(), np.abs(pose[3:] - t1.ravel()).max())
        # two P3P roots whose 4th-point errors tie can legitimately be ordered differently; that is rare
        if d < 1e-2:
            assert d < 1e-8
            n_cmp += 1
            # the reference's 4-point gate (esac_util.h:202-223) on float-rounded projections
            proj, _ = cv2.projectPoints(X.reshape(-1, 1, 3), r1, t1, K, None)
            dd = uv - proj.reshape(-1, 2).astype(np.float32)
            n = np.sqrt(dd[:, 0].astype(np.float64) ** 2 + dd[:, 1].astype(np.float64) ** 2)
            assert bool(gate.value) == bool(np.all(n < 10.0))
    assert n_cmp > 1400


def test_p3p_degenerate_inputs_fail(lib):
    uv = np.array([[4, 4], [12, 4], [4, 12], [20, 20]], np.float32)
    for X in (np.zeros((4, 3), np.float32), np.array([[0, 0, 1], [1, 0, 1], [2, 0, 1], [3, 0, 1]], np.float32)):
        pose = np.zeros(6); gate = C.c_int()
        assert lib.esacb200_host_p3p_pose(_ptr(X), _ptr(uv), 525.0, 320.0, 240.0, 10.0, _ptr(pose), C.byref(gate)) == 0


def test_p3p_all_roots_match_cv2_solvep3p(lib):
    rng = np.random.default_rng(5)
    for _ in range(300):
        X, uv = _random_case(rng, 0.0)
        n, rvs, tvs = cv2.solveP3P(X[:3].reshape(-1, 1, 3), uv[:3].reshape(-1, 1, 2), K, None, flags=cv2.SOLVEPNP_P3P)
        y = np.zeros((3, 3))
        for i in range(3):
            b = np.array([np.float32((float(uv[i, 0]) - 320.0) * (1 / 525.0)), np.float32((float(uv[i, 1]) - 240.0) * (1 / 525.0)), 1.0],
                         np.float64)
            y[i] = b / np.linalg.norm(b)
        x = np.ascontiguousarray(X[:3].astype(np.float64))
        Rs = np.zeros(36); ts = np.zeros(12)
        m = lib.esacb200_host_p3p_all(_ptr(y), _ptr(x), _ptr(Rs), _ptr(ts))
        assert m >= n
        for rv, tv in zip(rvs, tvs):
            Rc, _ = cv2.Rodrigues(rv)
            best = min(np.abs(Rs[9 * s:9 * s + 9].reshape(3, 3) - Rc).max() + np.abs(ts[3 * s:3 * s + 3] - tv.ravel()).max()
                       for s in range(m))
            assert best < 1e-7


def test_projection_matches_cv2_projectpoints(lib):
    rng = np.random.default_rng(2)
    for _ in range(200):
        pose = np.concatenate([rng.normal(0, 0.4, 3), rng.normal(0, 1, 3) + [0, 0, 2]])
        X = rng.uniform(-2, 2, 3).astype(np.float32)
        uvf = np.zeros(2, np.float32); uv = np.zeros(2); J = np.zeros(12)
        lib.esacb200_host_project(_ptr(pose), 525.0, 320.0, 240.0, _ptr(X), _ptr(uvf), _ptr(uv), _ptr(J))
        p, Jc = cv2.projectPoints(X.reshape(1, 1, 3), pose[:3].reshape(3, 1), pose[3:].reshape(3, 1), K, None)
        assert p.dtype == np.float32
        assert np.array_equal(p.ravel(), uvf)  # bit-exact float rounding
        assert np.abs(J.reshape(2, 6) - Jc[:, :6]).max() < 1e-9 * (1 + np.abs(Jc[:, :6]).max())


def test_loss_dloss_pose_conversions_match_oracle(lib):
    rng = np.random.default_rng(3)
    for i in range(100):
        est = np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 2, 3)])
        gt = est + rng.normal(0, 0.05 if i % 2 else 1.0, 6)
        T1 = np.zeros(16); T2 = np.zeros(16)
        lib.esacb200_host_pose2trans(_ptr(est), _ptr(T1))
        lib.esacb200_host_pose2trans(_ptr(gt), _ptr(T2))
        assert np.abs(T1.reshape(4, 4) - O.pose2trans(est[:3], est[3:])).max() < 1e-12
        for cut in (100.0, 1.0):
            l = lib.esacb200_host_loss(_ptr(T1), _ptr(T2), 1.0, 100.0, cut)
            assert abs(l - O.loss(T1.reshape(4, 4), T2.reshape(4, 4), 1.0, 100.0, cut)) < 1e-9 * max(1, l)
            d = np.zeros(6)
            lib.esacb200_host_dloss(_ptr(est), _ptr(gt), 1.0, 100.0, cut, _ptr(d))
            ref = O.d_loss(est[:3], est[3:], gt[:3], gt[3:], 1.0, 100.0, cut).ravel()
            assert np.abs(d - ref).max() < 1e-8 * (1 + np.abs(ref).max())
        back = np.zeros(6)
        lib.esacb200_host_trans2pose(_ptr(T1), _ptr(back))
        r, t = O.trans2pose(T1.reshape(4, 4))
        assert np.abs(back - np.concatenate([r.ravel(), t.ravel()])).max() < 1e-9


def test_dloss_is_gradient_of_loss_without_cut(lib):
    """Below the cut, dLoss is the true gradient of loss() (finite differences on the fp64 hook)."""
    rng = np.random.default_rng(4)
    est = np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 1, 3)])
    gt = est + rng.normal(0, 0.01, 6)
    T2 = np.zeros(16)
    lib.esacb200_host_pose2trans(_ptr(gt), _ptr(T2))

    def f(p):
        T = np.zeros(16)
        lib.esacb200_host_pose2trans(_ptr(np.ascontiguousarray(p)), _ptr(T))
        return lib.esacb200_host_loss(_ptr(T), _ptr(T2), 1.0, 100.0, 1e9)

    d = np.zeros(6)
    lib.esacb200_host_dloss(_ptr(est), _ptr(gt), 1.0, 100.0, 1e9, _ptr(d))
    num = np.array([(f(est + 1e-6 * np.eye(6)[i]) - f(est - 1e-6 * np.eye(6)[i])) / 2e-6 for i in range(6)])
    # loss() uses PI=3.1415926, dLoss uses CV_PI: the rotation part differs by that ratio only
    assert np.abs(d - num).max() < 1e-4 * (1 + np.abs(num).max())


def test_dprojectdobj_matches_oracle(lib):
    rng = np.random.default_rng(6)
    for _ in range(100):
        pose = np.concatenate([rng.normal(0, 0.4, 3), rng.normal(0, 1, 3) + [0, 0, 2]])
        X = rng.uniform(-2, 2, 3).astype(np.float32)
        pt = rng.uniform(0, 640, 2).astype(np.float32)
        out = np.zeros(3)
        lib.esacb200_host_dprojectdobj(_ptr(pt), _ptr(X), _ptr(pose), 525.0, 320.0, 240.0, 100.0, _ptr(out))
        R, _ = cv2.Rodrigues(pose[:3].reshape(3, 1))
        ref = O.d_project_d_obj(pt, X, R, pose[3:].reshape(3, 1), K, 100.0).ravel()
        assert np.abs(out - ref).max() < 1e-9 * (1 + np.abs(ref).max())


def test_pinv6_matches_cv2_svd_inverse(lib):
    rng = np.random.default_rng(7)
    for rank in (6, 6, 6, 4):
        J = rng.normal(size=(50, 6))
        if rank < 6:
            J[:, rank:] = J[:, :6 - rank] * 2.0  # exactly dependent columns -> pseudo-inverse
        A = np.ascontiguousarray(J.T @ J)
        out = np.zeros(36)
        lib.esacb200_host_pinv6(_ptr(A), _ptr(out))
        ref = cv2.invert(A, flags=cv2.DECOMP_SVD)[1]
        assert np.abs(out.reshape(6, 6) - ref).max() < 1e-7 * np.abs(ref).max()


def test_cell_stream_matches_oracle(lib):
    for (W, H) in ((80, 60), (5, 3), (107, 60)):
        for h in range(5):
            for t in range(20):
                cells = np.zeros(8, np.int32)
                lib.esacb200_host_draw_cells(C.c_uint64(1305 + h), h, t, W, H, _ptr(cells))
                ref = O.draw_minimal_set(1305 + h, h, t, W, H)
                assert cells.reshape(4, 2).tolist() == [list(c) for c in ref]
                assert cells.reshape(4, 2)[:, 0].max() <= W - 2 and cells.reshape(4, 2)[:, 1].max() <= H - 2
                assert len({tuple(c) for c in cells.reshape(4, 2)}) == 4


@pytest.mark.parametrize("kw,which", [({"seed": 1}, "gt"), ({"seed": 1, "active_only": False}, "other"),
                                      ({"seed": 3, "outdoor": True}, "gt"), ({"seed": 2, "world_offset": 700.0}, "gt")])
def test_float_prefilter_never_rejects_an_accepted_try(lib, kw, which):
    """The sampling kernel's float prefilter (p3p_may_pass) may only discard tries the exact fp64 path rejects."""
    from esac_b200.synth import make_scene
    sc = make_scene(E=2, H=60, W=80, M=8, sub=8, **kw)
    e = sc.gt_expert if which == "gt" else 1 - sc.gt_expert
    rng = np.random.default_rng(5)
    pl = sc.coords[e]
    mp, ac, ev = C.c_int(), C.c_int(), C.c_int()
    n_acc = n_may = 0
    n = 20000
    for _ in range(n):
        while True:
            xs = rng.integers(0, 79, 4); ys = rng.integers(0, 59, 4)
            if len({(a, b) for a, b in zip(xs, ys)}) == 4:
                break
        obj = np.ascontiguousarray(pl[:, ys, xs].T.astype(np.float32))
        img = np.ascontiguousarray(np.stack([xs * 8 + 4, ys * 8 + 4], 1).astype(np.float32))
        lib.esacb200_host_try(_ptr(obj), _ptr(img), sc.f, sc.ppx, sc.ppy, sc.tau, 2.0, C.byref(mp), C.byref(ac))  # 2.0 = kPrefilterMargin, the shipping band
        assert not (ac.value and not mp.value)
        pv = np.zeros(6)
        lib.esacb200_host_try_verdict(_ptr(obj), _ptr(img), sc.f, sc.ppx, sc.ppy, sc.tau, C.byref(ev), _ptr(pv))
        assert ev.value == ac.value   # the early exit / favourite candidate of the verdict path never change the exact decision
        if ac.value:                  # ... nor the pose that gets staged for an accepted try
            pf = np.zeros(6); gate = C.c_int()
            lib.esacb200_host_p3p_pose(_ptr(obj), _ptr(img), sc.f, sc.ppx, sc.ppy, sc.tau, _ptr(pf), C.byref(gate))
            assert np.array_equal(pv, pf)
        n_acc += ac.value; n_may += mp.value
    if which == "gt":
        assert n_acc > 500           # the invariant was exercised on accepted tries
    else:
        assert n_may < 0.1 * n       # and the prefilter does reject most tries on wrong experts
