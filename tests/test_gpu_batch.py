"""GPU tests of the batched entry points of SURVEY 8f rank 2: device-side hypothesis assignment and backward_batch."""
import numpy as np
import pytest

from esac_b200.synth import make_scene
from oracle import esac_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import esac_b200.api as api
    api.context().set_option("fixed_seed", 0)
    return api


@pytest.mark.parametrize("B,E,M,keep,single", [(1, 1, 64, -1, False), (4, 7, 256, -1, False), (3, 19, 256, 5, False),
                                               (2, 20, 1024, -1, False), (5, 10, 33, 3, True), (1, 1000, 4096, 50, False)])
def test_assign_hypotheses_is_bit_exact_against_oracle(api, B, E, M, keep, single):
    import torch
    rng = np.random.default_rng(B * 100 + E)
    w = rng.random((B, E)).astype(np.float32) ** 4          # peaky, like a gating softmax
    w[0, rng.integers(0, E)] = 0.0 if E > 1 else w[0, 0]
    a_ref, h_ref = O.assign_hypotheses(w, M, seed=77, keep_top=keep, single=single)
    for kind in ("numpy", "cpu", "cuda"):
        x = w if kind == "numpy" else (torch.from_numpy(w) if kind == "cpu" else torch.from_numpy(w).cuda())
        a, h = api.assign_hypotheses(x, M, 77, maxExperts=keep, expertSelection=single)
        if kind != "numpy":
            assert a.device == x.device and h.device == x.device and a.dtype == torch.int64
            a, h = a.cpu().numpy(), h.cpu().numpy()
        assert np.array_equal(a, a_ref), kind
        assert np.array_equal(h, h_ref), kind


def test_assign_hypotheses_rejects_what_torch_multinomial_rejects(api):
    with pytest.raises(RuntimeError, match="sum of probabilities"):
        api.assign_hypotheses(np.zeros((2, 5), np.float32), 8, 1)
    with pytest.raises(RuntimeError, match="inf, nan or element < 0"):
        api.assign_hypotheses(np.array([[0.2, float("nan"), 0.1]], np.float32), 8, 1)
    with pytest.raises(RuntimeError, match="inf, nan or element < 0"):
        api.assign_hypotheses(np.array([[0.2, -1.0, 0.1]], np.float32), 8, 1)
    with pytest.raises(RuntimeError):
        api.assign_hypotheses(np.ones((2, 5), np.float64), 8, 1)


@pytest.mark.parametrize("workers", [1, 4])
@pytest.mark.parametrize("kind", ["cpu", "cuda"])
def test_backward_batch_equals_a_loop_of_backward(api, workers, kind):
    import torch
    B, E, H, W, M = 5, 3, 24, 32, 24
    scenes = [make_scene(E=E, H=H, W=W, M=M, sub=8, seed=40 + b) for b in range(B)]
    coords = np.stack([s.coords for s in scenes])
    assign = np.stack([s.assign for s in scenes])
    gts = np.stack([s.gt_pose for s in scenes])
    sx = [0, 2, -3, 1, 0]
    sy = [1, 0, -2, 4, 0]
    p = scenes[0].params  # (shiftX, shiftY, f, ppx, ppy, tau, alpha, beta, maxReproj, sub)
    ctx = api.context()
    ctx.set_option("batch_workers", workers)
    # reference: B consecutive calls
    api.set_seed(9)
    g_loop = np.zeros_like(coords)
    l_loop = []
    for b in range(B):
        l_loop.append(api.backward(coords[b], g_loop[b], assign[b], gts[b], 1.0, 100.0, 100.0, sx[b], sy[b], *p[2:]))
    api.set_seed(9)
    dev = "cuda" if kind == "cuda" else "cpu"
    t_coords = torch.from_numpy(coords).to(dev)
    t_grads = torch.zeros(coords.shape, device=dev)
    losses = api.backward_batch(t_coords, t_grads, torch.from_numpy(assign).to(dev), torch.from_numpy(gts).to(dev), 1.0, 100.0,
                                100.0, sx, sy, *p[2:])
    ctx.set_option("batch_workers", 4)
    assert np.allclose(losses, l_loop, rtol=1e-12, atol=0)
    assert np.array_equal(t_grads.cpu().numpy(), g_loop)
    assert api.last_stats()["kernel_launches"] > 0


def test_backward_batch_checks_shapes(api):
    import torch
    c = torch.zeros(2, 1, 3, 8, 10)
    with pytest.raises(RuntimeError):
        api.backward_batch(c, torch.zeros(2, 1, 3, 8, 9), torch.zeros(2, 4, dtype=torch.int64), torch.zeros(2, 4, 4), 1., 100., 100.,
                           0, 0, 525., 40., 32., 10., 100., 0.5, 100., 8)
    with pytest.raises(RuntimeError):
        api.backward_batch(c, torch.zeros_like(c), torch.zeros(2, 4, dtype=torch.int64), torch.zeros(2, 4, 4), 1., 100., 100.,
                           [0, 0, 0], 0, 525., 40., 32., 10., 100., 0.5, 100., 8)
