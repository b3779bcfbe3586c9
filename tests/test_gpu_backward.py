"""GPU parity of esac.backward against the cv2 oracle (run with `-m gpu`)."""
import numpy as np
import pytest

from esac_b200.synth import make_scene
from oracle import esac_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import esac_b200.api as api
    api.context().set_option("fixed_seed", 1)
    return api


@pytest.mark.parametrize("seed,E,H,W,M,kw", [
    (1, 1, 30, 40, 32, {}),
    (2, 3, 30, 40, 48, {}),
    (3, 2, 24, 31, 32, {"shiftX": 2, "shiftY": -3}),
    (4, 2, 30, 40, 32, {"outlier_frac": 0.8}),      # flat hypothesis distribution: many contributing hypotheses
])
def test_backward_matches_oracle(api, seed, E, H, W, M, kw):
    sc = make_scene(E=E, H=H, W=W, M=M, sub=8, seed=seed, **kw)
    g_ref = np.zeros_like(sc.coords)
    l_ref, bt = O.backward(sc.coords, g_ref, sc.assign, sc.gt_pose, 1.0, 100.0, 100.0, *sc.params, seed=50 + seed, trace=True)
    api.set_seed(50 + seed)
    g = np.zeros_like(sc.coords)
    loss = api.backward(sc.coords, g, sc.assign, sc.gt_pose, 1.0, 100.0, 100.0, *sc.params)
    hy = api.last_hypotheses(losses=True)
    st = api.last_stats()
    assert np.abs(hy["scores"] - np.array(bt.scores)).max() < 1e-4
    assert st["n_contrib"] == int((bt.probs >= O.PROB_THRESH).sum())
    assert np.abs(hy["losses"] - np.array(bt.losses)).max() < 1e-5 * max(1.0, np.abs(bt.losses).max())
    assert abs(loss - l_ref) < 1e-6 * max(1.0, abs(l_ref))
    scale = max(np.abs(g_ref).max(), 1e-12)
    assert np.abs(g - g_ref).max() / scale < 1e-3, (np.abs(g - g_ref).max(), scale)


def test_backward_accumulates_in_place(api):
    sc = make_scene(E=1, H=24, W=32, M=16, sub=8, seed=7)
    api.set_seed(3)
    g0 = np.zeros_like(sc.coords)
    api.backward(sc.coords, g0, sc.assign, sc.gt_pose, 1.0, 100.0, 100.0, *sc.params)
    api.set_seed(3)
    g1 = np.full_like(sc.coords, 0.5)
    api.backward(sc.coords, g1, sc.assign, sc.gt_pose, 1.0, 100.0, 100.0, *sc.params)
    assert np.abs((g1 - 0.5) - g0).max() < 1e-6 * max(1.0, np.abs(g0).max())


def test_backward_cuda_tensors(api):
    import torch
    sc = make_scene(E=2, H=24, W=32, M=24, sub=8, seed=8)
    api.set_seed(4)
    g_cpu = torch.zeros(sc.coords.shape)
    l1 = api.backward(torch.from_numpy(sc.coords), g_cpu, torch.from_numpy(sc.assign), torch.from_numpy(sc.gt_pose),
                      1.0, 100.0, 100.0, *sc.params)
    api.set_seed(4)
    g_gpu = torch.zeros(sc.coords.shape, device="cuda")
    l2 = api.backward(torch.from_numpy(sc.coords).cuda(), g_gpu, torch.from_numpy(sc.assign).cuda(),
                      torch.from_numpy(sc.gt_pose).cuda(), 1.0, 100.0, 100.0, *sc.params)
    assert l1 == l2
    assert torch.equal(g_cpu, g_gpu.cpu())


def test_autograd_wrapper_matches_manual_gradients(api):
    """EsacLoss = esac.backward + the trainer's hand-made gating gradient (train_esac.py:171-180)."""
    import torch
    from esac_b200.autograd import esac_loss
    sc = make_scene(E=3, H=24, W=32, M=24, sub=8, seed=12)
    coords = torch.from_numpy(sc.coords).cuda().requires_grad_(True)
    gating = torch.log_softmax(torch.zeros(1, 3, device="cuda"), dim=1).requires_grad_(True)
    assign = torch.from_numpy(sc.assign)
    api.set_seed(9)
    loss = esac_loss(coords, gating, assign, torch.from_numpy(sc.gt_pose), 1.0, 100.0, 100.0, *sc.params)
    loss.backward()
    api.set_seed(9)
    g = np.zeros_like(sc.coords)
    l_ref = api.backward(sc.coords, g, sc.assign, sc.gt_pose, 1.0, 100.0, 100.0, *sc.params)
    assert abs(loss.item() - l_ref) < 1e-6 * max(1.0, abs(l_ref))
    assert np.array_equal(coords.grad.cpu().numpy(), g)
    hist = np.bincount(sc.assign, minlength=3).astype(np.float32)
    assert np.allclose(gating.grad.cpu().numpy().ravel(), l_ref * hist, rtol=1e-6)


def test_autograd_wrapper_expert_selection_branch(api):
    """train_esac.py:133-136,171-173: one expert drawn and expanded to all hypotheses -> the gating gradient is `loss` at
    that expert (not loss * M, which the histogram formula of the other branch would give)."""
    import torch
    from esac_b200.autograd import esac_loss
    sc = make_scene(E=3, H=24, W=32, M=24, sub=8, seed=13)
    expert = torch.tensor([sc.gt_expert], dtype=torch.int64)
    e_hyps = expert.expand(24)  # stride-0 view, exactly what the trainer builds
    for flag in (None, True):
        coords = torch.from_numpy(sc.coords).cuda().requires_grad_(True)
        gating = torch.log_softmax(torch.zeros(1, 3, device="cuda"), dim=1).requires_grad_(True)
        api.set_seed(10)
        loss = esac_loss(coords, gating, e_hyps, torch.from_numpy(sc.gt_pose), 1.0, 100.0, 100.0, *sc.params, expert_selection=flag)
        loss.backward()
        want = np.zeros(3, np.float32)
        want[sc.gt_expert] = loss.item()
        assert np.allclose(gating.grad.cpu().numpy().ravel(), want, rtol=1e-6)
    # the same assignment as a materialised tensor with expert_selection=False is the histogram branch
    coords = torch.from_numpy(sc.coords).cuda().requires_grad_(True)
    gating = torch.log_softmax(torch.zeros(1, 3, device="cuda"), dim=1).requires_grad_(True)
    api.set_seed(10)
    loss = esac_loss(coords, gating, e_hyps.contiguous(), torch.from_numpy(sc.gt_pose), 1.0, 100.0, 100.0, *sc.params)
    loss.backward()
    assert np.allclose(gating.grad.cpu().numpy().ravel()[sc.gt_expert], 24 * loss.item(), rtol=1e-6)


def test_reference_training_step_runs_through_the_drop_in_module():
    """examples/train_step_synthetic.py = train_esac.py:105-183 with stand-in networks: esac.backward drives autograd."""
    import runpy
    import sys
    from pathlib import Path
    ex = Path(__file__).resolve().parents[1] / "examples" / "train_step_synthetic.py"
    argv = sys.argv
    sys.argv = [str(ex), "--iterations", "2", "--hypotheses", "64"]
    try:
        mod = runpy.run_path(str(ex), run_name="example")
        losses = mod["main"]()
    finally:
        sys.argv = argv
    assert len(losses) == 2 and all(np.isfinite(l) and l >= 0 for l in losses)
