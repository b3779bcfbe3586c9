"""Multi-GPU esac.forward: experts sharded expert-major across ranks, ONE all-gather of per-shard scores.

Hypotheses are independent through sampling, P3P, scoring and (per shard) refinement; the only exchange in
esac_forward is the softmax/argmax over all scores (esac.cpp:153-155).  Every rank therefore runs the complete
local pipeline on the experts it owns -- including the refinement of its local best hypothesis, which costs no
extra latency because the ranks run concurrently -- and a single all-gather of
    [ local scores (M_local) | local refined pose (16) | global expert id | local winner index ]
lets every rank pick the global winner with the reference's rule (first strict maximum, esac_util.h:519-523).
Messages are KB-sized: the collective is latency-bound and is issued once per image.
"""
from __future__ import annotations

import numpy as np


def pack_local(scores, pose16, expert_global: int, local_winner: int):
    """float64 vector [M_local + 18]."""
    import torch
    tail = torch.tensor([float(expert_global), float(local_winner)], dtype=torch.float64, device=scores.device)
    return torch.cat([scores.reshape(-1), pose16.reshape(16).to(torch.float64), tail])


def select_global(gathered: np.ndarray, M_local: int):
    """gathered: [world, M_local + 18].  Returns (global winner index, owning rank, pose 4x4 float32, expert id,
    probabilities of all hypotheses) with softMax / draw(training=false) semantics (esac_util.h:461-530)."""
    world = gathered.shape[0]
    scores = gathered[:, :M_local].reshape(-1)
    sf = np.exp(scores - scores.max())
    probs = sf / sf.sum()
    # first strict maximum among p >= EPS (draw(), training=false): argmax returns the first maximum
    winner = int(np.argmax(probs)) if probs.max() >= 1e-8 else 0
    rank = winner // M_local
    assert rank < world
    pose = gathered[rank, M_local:M_local + 16].reshape(4, 4).astype(np.float32)
    expert = int(gathered[rank, M_local + 16])
    return winner, rank, pose, expert, probs


def make_exchange(group=None, device=None):
    """The two reductions of the sharded backward as torch.distributed collectives (NCCL on `device`, gloo on CPU):
    phase 1 all-gathers (max, sum exp) pairs and merges them into the global softmax normalisation, phase 2 sums the
    partial expectations."""
    import torch
    import torch.distributed as dist

    def exchange(phase, values):
        world = dist.get_world_size(group)
        t = torch.tensor(values, dtype=torch.float64, device=device)
        if phase == 1:
            g = torch.empty(world * 2, dtype=torch.float64, device=device)
            dist.all_gather_into_tensor(g, t, group=group)
            g = g.view(world, 2).cpu().numpy()
            gmax = float(g[:, 0].max())
            gsum = float((g[:, 1] * np.exp(g[:, 0] - gmax)).sum())
            return [gmax, gsum]
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return [float(v) for v in t.cpu()]

    return exchange


def backward_sharded(coords_local, grads_local, assign_local, gt_pose, w_rot, w_trans, cut, params, hyp_offset: int, group=None):
    """esac.backward with experts sharded expert-major across ranks: every rank owns its experts' planes and gradient
    slices (no gradient reduction); two KB-sized collectives give every rank the global softmax and the global expected
    loss, which it returns.  `hyp_offset` = number of hypotheses owned by lower ranks."""
    from . import api
    dev = coords_local.device if getattr(coords_local, "is_cuda", False) else None
    ex = make_exchange(group, dev)
    return api.backward_sharded(coords_local, grads_local, assign_local, gt_pose, w_rot, w_trans, cut, *params, exchange=ex,
                                hyp_offset=hyp_offset)


def forward_sharded(coords_local, assign_local, out_pose, params, expert_offset: int, group=None, local_forward=None,
                    hyp_offset: int = 0):
    """esac.forward over experts sharded across the ranks of `group`.  coords_local [E_local,3,H,W] and
    assign_local [M_local] (expert indices local to the shard) live on this rank; out_pose [4,4] receives the
    global winner's camera pose on every rank; returns the global expert index."""
    import torch
    import torch.distributed as dist
    from . import api

    M = int(assign_local.shape[0])
    if local_forward is None:
        if not hasattr(coords_local, "is_cuda"):
            coords_local, assign_local = torch.from_numpy(coords_local), torch.from_numpy(assign_local)
        dev = coords_local.device if coords_local.is_cuda else torch.device("cuda", torch.cuda.current_device())
        # host inputs (the reference's callers hold CPU tensors): staged on the current stream, asynchronously if pinned
        coords_local = coords_local.to(dev, non_blocking=True)
        assign_local = assign_local.to(dev, non_blocking=True)
        ctx = api.context(dev.index)
        ctx.set_option("hyp_offset", hyp_offset)
        buf = torch.empty(M + 18, dtype=torch.float64, device=dev)
        try:
            api.forward_pack(coords_local, assign_local, params, expert_offset, buf)   # enqueued, no host sync
        finally:
            ctx.set_option("hyp_offset", 0)
    else:
        scores, pose, e_local, lw = local_forward(coords_local, assign_local, params)
        buf = pack_local(scores, pose, expert_offset + e_local, lw)
    world = dist.get_world_size(group)
    gathered = torch.empty(world * (M + 18), dtype=torch.float64, device=buf.device)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    g = gathered.view(world, M + 18)
    if g.is_cuda:
        # selection on the device: first maximum = draw(training=false); ONE 17-value read-back (the only host sync of the step)
        w = torch.argmax(g[:, :M].reshape(-1))
        tail = g.reshape(-1)[(w // M) * (M + 18) + M + torch.arange(17, device=g.device)]
        if hasattr(out_pose, "is_cuda") and out_pose.is_cuda:
            out_pose.copy_(tail[:16].reshape(4, 4))
            expert = int(tail[16].item())
            if expert < 0:
                raise RuntimeError("hypAssignment holds an expert index outside the shard's experts")
            return expert
        small = tail.cpu().numpy()
        gpose, expert = small[:16].reshape(4, 4).astype(np.float32), int(small[16])
        if expert < 0:
            raise RuntimeError("hypAssignment holds an expert index outside the shard's experts")
    else:
        _, _, gpose, expert, _ = select_global(g.numpy(), M)
    if hasattr(out_pose, "copy_"):
        out_pose.copy_(torch.from_numpy(gpose))
    else:
        out_pose[:, :] = gpose
    return expert
