#!/usr/bin/env python
"""Benchmark of the ESAC hot path (esac.forward) on B200s.

Metric (BASELINE.json): pose hypotheses scored per second at 640x480, 256 hypotheses x E experts.
A step = one esac.forward call (sample -> score -> select -> refine) over one synthetic image.

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repository's CUDA path
  python bench.py --impl reference ...                         the reference's CPU path (cv2 oracle port,
                                                               all host cores; the reference extension
                                                               itself cannot be built in this image)
For N > 1 launch under torchrun (one rank per GPU); experts are sharded expert-major across the ranks,
every rank runs the full local pipeline on its shard and ONE NCCL all-gather of the per-shard scores
(+ the 17-float local result) picks the global winner.

One JSON line on stdout (rank 0): value = device-resident whole-job hypotheses/s; e2e = the same through
esac.forward with pinned HOST tensors (H2D copy of the coordinate maps and D2H of the pose inside the
timed region); roofline = the scoring kernel's algorithmic bytes (12*N per hypothesis) / its CUDA-event
time / the measured HBM peak; cpu_baseline = the oracle on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "pose hypotheses scored/sec (640x480, 256 hyp x E experts)"
UNIT = "hypotheses/s"
E_PER_GPU, HYPS_PER_EXPERT, H, W, SUB = 7, 256, 480, 640, 1
N_SCENES = 6  # rotating inputs: 6 x 25.8 MB > 126 MB L2, so no step finds its planes in L2


def workload_config(n_gpus: int) -> dict:
    return {"workload": f"BASELINE configs[1]: 7Scenes-style ensemble, {E_PER_GPU} experts/GPU x {HYPS_PER_EXPERT} "
                        f"hypotheses each ({E_PER_GPU * HYPS_PER_EXPERT * n_gpus} total), {W}x{H} scene-coordinate maps "
                        f"(subSampling=1), f=525, tau=10 alpha=100 beta=0.5 maxReproj=100, 40% outliers",
            "experts_per_gpu": E_PER_GPU, "hyps_per_expert": HYPS_PER_EXPERT, "map": [H, W],
            "hyps_total": E_PER_GPU * HYPS_PER_EXPERT * n_gpus,
            "parallelism": "single GPU" if n_gpus == 1 else f"expert-major shard over {n_gpus} GPUs + 1 all-gather of scores",
            "l2_policy": f"inputs larger than L2: {N_SCENES} distinct scenes ({N_SCENES * E_PER_GPU * 3 * H * W * 4 / 1e6:.0f} MB) used round-robin"}


def make_inputs(rank: int, n: int = N_SCENES):
    from esac_b200.synth import make_scene
    return [make_scene(E=E_PER_GPU, H=H, W=W, M=HYPS_PER_EXPERT, sub=SUB, seed=100 * rank + i, per_expert=True,
                       active_only=False) for i in range(n)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx = float(p[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference_run(scene, take: int, seed: int):
    from oracle.parallel import forward_parallel, scene_dict
    e, T, secs, workers = forward_parallel(scene_dict(scene, take), seed=seed)
    return take / secs, secs, workers


# -------------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU path (oracle port on cv2) on a bounded sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    take = max(cores, 32)
    take = min(take, E_PER_GPU * HYPS_PER_EXPERT)
    from esac_b200.synth import make_scene
    sc = make_scene(E=E_PER_GPU, H=H, W=W, M=HYPS_PER_EXPERT, sub=SUB, seed=0, per_expert=True, active_only=False)
    # bounded sample: `take` hypotheses spread over the experts (every 1792/take-th hypothesis)
    idx = np.linspace(0, len(sc.assign) - 1, take).astype(int)
    sc.assign = sc.assign[idx]
    # every step is one bounded sample; the number of steps actually run is capped so that the arm ends within minutes
    # whatever K the driver passes (a step costs seconds of CPU time, the metric is a rate and does not depend on K)
    warm = min(args.warmup, 3)
    steps = min(args.steps, 24)
    for _ in range(warm):
        cpu_reference_run(sc, take, 1)
    t0 = time.perf_counter()
    for k in range(steps):
        cpu_reference_run(sc, take, 2 + k)
    dt = time.perf_counter() - t0
    val = take * steps / dt
    cfg = workload_config(args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "steps_run": steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64/f32 mix (OpenCV)", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{take} of {E_PER_GPU * HYPS_PER_EXPERT} hypotheses per step, full esac.forward "
                                       f"(sample+score+select+refine) via the cv2 oracle, fork pool over hypotheses"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this implementation has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # the driver's own setting (it greps NCCL's communicator lines) wins
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import esac
    import esac_b200.api as api
    from esac_b200 import sharded

    ctx = api.context(local_rank)
    scenes = make_inputs(rank)
    M_local = len(scenes[0].assign)
    M_total = M_local * world
    N = H * W
    dev = torch.device("cuda", local_rank)
    d_coords = [torch.from_numpy(s.coords).to(dev) for s in scenes]
    d_assign = [torch.from_numpy(s.assign).to(dev) for s in scenes]
    h_coords = [torch.from_numpy(s.coords).pin_memory() for s in scenes]
    h_assign = [torch.from_numpy(s.assign).pin_memory() for s in scenes]
    d_out = torch.zeros(4, 4, device=dev)
    h_out = torch.zeros(4, 4).pin_memory()
    # input preparation, not a step: the first transfers out of freshly pinned pages run at a fraction of the steady PCIe rate
    # (measured: ~22 GB/s over the first 20 copies, >50 GB/s afterwards), so every pinned buffer is pushed through a few times
    scratch = torch.empty_like(d_coords[0])
    for _ in range(8):
        for hcrd in h_coords:
            scratch.copy_(hcrd, non_blocking=True)
    torch.cuda.synchronize()
    del scratch
    params = scenes[0].params
    api.set_seed(1305 + rank, local_rank)

    def step(i, host: bool):
        j = i % N_SCENES
        if world == 1:
            if host:
                return esac.forward(h_coords[j], h_assign[j], h_out, *params)
            return esac.forward(d_coords[j], d_assign[j], d_out, *params)
        if host:
            return sharded.forward_sharded(h_coords[j], h_assign[j], h_out, params, expert_offset=rank * E_PER_GPU)
        return sharded.forward_sharded(d_coords[j], d_assign[j], d_out, params, expert_offset=rank * E_PER_GPU)

    def timed(host: bool, steps: int, warmup: int):
        for i in range(warmup):
            step(i, host)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms_score, launches, score_launches = 0.0, 0, 0
        e0.record()
        for i in range(steps):
            step(warmup + i, host)
            st = ctx.stats()
            ms_score += st["ms_score"]
            launches += st["kernel_launches"]
            score_launches += st["score_launches"]
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ms_score, launches, score_launches, ctx.stats()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, ms_score, launches, score_launches, last = timed(False, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_e2e, _, _, _, _ = timed(True, args.steps, args.warmup)
    if world > 1:
        # the sharded step enqueues the local pipeline without a host synchronisation, so it collects no stage timers;
        # the scoring kernel's launch time (roofline) and the stage breakdown come from an untimed probe of the same local
        # shard through the synchronising single-GPU entry
        ms_score, score_launches = 0.0, 0
        for i in range(12):
            esac.forward(d_coords[i % N_SCENES], d_assign[i % N_SCENES], d_out, *params)
            if i >= 2:
                st = ctx.stats()
                ms_score += st["ms_score"]
                score_launches += st["score_launches"]
        last = ctx.stats()
        dist.barrier()

    # informative: the batched entry point (8 images per call, pinned host tensors, copy of image b+1 overlapping image b)
    batched = None
    if world == 1:
        Bsz = 8
        hb_coords = torch.stack([h_coords[i % N_SCENES] for i in range(Bsz)]).pin_memory()
        hb_assign = torch.stack([h_assign[i % N_SCENES] for i in range(Bsz)])
        hb_out = torch.zeros(Bsz, 4, 4).pin_memory()
        reps = max(2, args.steps // 8)
        api.forward_batch(hb_coords, hb_assign, hb_out, *params)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            api.forward_batch(hb_coords, hb_assign, hb_out, *params)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        batched = {"value": M_local * Bsz * reps / dt, "unit": UNIT, "images_per_call": Bsz, "ms_per_image": 1e3 * dt / (reps * Bsz),
                   "note": "esac_b200.api.forward_batch, pinned host maps, H2D overlapped with compute, one sync per call (host wall clock)"}

    # informative: esac.backward on the same workload, device-resident tensors
    bwd = None
    if world == 1:
        gt = torch.from_numpy(scenes[0].gt_pose)
        g = torch.zeros_like(d_coords[0])
        reps = max(3, min(10, args.steps // 5))
        esac.backward(d_coords[0], g, d_assign[0], gt, 1.0, 100.0, 100.0, *params)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        contrib = 0
        for i in range(reps):
            j = i % N_SCENES
            esac.backward(d_coords[j], g, d_assign[j], torch.from_numpy(scenes[j].gt_pose), 1.0, 100.0, 100.0, *params)
            contrib += ctx.stats()["n_contrib"]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        bwd = {"ms_per_call": 1e3 * dt / reps, "value": M_local * reps / dt, "unit": UNIT,
               "contributing_hypotheses_per_call": contrib / reps,
               "note": "esac.backward (sample+score+refine every hypothesis with p>=1e-3 + gradients), CUDA tensors, host wall clock"}

    if rank == 0:
        value = M_total * args.steps / (ms * 1e-3)
        e2e = M_total * args.steps / (ms_e2e * 1e-3)
        peak, peak_src = measured_peak()
        alg_bytes = M_local * 12.0 * N  # per scoring launch on this rank
        t_score = ms_score / max(score_launches, 1) * 1e-3
        achieved = alg_bytes / t_score / 1e9 if t_score > 0 else 0.0
        # the resource that actually binds the kernel: 3 MUFU ops per cell-hypothesis at 16 lanes/clk/SM (profiles/r01_pipes.txt)
        sm_count = api.context(local_rank).device_info()["sm_count"]
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        mufu_peak = sm_count * (16.0 / 3.0) * mhz * 1e6 / N  # hypotheses/s
        on_chip = {"pipe": "MUFU (rsqrt, ex2, rcp per cell-hypothesis; 16 lanes/clk/SM measured)", "peak_hyps_per_s": mufu_peak,
                   "achieved_hyps_per_s": M_local / t_score if t_score > 0 else None,
                   "frac": (M_local / t_score) / mufu_peak if t_score > 0 else None}
        traffic = None
        tp = ROOT / "profiles" / "score_kernel_traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (scoring, f32x2 FFMA2) / f64 (P3P, refinement)", "data": "synthetic",
                "config": workload_config(world), "clocks": clocks,
                "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": int(scenes[0].coords.nbytes + scenes[0].assign.nbytes), "d2h_bytes_per_step": 68},
                "gpu_launches": int(launches), "batched_e2e": batched, "backward": bwd,
                "roofline": {"kernel": "esacb200::score_kernel_tma<8>", "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": t_score * 1e3,
                             "hyps_per_s_kernel_only": M_local / t_score if t_score > 0 else None,
                             "note": "achieved = hypotheses x 12 B x 307200 cells / CUDA-event time of the scoring launch; "
                                     "each plane is read once per 64-hypothesis chunk, so frac > 1 is expected",
                             "on_chip_ceiling": on_chip},
                "stages_ms_last_step": {k: last[k] for k in ("ms_h2d", "ms_prep", "ms_sample", "ms_score", "ms_select",
                                                              "ms_refine", "ms_total")},
                "score_launch": {"ppt": last["score_ppt"], "grid": last["score_grid"], "refine_group": last["refine_group"]}}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            take = min(max(cores, 32), M_local)
            sc = scenes[0]
            idx = np.linspace(0, M_local - 1, take).astype(int)
            from copy import copy
            sc2 = copy(sc)
            sc2.assign = sc.assign[idx]
            v, secs, workers = cpu_reference_run(sc2, take, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": workers, "kind": "port",
                                    "sample": f"{take} of {M_local} hypotheses, one full esac.forward "
                                              f"(sample+score+select+refine) via the cv2 oracle, {secs:.1f} s"}
            # informative: the scoring stage alone in compiled C/OpenMP (oracle/esac_oracle_c.c, no OpenCV, no Python in the
            # loop) -- an upper bound on what a compiled CPU forward could reach on this host, next to the cv2 port above
            try:
                from oracle.build import c_score
                esac.forward(d_coords[0], d_assign[0], d_out, *params)
                poses = ctx.hypotheses()["poses"]
                n_c = min(M_local, max(256, 4 * cores))
                idx_c = np.linspace(0, M_local - 1, n_c).astype(int)
                c_score(sc.coords, sc.assign[idx_c[:cores]], poses[idx_c[:cores]], *params)      # warm-up, builds the library
                t0 = time.perf_counter()
                _, used = c_score(sc.coords, sc.assign[idx_c], poses[idx_c], *params)
                dtc = time.perf_counter() - t0
                line["cpu_baseline"]["compiled_scoring_only"] = {
                    "value": n_c / dtc, "unit": UNIT, "cores": int(used),
                    "sample": f"{n_c} of {M_local} hypotheses, getReproErrs+getHypScores restated in C/OpenMP, {dtc:.1f} s"}
            except Exception as exc:  # never let the informative leg break the bench line
                line["cpu_baseline"]["compiled_scoring_only"] = {"error": str(exc)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
