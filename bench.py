#!/usr/bin/env python
"""bench.py -- view-synthesis loss fwd+bwd throughput (BASELINE.json metric) on N B200s.

One "step" = generate_images_pred + compute_losses + loss.backward() for one KITTI-shaped batch
(12x3x192x640, frame_ids [0,-1,1], 4 scales, mono path: automask + selec_reproj + pose gradients;
/root/reference/ppeadepth/trainer.py:871-918, 1032-1160), i.e. BASELINE.json configs[1].

  value     Mpixels/s = N_gpus * B*H*W * steps / time, inputs resident in HBM, forward+backward of the
            C ABI replayed as CUDA graphs, rotating over input sets larger than L2.  `loss_check`: after
            the timed loop the replayed plan's loss is compared with the public API's on the same inputs.
  e2e       the same metric through the public host API (ppea_depth_b200.loss.ViewSynthesisLoss) with
            pinned HOST inputs: H2D of every input, forward, backward, D2H of the loss, every step; colour
            frames as the dataset's uint8 planes (expanded on the device, bit-identical to ToTensor), noise
            drawn on the device; median of the repetitions.  `e2e.float32_frames`: the same with fp32 frames.
  roofline  SURVEY.md §8d algorithmic bytes of the step / the dominant kernel's CUDA-event duration /
            measured HBM peak (MEASURED_PEAKS.json).
  cpu_baseline / --impl reference   the reference's own CPU implementation (the unmodified ppeadepth loss
            methods from oracle/_ref when present -- kind "reference" -- else the oracle port) timed on
            this box's host cores on a bounded sample; `threads_1` is the reference-faithful single-thread
            number (trainer.py:8-10 pins the pools to one thread), `torch_cuda_eager` the same ATen op
            sequence run by PyTorch on this GPU (what a user of the reference has today).
  N > 1     weak scaling (every rank owns its own 12-image batch; the loss path has no data-path collective,
            SURVEY.md §8e) is the headline; the line also carries the STRONG-scaling point of BASELINE
            configs[4] (`sharded_sweep96`: global batch 96 split over the ranks) and one NCCL all-reduce of a
            0.3 GB fp32 stand-in for the adapter gradients (trainer.py:350), alone and overlapped with the step.

Launch: `python bench.py` (1 GPU) or torchrun --nproc-per-node N bench.py --gpus N.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    "kitti": dict(batch=12, height=192, width=640, num_scales=4),
    "cityscapes": dict(batch=24, height=192, width=512, num_scales=4),
    "hires": dict(batch=8, height=320, width=1024, num_scales=4),
    "sweep96": dict(batch=96, height=192, width=640, num_scales=4),
}
L2_BYTES = 126 * 1024 * 1024
ALLREDUCE_FLOATS = 75_000_000        # 0.3 GB fp32: the adapter / decoder-adapter gradients of RepLKNet-31B stage 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS))
    ap.add_argument("--path", default="mono", choices=["mono", "multi"])
    ap.add_argument("--deterministic", action="store_true", help="(the default) coarse-scale gradient fields accumulated in 64-bit fixed point: bit-reproducible")
    ap.add_argument("--float-atomics", action="store_true", help="float atomics for the coarse-scale gradient fields instead (run-to-run summation order)")
    ap.add_argument("--shard", action="store_true", help="strong scaling: the workload's batch is split over the ranks")
    ap.add_argument("--no-fused", action="store_true", help="forward + backward kernel pair instead of the fused step")
    ap.add_argument("--tiles", action="store_true", help="fused step by the shared-memory tile kernel (round 1) instead of the streaming kernel")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sharded sweep / all-reduce legs (N > 1) and the torch-CUDA eager leg (N = 1)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of one repetition of the e2e leg (default: min(steps, 100))")
    args = ap.parse_args()
    args.deterministic = not args.float_atomics
    return args


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU, sampled every 100 ms while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._pump, daemon=True)
        self.t.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            inside = t0 - 0.05 <= t <= t1 + 0.15
            try:
                if inside:
                    sm.append(float(f[0]))
                    power.append(float(f[2]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            if inside:
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


# ----------------------------------------------------------------------------- workload
def make_sets(wl, n_sets, seed0, is_multi):
    """n_sets independent synthetic batches (host tensors)."""
    from ppea_depth_b200.synth import SynthConfig, make_batch, make_noise
    sets = []
    for i in range(n_sets):
        cfg = SynthConfig(seed=seed0 + i, **wl)
        inputs, outputs = make_batch(cfg)
        sets.append((inputs, outputs, make_noise(cfg, wl["num_scales"])))
    return sets


def _nbytes(shape, dtype):
    n = 1
    for d in shape:
        n *= d
    return n * torch.empty((), dtype=dtype).element_size()


def tensors_of_step(inputs, outputs, noise, S, is_multi):
    """The tensors one step reads (what the e2e leg copies host->device every step)."""
    t = {("in",) + k: v for k, v in inputs.items() if k[0] in ("K", "inv_K") and k[1] == 0}
    for f in (0, -1, 1):
        t[("in", "color", f, 0)] = inputs[("color", f, 0)]
    for s in range(1, S):
        t[("in", "color", 0, s)] = inputs[("color", 0, s)]
    for s in range(S):
        t[("out", "disp", s)] = outputs[("disp", s)]
    for f in (-1, 1):
        t[("out", "cam_T_cam", 0, f)] = outputs[("cam_T_cam", 0, f)]
    if is_multi:
        t[("out", "consistency_mask")] = outputs["consistency_mask"]
        t[("out", "augmentation_mask")] = outputs["augmentation_mask"]
        for s in range(S):
            t[("out", "mono_depth", 0, s)] = outputs[("mono_depth", 0, s)]
    else:
        for s in range(S):
            t[("noise", s)] = noise[s]
    return t


def build_plan(tset, wl, device, is_multi, deterministic, fused=None):
    from ppea_depth_b200.functional import VslConfig
    from ppea_depth_b200.runner import FusedPlan
    inputs, outputs, noise = tset
    S = wl["num_scales"]
    d = lambda x: x.to(device)
    cfg = VslConfig(is_multi=is_multi, deterministic=deterministic)
    if is_multi:
        kw = dict(cons_mask=d(outputs["consistency_mask"]), aug_mask=d(outputs["augmentation_mask"]),
                  mono_depth=[d(outputs[("mono_depth", 0, s)]) for s in range(S)])
    else:
        kw = dict(noise=[d(z) for z in noise])
    return FusedPlan(cfg, [d(outputs[("disp", s)]) for s in range(S)],
                     [d(outputs[("cam_T_cam", 0, -1)]), d(outputs[("cam_T_cam", 0, 1)])],
                     d(inputs[("color", 0, 0)]), [d(inputs[("color", -1, 0)]), d(inputs[("color", 1, 0)])],
                     d(inputs[("K", 0)]), d(inputs[("inv_K", 0)]), [d(inputs[("color", 0, s)]) for s in range(S)], fused=fused, **kw)


def barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world, device):
    if world == 1:
        return ms
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def time_replays(plans, K, Wm, world, device):
    """Wm warm-up + K timed graph replays rotating over `plans`; ms per step, max over ranks."""
    n = len(plans)
    for i in range(Wm):
        plans[i % n].replay()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        plans[i % n].replay()
    e1.record()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world, device) / K


def api_opt(B, H, W, S):
    from types import SimpleNamespace
    return SimpleNamespace(sclm=S - 1, v1_multiscale=False, height=H, width=W, min_depth=0.1, max_depth=100.0,
                           frame_ids=[0, -1, 1], disable_automasking=False, no_ssim=False, selec_reproj=True,
                           disable_motion_masking=False, no_matching_augmentation=False, batch_size=B,
                           disparity_smoothness=1e-3)


# ----------------------------------------------------------------------------- CPU reference arm
def time_cpu_reference(wl, is_multi, steps, warmup, budget_s=150.0, threads=None):
    """The reference's CPU implementation on the host cores: fwd+bwd on a bounded sample of the workload's batch.
    kind "reference": the UNMODIFIED ppeadepth.trainer.Trainer loss methods (oracle/_ref or /root/reference through
    oracle/ref_import.py); kind "port": the oracle restatement (the same ATen CPU ops in the same order)."""
    from oracle import ref_import as R
    from oracle import vsl_oracle as O
    from ppea_depth_b200.synth import SynthConfig, make_batch, make_noise
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if threads is not None:
        cores = threads
    kind = "port"
    if R.available():
        try:
            R.load_reference()
            kind = "reference"
        except Exception:                    # (an import the stubs do not cover: fall back to the port)
            kind = "port"
    torch.set_num_threads(cores)
    H, W, S = wl["height"], wl["width"], wl["num_scales"]

    def run(n_items, n_steps, n_warm):
        cfg = SynthConfig(seed=0, **dict(wl, batch=n_items))
        inputs, outputs = make_batch(cfg)
        noise = make_noise(cfg, S)
        opt = O.default_opt(sclm=S - 1, height=H, width=W, batch_size=n_items)
        ts = []
        for i in range(n_warm + n_steps):
            t0 = time.perf_counter()
            if kind == "reference":
                R.run_reference(inputs, outputs, opt, is_multi, None if is_multi else noise)
            else:
                O.run_fwd_bwd(inputs, outputs, opt, is_multi, noise)
            if i >= n_warm:
                ts.append(time.perf_counter() - t0)
        return ts

    t1 = run(1, 1, 1)[0]                       # calibrate on one image
    per_item_budget = budget_s / max(1, steps + warmup)
    n_items = int(max(1, min(wl["batch"], per_item_budget // max(t1, 1e-3))))
    ts = run(n_items, steps, warmup)
    mean_t = sum(ts) / len(ts)
    mpix = n_items * H * W / mean_t / 1e6
    sample = "%d of %d images per step (%dx%d, %d scales, %s path), %d steps after %d warm-up, mean %.3f s/step, %d threads" % (
        n_items, wl["batch"], H, W, S, "multi" if is_multi else "mono", steps, warmup, mean_t, cores)
    return mpix, mean_t * 1e3, cores, sample, kind


def time_torch_cuda_eager(wl, device, steps=10):
    """The reference's ATen op sequence (oracle port) run by PyTorch eager on THIS GPU: what a user of the reference has
    today (minus its nonzero() host syncs and its CPU noise).  Resident inputs, device noise."""
    from oracle import vsl_oracle as O
    from ppea_depth_b200.synth import SynthConfig, make_batch
    B, H, W, S = wl["batch"], wl["height"], wl["width"], wl["num_scales"]
    inputs, outputs = make_batch(SynthConfig(seed=7, **wl))
    ins = {k: v.to(device) for k, v in inputs.items()}
    base = {k: v.to(device) for k, v in outputs.items()}
    opt = O.default_opt(sclm=S - 1, height=H, width=W, batch_size=B)

    def step():
        outs = dict(base)
        for s in range(S):
            outs[("disp", s)] = base[("disp", s)].detach().requires_grad_(True)
        for f in (-1, 1):
            outs[("cam_T_cam", 0, f)] = base[("cam_T_cam", 0, f)].detach().requires_grad_(True)
        noise = [torch.randn(B, 1, H, W, device=device) for _ in range(S)]
        losses, _ = O.view_synthesis_losses(ins, outs, opt, False, noise)
        losses["loss"].backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": ms, "value": B * H * W / ms / 1e3, "unit": "Mpixels/s", "steps": steps,
            "what": "reference op sequence (oracle port of trainer.py:871-918 + 1032-1160) in PyTorch CUDA eager on this GPU, resident inputs, device noise"}


def pin_rank_to_cores(rank, world):
    """Every rank gets its own slice of the host cores (the box reports one NUMA node and one affinity mask for all
    GPUs): the ranks' enqueue threads and pinned-memory copies stop competing for the same cores."""
    if world <= 1 or not hasattr(os, "sched_getaffinity"):
        return None
    cpus = sorted(os.sched_getaffinity(0))
    per = max(1, len(cpus) // world)
    mine = cpus[rank * per:(rank + 1) * per] or cpus
    try:
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, min(4, len(mine))))
    except OSError:
        return None
    return mine


def main():
    args = parse()
    rank, world, local = dist_env()
    wl = dict(WORKLOADS[args.workload])
    is_multi = args.path == "multi"
    global_batch = wl["batch"]
    if args.shard and world > 1:
        wl["batch"] = max(1, wl["batch"] // world)
    B, H, W, S = wl["batch"], wl["height"], wl["width"], wl["num_scales"]
    cfg_desc = {"workload": "%s %dx3x%dx%d%s, frame_ids [0,-1,1], %d scales, %s path (generate_images_pred + compute_losses + backward)"
                % (args.workload, global_batch if args.shard else B, H, W, " split over the ranks" if args.shard else "", S, args.path),
                "pixels_per_step_per_gpu": B * H * W, "scales": S, "path": args.path, "deterministic_backward": bool(args.deterministic)}

    if args.impl == "reference":
        if rank != 0:
            return
        mpix, ms, cores, sample, kind = time_cpu_reference(wl, is_multi, max(1, args.steps), max(0, args.warmup))
        line = {"impl": "reference", "metric": "view-synthesis loss fwd+bwd throughput", "value": mpix, "unit": "Mpixels/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong" if args.shard else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg_desc,
                "cpu_baseline": {"value": mpix, "unit": "Mpixels/s", "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": mpix, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path for the product); "
                         "use --impl reference for the CPU reference arm")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    cores_mine = pin_rank_to_cores(rank, world)
    if world > 1:
        # NCCL's version banner / warnings off stdout (rank 0 prints ONE JSON line)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        torch.distributed.init_process_group("nccl", device_id=device)
    from ppea_depth_b200 import _cabi
    from ppea_depth_b200.synth import algorithmic_bytes
    _cabi.lib()          # fail loudly if the extension is missing
    fused_arg = False if args.no_fused else ("tiles" if args.tiles else None)

    # ---- resident input sets, sized to exceed L2
    probe = make_sets(wl, 1, 1000 * rank, is_multi)
    step_bytes = sum(v.numel() * v.element_size() for v in tensors_of_step(*probe[0], S, is_multi).values())
    n_sets = min(12, max(4, int(2.5 * L2_BYTES // step_bytes) + 1))
    if step_bytes > 2 * L2_BYTES:
        n_sets = 2
    sets = probe + make_sets(wl, n_sets - 1, 1000 * rank + 1, is_multi)
    plans = [build_plan(t, wl, device, is_multi, args.deterministic, fused=fused_arg) for t in sets]
    fused, tiles = plans[0].fused, plans[0].tiles
    kernels = ("fused step: preparation launch (packed sources, identity loss, target window sums) + warp-streaming kernel + smoothness launch in its shadow + finish + gradient finish "
               "(tails launched programmatically)" if fused and not tiles else
               "fused step: tile kernel (TMA-staged tiles) + finish + gradient finish" if fused else
               "vsl_forward_kernel + finish + vsl_backward_kernel + pose finish")
    for p in plans:
        p.capture()
    launches_per_step = plans[0].launches_forward + plans[0].launches_backward
    l2_note = "rotating %d resident input sets (%.0f MB read per step, %.0f MB total) > 126 MB L2" % (n_sets, step_bytes / 1e6, n_sets * step_bytes / 1e6)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # ---- value: HBM-resident, graph replay
    K, Wm = max(1, args.steps), max(3, args.warmup)
    t_clock0 = time.time()
    ms_step = time_replays(plans, K, Wm, world, device)
    value = world * B * H * W / (ms_step * 1e-3) / 1e6

    # ---- loss_check: the replayed plan against the public API on the same inputs
    from ppea_depth_b200.loss import ViewSynthesisLoss
    loss_check = None
    if rank == 0:
        plans[0].replay()
        torch.cuda.synchronize()
        loss_plan = float(plans[0].losses[0])
        inputs0, outputs0, noise0 = sets[0]
        ins = {k: v.to(device) for k, v in inputs0.items()}
        outs = {k: v.to(device) for k, v in outputs0.items()}
        for s in range(S):
            outs[("disp", s)].requires_grad_(True)
        mod = ViewSynthesisLoss(api_opt(B, H, W, S), deterministic=args.deterministic, fused=fused_arg)
        real_randn = torch.randn
        draws = iter(noise0)
        torch.randn = lambda *a, **k: next(draws).clone() if k.get("device") is None else real_randn(*a, **k)
        try:
            mod.generate_images_pred(ins, outs, is_multi)
            losses, _ = mod.compute_losses(ins, outs, is_multi)
        finally:
            torch.randn = real_randn
        losses["loss"].backward()
        loss_api = float(losses["loss"])
        g_api, g_plan = outs[("disp", 0)].grad, plans[0].grad_disp[0]
        gerr = float((g_api - g_plan).abs().max() / g_api.abs().max())
        loss_check = {"graph_replay_loss": loss_plan, "api_loss": loss_api, "rel_diff": abs(loss_plan - loss_api) / abs(loss_api),
                      "grad_disp0_rel_diff": gerr, "ok": abs(loss_plan - loss_api) <= 1e-6 * abs(loss_api) and gerr <= 1e-5}
        assert loss_check["ok"], loss_check

    # ---- roofline: per-kernel CUDA events (eager launches with the library's trace events), same rotation
    roof = None
    if rank == 0:
        acc = {}
        n_tr = min(K, 60)
        for p in plans:
            p.enable_trace()
        # (two steps in flight before the first trace is read: the host's launch latency stays out of the stage times)
        for i in range(n_tr):
            p = plans[i % n_sets]
            p.step()
            p.step()
            for k, v in p.trace_ms().items():
                acc[k] = acc.get(k, 0.0) + v
        for p in plans:
            p.disable_trace()
        stage_ms = {k: v / n_tr for k, v in acc.items()}
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        n_px = B * H * W
        # 86 + 24/4^s bytes per full-resolution pixel and scale on the mono path (forward 49 + 16/4^s,
        # backward 37 + 8/4^s), 98 + 24/4^s on the multi path, +8 with the deterministic backward: 375.9 B/px for 4 scales.
        step_bytes_alg = algorithmic_bytes(B, H, W, S, is_multi, args.deterministic and not fused)   # (the fused step has no full-resolution scratch in either mode: SURVEY 8d bytes)
        if fused:
            dom = "vsl_fused_kernel" if tiles else "vsl_stream_kernel"
        else:
            dom = "vsl_backward_kernel" if stage_ms["vsl_backward_kernel"] >= stage_ms["vsl_forward_kernel"] else "vsl_forward_kernel"
        # the fused kernels do the forward AND backward work of every pixel and scale in one launch: the §8d step bytes are
        # the algorithmic bytes of that launch (the pair path splits them between its two kernels)
        if fused:
            dom_bytes = step_bytes_alg
        else:
            fwd_b = sum((53.0 if is_multi else 49.0) + 16.0 / 4 ** s for s in range(S)) * n_px
            dom_bytes = step_bytes_alg - fwd_b if dom == "vsl_backward_kernel" else fwd_b
        achieved = dom_bytes / (stage_ms[dom] * 1e-3) / 1e9
        traffic = None            # DRAM bytes of one launch of that kernel from the committed ncu capture (same workload only)
        tpath = os.path.join(ROOT, "profiles", "r2_traffic.json" if (fused and not tiles) else ("r1j_traffic.json" if fused else "r1f_traffic.json"))
        if os.path.exists(tpath) and args.workload == "kitti" and not is_multi and (not args.deterministic or (fused and not tiles)) and not args.shard:
            traffic = json.load(open(tpath)).get(dom)
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "bytes_per_pixel": dom_bytes / n_px, "kernel_ms": stage_ms[dom], "stage_ms": stage_ms,
                "step_frac_of_peak": step_bytes_alg / (ms_step * 1e-3) / 1e9 / peak,
                "frac_of_8TBs_nominal": achieved / 8000.0,
                "note": "the step is bound by instruction issue and dependent-issue latency, not by HBM (profiles/README.md): "
                        "DRAM traffic of the launch is far below the algorithmic bytes"}

    # ---- e2e: public host API, pinned host inputs, H2D + fwd + bwd + D2H(loss) every step
    e2e = None
    if not args.no_e2e:
        mod = ViewSynthesisLoss(api_opt(B, H, W, S), deterministic=args.deterministic, noise_mode="device", fused=fused_arg)

        def collate(t):
            """One pinned arena per batch (what a collate_fn writing into a pinned buffer gives): a step's inputs cross
            PCIe as ONE copy.  Returns (arena, layout) with layout[key] = (byte offset, shape, dtype); uint8 frames first,
            contiguous, so that they can be expanded on the device by one call."""
            layout, off = {}, 0
            for k in sorted(t, key=lambda k: (t[k].dtype != torch.uint8, str(k))):
                v = t[k]
                off = (off + 255) // 256 * 256
                layout[k] = (off, tuple(v.shape), v.dtype)
                off += v.numel() * v.element_size()
            arena = torch.empty(off, dtype=torch.uint8).pin_memory()
            for k, (o, shape, dt) in layout.items():
                arena[o:o + t[k].numel() * t[k].element_size()].view(dt).view(shape).copy_(t[k])
            return arena, layout

        step_tensors = []
        for (inputs, outputs, noise) in sets[:4]:
            t = tensors_of_step(inputs, outputs, noise, S, is_multi)
            step_tensors.append({k: v for k, v in t.items() if k[0] != "noise"})   # the noise is drawn on the device (noise_mode="device")
        Ke = args.e2e_steps or min(K, 100)    # (the two-batch pipeline fill at the start is inside the timed region)
        copy_stream = torch.cuda.Stream(device=device)

        DEPTH = 2     # batches in flight on the copy stream (a pin_memory DataLoader's prefetch_factor)
        NSLOT = DEPTH + 1
        stage = {}    # device staging of the current leg: a ring of NSLOT input arenas with their tensor views built once

        def build_stage(host_sets):
            """The device side of the input pipeline, built once per leg: NSLOT arenas (step i uses slot i % NSLOT while the
            copy stream fills the next two), the views of every input tensor inside each arena, and -- for uint8 frames --
            the float32 block they are expanded into (images_to_float(out=...))."""
            arena0, layout = host_sets[0]
            assert all(h[1] == layout and h[0].numel() == arena0.numel() for h in host_sets)
            u8 = [k for k, (o, shape, dt) in layout.items() if dt == torch.uint8]
            lo = min((layout[k][0] for k in u8), default=0)
            hi = max((layout[k][0] + _nbytes(layout[k][1], layout[k][2]) for k in u8), default=0)
            slots = []
            for _ in range(NSLOT):
                d_arena = torch.empty(arena0.numel(), dtype=torch.uint8, device=device)
                dev = {k: d_arena[o:o + _nbytes(shape, dt)].view(dt).view(shape) for k, (o, shape, dt) in layout.items()}
                f32 = torch.empty(hi - lo, dtype=torch.float32, device=device) if u8 else None
                for k in u8:
                    o, shape, _ = layout[k]
                    dev[k] = f32[o - lo:o - lo + _nbytes(shape, torch.uint8)].view(shape)
                ins = {k[1:]: v for k, v in dev.items() if k[0] == "in"}
                outs = {(k[1] if len(k) == 2 else k[1:]): v for k, v in dev.items() if k[0] == "out"}
                slots.append({"arena": d_arena, "u8": d_arena[lo:hi] if u8 else None, "f32": f32, "ins": ins, "outs": outs,
                              "arrived": torch.cuda.Event(), "free": None})
            stage["slots"] = slots

        def start_h2d(hs, i):
            """Enqueues the host->device copy of step i's inputs on the copy stream (as a pin_memory dataloader does ahead of
            the step) into slot i % NSLOT, once the step that last used the slot is done with it."""
            sl = stage["slots"][i % NSLOT]
            with torch.cuda.stream(copy_stream):
                if sl["free"] is not None:
                    copy_stream.wait_event(sl["free"])
                sl["arena"].copy_(hs[0], non_blocking=True)
                sl["arrived"].record(copy_stream)
            return sl

        def e2e_compute(sl):
            from ppea_depth_b200 import images_to_float
            cur = torch.cuda.current_stream()
            cur.wait_event(sl["arrived"])
            if sl["u8"] is not None:
                # uint8 frames: one expansion call over their contiguous block (public API: images_to_float)
                images_to_float(sl["u8"], out=sl["f32"])
            ins, outs = sl["ins"], dict(sl["outs"])
            for s in range(S):
                outs[("disp", s)] = outs[("disp", s)].detach().requires_grad_(True)
            if not is_multi:
                for f in (-1, 1):
                    outs[("cam_T_cam", 0, f)] = outs[("cam_T_cam", 0, f)].detach().requires_grad_(True)
            mod.generate_images_pred(ins, outs, is_multi)
            losses, _ = mod.compute_losses(ins, outs, is_multi)
            losses["loss"].backward()
            if sl["free"] is None:
                sl["free"] = torch.cuda.Event()
            sl["free"].record(cur)
            return losses["loss"]

        loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event() for _ in range(2)]

        def e2e_run(host_sets, n):
            """n steps; the inputs of steps i+1 .. i+DEPTH are copied on the copy stream while step i computes; every
            step ends with the D2H of its loss, which the host reads one step later (as a training loop logs it)."""
            from collections import deque
            q = deque(start_h2d(host_sets[j % len(host_sets)], j) for j in range(min(DEPTH, n)))
            for i in range(n):
                cur = q.popleft()
                if i + DEPTH < n:
                    q.append(start_h2d(host_sets[(i + DEPTH) % len(host_sets)], i + DEPTH))
                loss = e2e_compute(cur)
                slot = i % 2
                loss_host[slot].copy_(loss.detach().reshape(1), non_blocking=True)
                loss_ev[slot].record()
                if i >= 1:
                    loss_ev[1 - slot].synchronize()
                    float(loss_host[1 - slot][0])
            loss_ev[(n - 1) % 2].synchronize()
            return float(loss_host[(n - 1) % 2][0])

        def time_e2e(host_sets, repeats=3, alone=False):
            """Ke steps, `repeats` times: the MEDIAN repetition is reported, every repetition is listed.  alone: this rank
            only, the others idle (no barriers, no reduction)."""
            e2e_run(host_sets, 3)
            ms = []
            for _ in range(repeats):
                if not alone:
                    barrier(world)
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                e2e_run(host_sets, Ke)
                f1.record()
                if alone:
                    torch.cuda.synchronize()
                    ms.append(f0.elapsed_time(f1) / Ke)
                else:
                    barrier(world)
                    ms.append(max_over_ranks(f0.elapsed_time(f1), world, device) / Ke)
            return statistics.median(ms), ms

        def leg(tensors):
            host_sets = [collate(t) for t in tensors]
            build_stage(host_sets)
            h2d = sum(v.numel() * v.element_size() for v in tensors[0].values())
            alone_ms = None
            if world > 1:
                # one rank with the box to itself, then all ranks together: the ratio names what the ranks share (host memory
                # and PCIe root complexes), with a number
                if rank == 0:
                    alone_ms, _ = time_e2e(host_sets, repeats=1, alone=True)
                barrier(world)
            ms, reps = time_e2e(host_sets)
            # the pipeline's arithmetic: one step on host set 0 (what loss_check evaluated through the resident API; the
            # device noise only breaks exact ties) must give that loss
            last = e2e_run(host_sets[:1], 1)
            if rank == 0 and loss_check is not None and not is_multi:
                assert abs(last - loss_check["api_loss"]) <= 1e-4 * abs(loss_check["api_loss"]), (last, loss_check)
            out = {"value": world * B * H * W / (ms * 1e-3) / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                   "ms_per_step": ms, "steps": Ke, "repeats_ms_per_step": reps, "h2d_gb_per_s": h2d / (ms * 1e-3) / 1e9,
                   "loss_of_set0_through_the_pipeline": last}
            if world > 1:
                out["h2d_gb_per_s_all_ranks"] = world * h2d / (ms * 1e-3) / 1e9
                if alone_ms is not None:
                    out["rank0_alone_ms_per_step"] = alone_ms
                    out["all_ranks_vs_rank0_alone"] = alone_ms / ms
            return out

        f32_leg = leg(step_tensors)
        # the dataset's frames are uint8 (ToTensor divides them by 255 on the CPU, mono_dataset.py:62,106): handed over as
        # uint8 planes and expanded on the device (ppea_images_u8_to_f32, bit-identical), a quarter of the image bytes cross PCIe
        u8_tensors = [{k: (torch.round(v * 255).to(torch.uint8) if (k[0] == "in" and k[1] == "color") else v) for k, v in t.items()} for t in step_tensors]
        e2e = leg(u8_tensors)
        e2e["frames"] = "uint8"
        e2e["float32_frames"] = f32_leg
        e2e["api"] = ("ppea_depth_b200.loss.ViewSynthesisLoss.generate_images_pred + compute_losses + backward (noise_mode=device, cached step "
                      "plans); colour frames cross PCIe as the dataset's uint8 planes and are expanded on the device (images_to_float); a "
                      "step's inputs sit in one pinned arena and cross as one copy into a ring of three device arenas (tensor views built "
                      "once); the next two steps are copied on a second stream while step i computes; the loss is copied to pinned host memory every step and read by the host one step later.  With "
                      "the reference's own CPU noise (noise_mode=reference: 4 torch.randn of (B,1,H,W) on the host + 23.6 MB H2D) a step "
                      "is bound by ~50 ms of host RNG instead.")

    # ---- N > 1: strong-scaling point of BASELINE configs[4] and the adapter-gradient all-reduce; N = 1: context numbers
    extras = {}
    if not args.no_extras and not args.shard and args.workload == "kitti":
        del plans
        torch.cuda.empty_cache()
        per = max(1, WORKLOADS["sweep96"]["batch"] // world)
        wl96 = dict(WORKLOADS["sweep96"], batch=per)
        sets96 = make_sets(wl96, 2, 5000 + 10 * rank, is_multi)
        plans96 = [build_plan(t, wl96, device, is_multi, args.deterministic, fused=fused_arg) for t in sets96]
        for p in plans96:
            p.capture()
        K96 = max(10, min(K, 50))
        ms96 = time_replays(plans96, K96, 5, world, device)
        extras["sharded_sweep96"] = {"global_batch": per * world, "batch_per_gpu": per, "ms_per_step": ms96, "steps": K96,
                                     "value": per * world * H * W / (ms96 * 1e-3) / 1e6, "unit": "Mpixels/s", "scaling": "strong",
                                     "note": "BASELINE configs[4]: global batch 96 at 192x640 split over the ranks, per-rank normalisation "
                                             "(what the reference's DDP does), no data-path collective"}
        if world > 1:
            buf = torch.zeros(ALLREDUCE_FLOATS, device=device, dtype=torch.float32)
            dist = torch.distributed

            def timed(fn, n):
                barrier(world)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(n):
                    fn(i)
                e1.record()
                barrier(world)
                return max_over_ranks(e0.elapsed_time(e1), world, device) / n

            for _ in range(3):
                dist.all_reduce(buf)
            ar_ms = timed(lambda i: dist.all_reduce(buf), 10)

            def overlapped(i):
                work = dist.all_reduce(buf, async_op=True)      # NCCL's own stream: runs beside the step's kernels
                plans96[i % 2].replay()
                work.wait()

            for i in range(3):
                overlapped(i)
            ov_ms = timed(overlapped, 10)
            nbytes = ALLREDUCE_FLOATS * 4
            extras["allreduce"] = {"bytes": nbytes, "allreduce_ms": ar_ms, "step_ms": ms96, "overlap_ms": ov_ms,
                                   "serial_ms": ar_ms + ms96, "bus_gb_per_s": 2.0 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9,
                                   "limiter": "the all-reduce (%.2f ms for 0.3 GB) is %s than the sharded step (%.2f ms): overlapped they take %.2f ms "
                                              "= %.0f %% of the longer one" % (ar_ms, "longer" if ar_ms > ms96 else "shorter", ms96, ov_ms,
                                                                                100.0 * ov_ms / max(ar_ms, ms96)),
                                   "what": "one flat-bucket NCCL all-reduce of a 0.3 GB fp32 stand-in for the adapter / decoder-adapter "
                                           "gradients (trainer.py:350, size from SURVEY.md §2.2) on NCCL's stream, beside the sharded step"}
            del buf
        del plans96
        if world == 1 and rank == 0:
            try:
                extras["torch_cuda_eager"] = time_torch_cuda_eager(WORKLOADS["kitti"], device)
            except Exception as exc:         # (context only: never fail the bench for it)
                extras["torch_cuda_eager"] = {"error": repr(exc)[:200]}
    t_clock1 = time.time()

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    clocks = sampler.stop(t_clock0, t_clock1)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        mpix, ms, cores, sample, kind = time_cpu_reference(wl, is_multi, steps=3, warmup=1, budget_s=20.0)
        cpu = {"value": mpix, "unit": "Mpixels/s", "cores": cores, "kind": kind, "sample": sample}
        mpix1, ms1, _, sample1, _ = time_cpu_reference(wl, is_multi, steps=2, warmup=1, budget_s=12.0, threads=1)
        cpu["threads_1"] = {"value": mpix1, "unit": "Mpixels/s", "cores": 1, "sample": sample1,
                            "note": "the reference pins OMP / MKL / NUMEXPR to one thread at import (trainer.py:8-10)"}

    line = {"metric": "view-synthesis loss fwd+bwd throughput", "value": value, "unit": "Mpixels/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.shard else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg_desc,
            "mpixel_scales_per_s": value * S, "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step,
            "kernels": kernels, "l2": l2_note, "host_cores_of_rank0": cores_mine,
            "clocks": clocks, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu, "loss_check": loss_check}
    line.update(extras)
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
