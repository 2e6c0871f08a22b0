#!/usr/bin/env python
"""bench.py -- view-synthesis loss fwd+bwd throughput (BASELINE.json metric) on N B200s.

One "step" = generate_images_pred + compute_losses + loss.backward() for one KITTI-shaped batch
(12x3x192x640, frame_ids [0,-1,1], 4 scales, mono path: automask + selec_reproj + pose gradients;
/root/reference/ppeadepth/trainer.py:871-918, 1032-1160), i.e. BASELINE.json configs[1].

  value     Mpixels/s = N_gpus * B*H*W * steps / time, inputs resident in HBM, forward+backward of the
            C ABI replayed as CUDA graphs, rotating over input sets larger than L2.
  e2e       the same metric through the public host API (ppea_depth_b200.loss.ViewSynthesisLoss) with
            pinned HOST inputs: H2D of every input, forward, backward, D2H of the loss, every step.
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event duration / measured HBM peak.
  cpu_baseline / --impl reference   the reference algorithm (oracle port, PyTorch CPU ops exactly as
            the reference calls them) timed on this box's host cores on a bounded sample.

Launch: `python bench.py` (1 GPU) or torchrun --nproc-per-node N bench.py --gpus N (weak scaling: every
rank owns its own 12-image batch; the loss path has no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS))
    ap.add_argument("--path", default="mono", choices=["mono", "multi"])
    ap.add_argument("--deterministic", action="store_true")
    ap.add_argument("--no-fused", action="store_true", help="forward + backward kernel pair instead of the single-launch step")
    ap.add_argument("--tiles", action="store_true", help="fused step by the shared-memory tile kernel (round 1) instead of the streaming kernel")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of one repetition of the e2e leg (default: min(steps, 100))")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


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
def make_sets(wl, n_sets, seed0, device, is_multi):
    """n_sets independent synthetic batches resident on `device` (+ their pinned host copies)."""
    from ppea_depth_b200.synth import SynthConfig, make_batch, make_noise
    sets = []
    for i in range(n_sets):
        cfg = SynthConfig(seed=seed0 + i, **wl)
        inputs, outputs = make_batch(cfg)
        noise = make_noise(cfg, wl["num_scales"])
        sets.append((inputs, outputs, noise))
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
    kw = {}
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


# ----------------------------------------------------------------------------- CPU reference arm
def time_cpu_reference(wl, is_multi, steps, warmup, budget_s=150.0):
    """The reference algorithm (oracle port = the same ATen CPU ops the reference dispatches) on the host
    cores: fwd+bwd on a bounded sample of the workload's batch."""
    from oracle import vsl_oracle as O
    from ppea_depth_b200.synth import SynthConfig, make_batch, make_noise
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
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
    sample = "%d of %d images per step (%dx%d, %d scales, %s path), %d steps after %d warm-up, mean %.3f s/step" % (
        n_items, wl["batch"], H, W, S, "multi" if is_multi else "mono", steps, warmup, mean_t)
    return mpix, mean_t * 1e3, cores, sample


def main():
    args = parse()
    rank, world, local = dist_env()
    wl = WORKLOADS[args.workload]
    is_multi = args.path == "multi"
    B, H, W, S = wl["batch"], wl["height"], wl["width"], wl["num_scales"]
    cfg_desc = {"workload": "%s %dx3x%dx%d, frame_ids [0,-1,1], %d scales, %s path (generate_images_pred + compute_losses + backward)"
                % (args.workload, B, H, W, S, args.path), "pixels_per_step_per_gpu": B * H * W, "scales": S,
                "path": args.path, "deterministic_backward": bool(args.deterministic)}

    if args.impl == "reference":
        if rank != 0:
            return
        mpix, ms, cores, sample = time_cpu_reference(wl, is_multi, max(1, args.steps), max(0, args.warmup))
        line = {"impl": "reference", "metric": "view-synthesis loss fwd+bwd throughput", "value": mpix, "unit": "Mpixels/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg_desc,
                "cpu_baseline": {"value": mpix, "unit": "Mpixels/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": mpix, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path for the product); "
                         "use --impl reference for the CPU reference arm")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # NCCL's version banner / warnings off stdout (rank 0 prints ONE JSON line): NCCL honours NCCL_DEBUG_FILE only above
        # the VERSION level, which is what this image sets
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        torch.distributed.init_process_group("nccl", device_id=device)
    from ppea_depth_b200 import _cabi
    from ppea_depth_b200.synth import algorithmic_bytes
    _cabi.lib()          # fail loudly if the extension is missing

    # ---- resident input sets, sized to exceed L2
    probe = make_sets(wl, 1, 1000 * rank, device, is_multi)
    step_bytes = sum(v.numel() * v.element_size() for v in tensors_of_step(*probe[0], S, is_multi).values())
    n_sets = max(4, int(2.5 * L2_BYTES // step_bytes) + 1)
    n_sets = min(n_sets, 12)
    sets = probe + make_sets(wl, n_sets - 1, 1000 * rank + 1, device, is_multi)
    plans = [build_plan(t, wl, device, is_multi, args.deterministic, fused=False if args.no_fused else ("tiles" if args.tiles else None)) for t in sets]
    fused = plans[0].fused
    tiles = plans[0].tiles
    cfg_desc["kernels"] = ("single-launch fused step (vsl_fused_kernel, TMA-staged tiles) + finish + gradient finish (tails launched programmatically)" if fused
                           else "vsl_forward_kernel + finish + vsl_backward_kernel + pose finish")
    for p in plans:
        p.capture()
    launches_per_step = plans[0].launches_forward + plans[0].launches_backward
    cfg_desc["l2"] = "rotating %d resident input sets (%.0f MB read per step, %.0f MB total) > 126 MB L2" % (
        n_sets, step_bytes / 1e6, n_sets * step_bytes / 1e6)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # ---- value: HBM-resident, graph replay
    K, Wm = max(1, args.steps), max(3, args.warmup)
    for i in range(Wm):
        plans[i % n_sets].replay()
    barrier(world)
    t_clock0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        plans[i % n_sets].replay()
    e1.record()
    barrier(world)
    ms_total = max_over_ranks(e0.elapsed_time(e1), world, device)
    ms_step = ms_total / K
    value = world * B * H * W / (ms_step * 1e-3) / 1e6

    # ---- roofline: per-kernel CUDA events (eager launches with the library's trace events), same rotation
    roof = None
    if rank == 0:
        acc = {}
        n_tr = min(K, 60)
        for p in plans:
            p.enable_trace()
        for i in range(n_tr):
            p = plans[i % n_sets]
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
        # SURVEY.md §8d per-pixel algorithmic bytes, split per launch: forward 49 + 16/4^s (mono) and backward 37 + 8/4^s,
        # multi: forward 53 + 16/4^s, backward 45 + 8/4^s; deterministic backward +8.
        fwd_b = sum((53.0 if is_multi else 49.0) + 16.0 / 4 ** s for s in range(S)) * n_px
        bwd_b = sum((45.0 if is_multi else 37.0) + (8.0 if args.deterministic else 0.0) + 8.0 / 4 ** s for s in range(S)) * n_px
        assert abs((fwd_b + bwd_b) - algorithmic_bytes(B, H, W, S, is_multi, args.deterministic)) < 1.0
        if fused:
            # The fused launch does the forward AND the backward work of every pixel.scale but has to move less than the
            # two-launch figure of SURVEY.md §8d: target/sources are read once, sel never round-trips.  Its own compulsory
            # bytes: tgt 12 + src 24 + noise 4 + depth 4 + sel 1 + loss 4 + (disp 4 + colour 12 + raw grad 4 + stencil 4)/4^s;
            # multi path: cons_mask 4 + mono_depth 4 instead of the noise, + consistency field 4/4^s.
            dom = "vsl_fused_kernel" if tiles else "vsl_stream_kernel"
            dom_bytes = sum((53.0 if is_multi else 49.0) + (28.0 if is_multi else 24.0) / 4 ** s for s in range(S)) * n_px
        else:
            dom = "vsl_backward_kernel" if stage_ms["vsl_backward_kernel"] >= stage_ms["vsl_forward_kernel"] else "vsl_forward_kernel"
            dom_bytes = bwd_b if dom == "vsl_backward_kernel" else fwd_b
        achieved = dom_bytes / (stage_ms[dom] * 1e-3) / 1e9
        traffic = None            # DRAM bytes of one launch of that kernel from the committed ncu capture (same workload only)
        tpath = os.path.join(ROOT, "profiles", "r1j_traffic.json" if fused else "r1f_traffic.json")
        if os.path.exists(tpath) and args.workload == "kitti" and not is_multi and not args.deterministic:
            traffic = json.load(open(tpath)).get(dom)
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "kernel_ms": stage_ms[dom], "stage_ms": stage_ms,
                "step_algorithmic_bytes": fwd_b + bwd_b,
                "kernel_frac_at_survey_step_bytes": (fwd_b + bwd_b) / (stage_ms[dom] * 1e-3) / 1e9 / peak if fused else None,
                "step_frac_of_peak": (fwd_b + bwd_b) / (ms_step * 1e-3) / 1e9 / peak,
                "frac_of_8TBs_nominal": achieved / 8000.0}

    # ---- e2e: public host API, pinned host inputs, H2D + fwd + bwd + D2H(loss) every step
    e2e = None
    if not args.no_e2e:
        from types import SimpleNamespace
        from ppea_depth_b200.loss import ViewSynthesisLoss
        opt = SimpleNamespace(sclm=S - 1, v1_multiscale=False, height=H, width=W, min_depth=0.1, max_depth=100.0,
                              frame_ids=[0, -1, 1], disable_automasking=False, no_ssim=False, selec_reproj=True,
                              disable_motion_masking=False, no_matching_augmentation=False, batch_size=B,
                              disparity_smoothness=1e-3)
        mod = ViewSynthesisLoss(opt, deterministic=args.deterministic, noise_mode="device", fused=False if args.no_fused else ("tiles" if args.tiles else None))
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
            step_tensors.append({k: v for k, v in t.items() if k[0] != "noise"})   # e2e draws the noise on the device (noise_mode="device")
        host_sets = [collate(t) for t in step_tensors]
        h2d = sum(v.numel() * v.element_size() for v in step_tensors[0].values())
        Ke = args.e2e_steps or min(K, 100)    # (the two-batch pipeline fill at the start is inside the timed region)

        copy_stream = torch.cuda.Stream(device=device)

        def start_h2d(hs):
            """Enqueues the step's host->device copy on the copy stream (as a pin_memory dataloader does ahead of
            the step); returns the device tensors (views of the device arena) and the event that marks their arrival."""
            arena, layout = hs
            with torch.cuda.stream(copy_stream):
                d_arena = arena.to(device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return (d_arena, layout), ev

        def e2e_compute(dev, ev):
            from ppea_depth_b200 import images_to_float
            d_arena, layout = dev
            torch.cuda.current_stream().wait_event(ev)
            d_arena.record_stream(torch.cuda.current_stream())
            dev = {k: d_arena[o:o + _nbytes(shape, dt)].view(dt).view(shape) for k, (o, shape, dt) in layout.items()}
            u8 = [k for k, (o, shape, dt) in layout.items() if dt == torch.uint8]
            if u8:
                # uint8 frames: one expansion call over their contiguous block (public API: images_to_float)
                lo = min(layout[k][0] for k in u8)
                hi = max(layout[k][0] + _nbytes(layout[k][1], layout[k][2]) for k in u8)
                f32 = images_to_float(d_arena[lo:hi])
                for k in u8:
                    o, shape, _ = layout[k]
                    dev[k] = f32[o - lo:o - lo + _nbytes(shape, torch.uint8)].view(shape)
            ins = {k[1:]: v for k, v in dev.items() if k[0] == "in"}
            outs = {(k[1] if len(k) == 2 else k[1:]): v for k, v in dev.items() if k[0] == "out"}
            for s in range(S):
                outs[("disp", s)].requires_grad_(True)
            if not is_multi:
                for f in (-1, 1):
                    outs[("cam_T_cam", 0, f)].requires_grad_(True)
            mod.generate_images_pred(ins, outs, is_multi)
            losses, _ = mod.compute_losses(ins, outs, is_multi)
            losses["loss"].backward()
            return losses["loss"]

        loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event() for _ in range(2)]
        DEPTH = 2     # batches in flight on the copy stream (a pin_memory DataLoader's prefetch_factor)

        def e2e_run(n):
            """n steps; the inputs of steps i+1 .. i+DEPTH are copied on the copy stream while step i computes (the copy
            engine never waits for the host: one step of host-side enqueue is ~0.6 ms, close to the copy time
            of a batch); every step ends with the D2H of its loss."""
            from collections import deque
            q = deque(start_h2d(host_sets[j % len(host_sets)]) for j in range(min(DEPTH, n)))
            last = 0.0
            for i in range(n):
                cur = q.popleft()
                if i + DEPTH < n:
                    q.append(start_h2d(host_sets[(i + DEPTH) % len(host_sets)]))
                loss = e2e_compute(*cur)
                # D2H of the step's result into pinned memory, every step; the host looks at it one step later (as a
                # training loop logs its loss), so the GPU is not left idle while the host enqueues the next step
                slot = i % 2
                loss_host[slot].copy_(loss.detach().reshape(1), non_blocking=True)
                loss_ev[slot].record()
                if i >= 1:
                    loss_ev[1 - slot].synchronize()
                    last = float(loss_host[1 - slot][0])
            loss_ev[(n - 1) % 2].synchronize()
            return float(loss_host[(n - 1) % 2][0])

        def time_e2e(repeats=3):
            """Ke steps, `repeats` times; the PCIe / host side of this leg is noisy on a shared box (single runs between
            1.22 and 1.85 ms/step were seen), so the best repetition is reported and every repetition is listed."""
            e2e_run(3)
            ms = []
            for _ in range(repeats):
                barrier(world)
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                e2e_run(Ke)
                f1.record()
                barrier(world)
                ms.append(max_over_ranks(f0.elapsed_time(f1), world, device) / Ke)
            return min(ms), ms

        ms_e2e, rep_f32 = time_e2e()
        e2e = {"value": world * B * H * W / (ms_e2e * 1e-3) / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e, "steps": Ke, "repeats_ms_per_step": rep_f32,
               "api": "ppea_depth_b200.loss.ViewSynthesisLoss.generate_images_pred + compute_losses + backward (noise_mode=device); "
                      "a step's inputs sit in one pinned arena and cross PCIe as one copy; the next two steps are copied on a second "
                      "stream while step i computes; the loss is copied "
                      "to pinned host memory every step and read by the host one step later"}
        # Same call, colour frames handed over as the dataset's uint8 planes and expanded on the device
        # (ppea_images_u8_to_f32, bit-identical to ToTensor): a quarter of the image bytes cross PCIe.  Reported beside the
        # reference-facing float32 number, not instead of it.
        for t in step_tensors:
            for k in list(t):
                if k[0] == "in" and k[1] == "color":
                    t[k] = torch.round(t[k] * 255).to(torch.uint8)
        host_sets = [collate(t) for t in step_tensors]
        h2d_u8 = sum(v.numel() * v.element_size() for v in step_tensors[0].values())
        ms_u8, rep_u8 = time_e2e()
        e2e["uint8_frames"] = {"value": world * B * H * W / (ms_u8 * 1e-3) / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": h2d_u8,
                               "d2h_bytes_per_step": 4, "ms_per_step": ms_u8, "steps": Ke, "repeats_ms_per_step": rep_u8}
    t_clock1 = time.time()

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    clocks = sampler.stop(t_clock0, t_clock1)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        mpix, ms, cores, sample = time_cpu_reference(wl, is_multi, steps=3, warmup=1, budget_s=20.0)
        cpu = {"value": mpix, "unit": "Mpixels/s", "cores": cores, "kind": "port", "sample": sample}

    line = {"metric": "view-synthesis loss fwd+bwd throughput", "value": value, "unit": "Mpixels/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg_desc,
            "mpixel_scales_per_s": value * S, "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step,
            "clocks": clocks, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
