#!/usr/bin/env python
"""bench.py -- LIC ray samples/s and fps of the 3D-LIC ray-cast hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl reference]

A "step" is one frame of the hot path (lic_raycast kernel + un-block/RGBA8 store) over the named synthetic
workload.  N = 1 default workload: cfg3 = 256^3 Crawfis tornado, cos^2 filter, length TF + gradient illumination,
LIC step 0.005, ray-cast step 1/128, 1024^2 (the configuration BASELINE.json's metric is quoted on).
N > 1 (torchrun, one rank per GPU): sort-first 16x16 image blocks dealt round-robin to ranks, volume replicated,
tile buffers all-gathered over NCCL; per-GPU work shrinks with N for a fixed frame => "strong" scaling.

`--impl reference` times the reference's own CPU implementation of the path (oracle/_ref when it was built from
/root/reference, else the oracle port) on the host cores, on a bounded pixel rectangle of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "lic_ray_samples_per_s"
UNIT = "ray samples/s"

# algorithmic gather bytes per ray sample, fp16x4 vector storage, S = 64:
#   (1+2S)*8*8 vector + (1+S)*8*B_noise + n_scalar*8 + (1+S)*2 kernel + 12 TF
B_GRAD = (1 + 2 * 64) * 8 * 8 + 65 * 8 * 4 + 8 + 130 + 12      # RGBA noise, gradient build  = 10486
B_SCALAR = (1 + 2 * 64) * 8 * 8 + 65 * 8 * 1 + 66 * 8 + 130 + 12  # scalar noise + band gate    = 9446


def make_scene(name):
    from vectorvisualization_b200 import configs
    mk = {"cfg1": configs.cfg1, "cfg2": configs.cfg2, "cfg3": configs.cfg3, "cfg4": configs.cfg4}[name]
    return mk()


def workload_string(s):
    n = s.field.shape
    return "%s: %dx%dx%d field, %d^3 noise%s, %dx%d view, raycast step 1/%d, LIC %d+%d steps h=%g" % (
        s.name, n[2], n[1], n[0], s.noise.shape[0], " +gradients" if s.with_gradients else "", s.width, s.height,
        round(1.0 / s.lic_params().stepSizeVol), s.lic_params().stepsForward, s.lic_params().stepsBackward, s.lic_params().stepSizeLIC)


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU reference arm

def make_cpu_runner(scene, use_ref=True):
    """the CPU implementation of the path: the reference's shader code compiled as C++ (oracle/_ref, kind "reference")
    when it was built from /root/reference, else the oracle port (kind "port")"""
    from oracle import vvo
    if use_ref:
        try:
            from oracle import refshim
            if refshim.available():
                return refshim.RefScene(scene), "reference", vvo.lib().vvo_num_threads()
        except Exception:
            pass
    return vvo.OracleScene(scene), "port", vvo.lib().vvo_num_threads()


def centred_rect(scene, side):
    w, h = scene.width, scene.height
    side = max(8, min(int(side), min(w, h)))
    x0, y0 = (w - side) // 2, (h - side) // 2
    return (x0, y0, x0 + side, y0 + side)


def cpu_rate(scene, runner, kind, cores, budget_s, side=None):
    """ray samples/s of the CPU implementation on a centred pixel rectangle sized for ~budget_s seconds"""
    if side is None:
        t = time.perf_counter()
        _, _, n = runner.raycast(centred_rect(scene, 24))
        dt = time.perf_counter() - t
        rate = max(n, 1) / max(dt, 1e-6)
        per_px = max(n, 1) / (24.0 * 24.0)
        side = (budget_s * rate / per_px) ** 0.5
    rc = centred_rect(scene, side)
    t = time.perf_counter()
    _, _, n = runner.raycast(rc)
    dt = time.perf_counter() - t
    return dict(value=n / dt, cores=cores, kind=kind, seconds=dt, samples=int(n), side=rc[2] - rc[0],
                sample="%s, centred %dx%d-pixel rectangle of the %dx%d frame (%d ray samples, %.1f s, %s)" % (
                    scene.name, rc[2] - rc[0], rc[3] - rc[1], scene.width, scene.height, n, dt,
                    "reference shader code compiled as C++ (oracle/_ref)" if kind == "reference" else "fp32 C++ port (oracle/)"))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scene = make_scene(args.config)
    runner, kind, cores = make_cpu_runner(scene)
    total_budget = 150.0
    per_step = max(1.0, min(12.0, total_budget / max(1, args.steps + args.warmup)))
    info = cpu_rate(scene, runner, kind, cores, per_step)          # sizes the rectangle (counts as the first warm-up)
    side = info["side"]
    for _ in range(max(0, args.warmup - 1)):
        cpu_rate(scene, runner, kind, cores, per_step, side)
    secs, samples = 0.0, 0
    for _ in range(args.steps):
        info = cpu_rate(scene, runner, kind, cores, per_step, side)
        secs += info["seconds"]; samples += info["samples"]
    value = samples / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_string(scene), "timing": "bounded pixel rectangle per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ CUDA arm

def run_cuda(args):
    import torch
    import torch.distributed as dist
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs
    from vectorvisualization_b200.dist import render_distributed, connect_p2p, render_distributed_p2p, disconnect_p2p

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    scene = make_scene(args.config)
    r = vv.Renderer(local_rank)
    configs.apply_scene(r, scene)
    stream = torch.cuda.Stream()          # a real (non-legacy) stream: the library, NCCL and the timing events share it
    torch.cuda.set_stream(stream)
    r.setStream(stream.cuda_stream)
    exchange = "none"
    if world > 1:
        r.setPartition(rank, world)
        # tile exchange: peer-to-peer stores over NVLink (vv_p2p_*), or one NCCL all_gather per frame.  Every rank must
        # take the same path, so the outcome of the IPC set-up is agreed on first.
        ok = 0
        if args.exchange == "p2p":
            try:
                connect_p2p(r)
                ok = 1
            except Exception as ex:
                print("bench.py: rank %d: peer-to-peer set-up failed (%r), falling back to the NCCL gather" % (rank, ex), file=sys.stderr)
        t = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        exchange = "p2p" if int(t.item()) == 1 else "nccl"

    def frame():
        if exchange == "p2p":
            render_distributed_p2p(r)
            return None
        if world > 1:
            return render_distributed(r)
        r.render(True)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    keep = None
    for _ in range(max(args.warmup, 3)):
        keep = frame()
    barrier()
    if exchange == "p2p":
        # the warm-up frames went through the peer-to-peer exchange: if any rank saw a time-out, every rank drops to NCCL
        ok = 1
        try:
            r.p2pStatus()
        except Exception as ex:
            ok = 0
            print("bench.py: rank %d: %r, falling back to the NCCL gather" % (rank, ex), file=sys.stderr)
        t = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) == 0:
            torch.cuda.synchronize()
            dist.barrier()
            r.p2pDisconnect()
            exchange = "nccl"
            for _ in range(3):
                keep = frame()
            barrier()
    samples_per_frame = r.lastRaySamples()
    launches_per_frame = r.lastLaunchCount()
    if world > 1:
        t = torch.tensor([samples_per_frame], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        samples_per_frame = int(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- device-timed region: K frames, inputs resident in HBM ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        keep = frame()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    # dominant-kernel duration, measured live with CUDA events on the launching stream (the library brackets the
    # lic_raycast launch with its own event pair every frame)
    for _ in range(min(args.steps, 10)):
        frame()
        torch.cuda.synchronize()
        kernel_ms.append(r.lastKernelMs())
    barrier()
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = samples_per_frame * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the C ABI with host buffers: camera/params in, RGBA8 frame out, every step ----
    host = torch.empty((scene.height, scene.width, 4), dtype=torch.uint8).pin_memory()
    host_np = host.numpy()
    cam = dict(scene.camera)
    lp = scene.lic_params()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        r.setCamera(**cam)               # host -> device: the frame's parameter block (kernel arguments)
        r.setLICParams(lp)
        keep = frame()
        if rank == 0:
            r.readRGBA8(host_np)         # device -> host: the stored RGBA8 frame (what saveTexture writes)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    e2e_value = samples_per_frame * args.steps / e2e_s
    h2d = 1024          # sizeof(DevParams) kernel-argument block (< 1 KiB) per frame
    d2h = scene.width * scene.height * 4

    if exchange == "p2p":
        disconnect_p2p(r)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    B = B_GRAD if scene.with_gradients and "ILLUM_GRADIENT" in scene.defines else B_SCALAR
    k_ms = float(np.median(kernel_ms)) if kernel_ms else ms_per_step
    # samples processed by THIS rank's launch (rank 0) for the per-launch roofline
    local_samples = r.lastRaySamples()
    achieved = B * local_samples / (k_ms * 1e-3) / 1e9
    n = scene.field.shape
    compulsory = n[0] * n[1] * n[2] * 16 + scene.noise.size * (16 if scene.with_gradients else 8) + scene.width * scene.height * 16
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_ray_sample": B, "kernel": "lic_sample_kernel",
                "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
                "note": "requested gather bytes (L1/L2-served, reuse makes frac > 1 legitimate); compulsory HBM bytes/frame = %d" % compulsory,
                "compulsory_hbm_frac": compulsory / (k_ms * 1e-3) / 1e9 / peak}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            entry = json.load(open(prof)).get(scene.name) or {}
            roofline["traffic"] = entry.get("bytes_per_launch")
            if entry.get("ncu"):
                # the kernel is not byte-bound: what it IS bound by, from the committed ncu capture of this kernel
                roofline["ncu"] = entry["ncu"]
                loop = entry["ncu"].get("sass_walk_loop")
                if loop and local_samples > 0:
                    # instruction roofline of the walk loop: every SM sub-partition issues at most one warp instruction per
                    # clock, so a frame cannot take less than (warp double steps) x (instructions per double step) cycles
                    steps = max(lp.stepsForward, lp.stepsBackward)
                    sms, mhz = 148, 1965.0
                    try:
                        mhz = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("sm_max_mhz", mhz))
                        sms = torch.cuda.get_device_properties(0).multi_processor_count
                    except Exception:
                        pass
                    ideal_ms = (local_samples / 32.0) * steps * loop["executed_instructions_per_double_step"] / (sms * 4) / (mhz * 1e3)
                    roofline["issue_roofline"] = {"ideal_kernel_ms": ideal_ms, "frac": ideal_ms / k_ms, "sm_mhz": mhz,
                                                  "what": "walk-loop warp instructions / (4 issue slots per SM per clock)"}
        except Exception:
            pass

    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            runner, kind, cores = make_cpu_runner(scene)
            c = cpu_rate(scene, runner, kind, cores, 15.0)
            cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"], "sample": c["sample"]}
        except Exception as ex:   # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "fps": 1e3 / ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(scene), "ray_samples_per_frame": samples_per_frame,
                   "l2": "inputs larger than L2 (field %.0f MB + noise %.0f MB vs 126 MB)" % (
                       n[0] * n[1] * n[2] * 16 / 1e6, scene.noise.size * (16 if scene.with_gradients else 8) / 1e6),
                   "parallelism": "sort-first 16x16 blocks over %d GPU(s), volume replicated" % world,
                   "exchange": {"none": "single GPU", "p2p": "peer-to-peer tile stores over NVLink + arrival counters (vv_p2p_render)",
                                "nccl": "NCCL all_gather_into_tensor of the tile buffers"}[exchange],
                   "field_layout": "x-pair fp16 (16 B/voxel)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "fps": args.steps / e2e_s, "what": "vv_set_camera + vv_set_lic_params + vv_render + vv_read_rgba8 into pinned host memory, per frame"},
        "gpu_launches": launches_per_frame * args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": sampler.summary() if sampler else None,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default=None)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="multi-GPU tile exchange (N > 1)")
    args = ap.parse_args()
    if args.config is None:
        args.config = "cfg3"
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
