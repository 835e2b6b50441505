#!/usr/bin/env python
"""bench.py -- LIC ray samples/s and fps of the 3D-LIC ray-cast hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl reference]

A "step" is one frame of the hot path (ray set-up, lic_sample_kernel, compositing, un-block / RGBA8 store) over the named
synthetic workload.  N = 1 default workload: cfg3 = 256^3 Crawfis tornado, cos^2 filter, length TF + gradient illumination,
LIC step 0.005, ray-cast step 1/128, 1024^2 (the configuration BASELINE.json's metric is quoted on).
N > 1 (torchrun, one rank per GPU): sort-first 16x16 image blocks dealt round-robin to ranks, volume replicated, tile
buffers exchanged by peer-to-peer stores over NVLink (or one NCCL all_gather); a fixed frame is split => "strong" scaling.

Other workloads: --config cfg1 | cfg2 | cfg4 (the remaining ray-cast configurations of BASELINE.json), cfg3o (cfg3 with the
reference's default, opaque transfer function: the early-ray-termination path), cfg5 (LIC-volume precompute in z-slabs over
the ranks + all-gather + plain ray-cast; --cfg5-n / --cfg5-size scale it down).  A cfg3 run also carries `extra.cfg4`
(3 frames of the 512^3 / 2048^2 scaling configuration at the run's N).

`--impl reference` times the reference's own CPU implementation of the path (oracle/_ref when it was built from
/root/reference, else the oracle port) on ALL host cores, on a bounded pixel rectangle of the same workload.
"""
import argparse
import ctypes
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
B_VOLRAY = 8 * 8 + 8 * 4 + 12                                   # volume ray-cast: fp16x4 field + fp32 LIC volume + TF = 108


def make_scene(name, args=None):
    from vectorvisualization_b200 import configs
    if name == "cfg3o":
        return configs.cfg3o()
    if name == "cfg5":
        return configs.cfg5(n=args.cfg5_n, size=args.cfg5_size)
    mk = {"cfg1": configs.cfg1, "cfg2": configs.cfg2, "cfg3": configs.cfg3, "cfg4": configs.cfg4}[name]
    return mk()


def workload_string(s):
    n = s.field.shape
    return "%s: %dx%dx%d field, %d^3 noise%s, %dx%d view, raycast step 1/%d, LIC %d+%d steps h=%g" % (
        s.name, n[2], n[1], n[0], s.noise.shape[0], " +gradients" if s.with_gradients else "", s.width, s.height,
        round(1.0 / s.lic_params().stepSizeVol), s.lic_params().stepsForward, s.lic_params().stepsBackward, s.lic_params().stepSizeLIC)


def config_dict(s):
    """the workload definition -- identical in the CUDA arm and the reference arm"""
    n = s.field.shape
    return {"workload": workload_string(s),
            "l2": "inputs larger than L2 (field %.0f MB at 16 B/voxel, twice that in the default xy-quad layout, + noise %.0f MB vs 126 MB)" % (
                n[0] * n[1] * n[2] * 16 / 1e6, s.noise.size * (16 if s.with_gradients else 8) / 1e6)}


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


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def profile_record(scene_name):
    """what the committed ncu captures say about the dominant kernel on this workload (profiles/traffic.json)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(scene_name) or {}
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------------ CPU reference arm

def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def force_all_cores():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must not inherit that (round-1 SCALE ratios were void)"""
    n = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    return n


def make_cpu_runner(scene, use_ref=True):
    """the CPU implementation of the path: the reference's shader code compiled as C++ (oracle/_ref, kind "reference")
    when it was built from /root/reference, else the oracle port (kind "port"); always on all host cores"""
    n = force_all_cores()
    from oracle import vvo
    lib = vvo.lib()
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)    # the OpenMP runtime both CPU libraries link
    except Exception:
        pass
    if use_ref:
        try:
            from oracle import refshim
            if refshim.available():
                return refshim.RefScene(scene), "reference", lib.vvo_num_threads()
        except Exception:
            pass
    return vvo.OracleScene(scene), "port", lib.vvo_num_threads()


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
    img, cnt, n = runner.raycast(rc)
    dt = time.perf_counter() - t
    return dict(value=n / dt, cores=cores, kind=kind, seconds=dt, samples=int(n), side=rc[2] - rc[0], rect=rc, image=img, counts=cnt,
                sample="%s, centred %dx%d-pixel rectangle of the %dx%d frame (%d ray samples, %.1f s, %d threads, %s)" % (
                    scene.name, rc[2] - rc[0], rc[3] - rc[1], scene.width, scene.height, n, dt, cores,
                    "reference shader code compiled as C++ (oracle/_ref)" if kind == "reference" else "fp32 C++ port (oracle/)"))


def parity_stamp(rect, cpu_img, cpu_cnt, gpu_img, gpu_cnt):
    """the frame that was timed against the CPU implementation's frame, inside the rectangle the CPU leg shaded"""
    from oracle import vvo
    x0, y0, x1, y1 = rect
    a = vvo.quantize_rgba8(np.ascontiguousarray(gpu_img[y0:y1, x0:x1])).astype(np.int32)
    b = vvo.quantize_rgba8(np.ascontiguousarray(cpu_img[y0:y1, x0:x1])).astype(np.int32)
    mse = float(np.mean((a - b) ** 2))
    out = {"max_diff_8bit": int(np.abs(a - b).max()), "psnr_db": None if mse == 0 else round(10 * np.log10(255.0 ** 2 / mse), 2),
           "identical_8bit": bool(mse == 0), "pixels": int((x1 - x0) * (y1 - y0)), "rect": [int(v) for v in rect],
           "tolerance": "<= 2/255 per channel, PSNR >= 45 dB (BASELINE.json)"}
    if gpu_cnt is not None:
        out["samples_equal"] = bool(np.array_equal(gpu_cnt[y0:y1, x0:x1], cpu_cnt[y0:y1, x0:x1]))
        out["sample_count_mismatches"] = int((gpu_cnt[y0:y1, x0:x1] != cpu_cnt[y0:y1, x0:x1]).sum())
    out["within_tolerance"] = bool(out["max_diff_8bit"] <= 2 and (out["psnr_db"] is None or out["psnr_db"] >= 45.0))
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    force_all_cores()
    scene = make_scene(args.config, args)
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
        "data": "synthetic", "config": config_dict(scene),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ CUDA arm

class Job:
    """one process per GPU: renderer, stream, partition and tile exchange of a ray-cast workload"""

    def __init__(self, scene, args):
        import torch
        import torch.distributed as dist
        import vectorvisualization_b200 as vv
        from vectorvisualization_b200 import configs
        from vectorvisualization_b200.dist import connect_p2p
        self.torch, self.dist, self.scene = torch, dist, scene
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.r = vv.Renderer(self.local_rank)
        configs.apply_scene(self.r, scene)
        self.stream = torch.cuda.current_stream()
        self.r.setStream(self.stream.cuda_stream)
        self.exchange = "none"
        if self.world > 1:
            self.r.setPartition(self.rank, self.world)
            # tile exchange: peer-to-peer stores over NVLink (vv_p2p_*), or one NCCL all_gather per frame.  Every rank must
            # take the same path, so the outcome of the IPC set-up is agreed on first.
            ok = 0
            if args.exchange == "p2p":
                try:
                    connect_p2p(self.r)
                    ok = 1
                except Exception as ex:
                    print("bench.py: rank %d: peer-to-peer set-up failed (%r), falling back to the NCCL gather" % (self.rank, ex), file=sys.stderr)
            self.exchange = "p2p" if self.all_min(ok) == 1 else "nccl"
        self.keep = None

    def all_min(self, v):
        t = self.torch.tensor([v], dtype=self.torch.int32, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t.item())

    def all_max(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_sum(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.int64, device="cuda")
        self.dist.all_reduce(t)
        return int(t.item())

    def frame(self):
        from vectorvisualization_b200.dist import render_distributed, render_distributed_p2p
        if self.exchange == "p2p":
            render_distributed_p2p(self.r)
        elif self.world > 1:
            self.keep = render_distributed(self.r)
        else:
            self.r.render(True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def warm(self, n):
        for _ in range(n):
            self.frame()
        self.barrier()
        if self.exchange == "p2p":
            # the warm-up frames went through the peer-to-peer exchange: if any rank saw a time-out, every rank drops to NCCL
            ok = 1
            try:
                self.r.p2pStatus()
            except Exception as ex:
                ok = 0
                print("bench.py: rank %d: %r, falling back to the NCCL gather" % (self.rank, ex), file=sys.stderr)
            if self.all_min(ok) == 0:
                self.torch.cuda.synchronize()
                self.dist.barrier()
                self.r.p2pDisconnect()
                self.exchange = "nccl"
                for _ in range(3):
                    self.frame()
                self.barrier()

    def timed(self, steps):
        """K frames between two CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks"""
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        for _ in range(steps):
            self.frame()
        e1.record(self.stream)
        self.barrier()
        return self.all_max(e0.elapsed_time(e1))

    def close(self):
        from vectorvisualization_b200.dist import disconnect_p2p
        if self.exchange == "p2p":
            disconnect_p2p(self.r)
        self.keep = None
        self.r.close()


def time_extra_cfg4(args, frames=3):
    """the scaling configuration of the north star (512^3 / 2048^2) at this run's N: `frames` frames after one warm-up"""
    t0 = time.perf_counter()
    scene = make_scene("cfg4", args)
    job = Job(scene, args)
    job.warm(2)
    samples = job.all_sum(job.r.lastRaySamples())
    ms = job.timed(frames) / frames
    k_ms = job.r.lastKernelMs()
    rec = {"workload": workload_string(scene), "frames": frames, "ms_per_frame": ms, "fps": 1e3 / ms, "ray_samples_per_frame": samples,
           "value": samples / (ms * 1e-3), "unit": UNIT, "n_gpus": job.world, "exchange": job.exchange, "rank0_kernel_ms": k_ms,
           "setup_s": None}
    job.close()
    rec["setup_s"] = round(time.perf_counter() - t0, 1)
    return rec


def run_cuda(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_stream(torch.cuda.Stream())   # a real (non-legacy) stream: the library, NCCL and the timing events share it
    if args.config == "cfg5":
        rc = run_cfg5(args)
        if world > 1:
            dist.destroy_process_group()
        return rc
    scene = make_scene(args.config, args)
    job = Job(scene, args)
    r = job.r
    warm = max(args.warmup, 3)
    job.warm(warm)
    local_samples = r.lastRaySamples()
    launches_per_frame = r.lastLaunchCount()
    field_layout = {0: "float4 (16 B/voxel)", 1: "x-pair fp16 (16 B/voxel)", 2: "xy-quad fp16 (32 B/voxel, one 256-bit load per cell face)"}.get(r.fieldLayout())
    samples_per_frame = job.all_sum(local_samples)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- device-timed region: K frames, inputs resident in HBM ----
    ms_total = job.timed(args.steps)
    ms_per_step = ms_total / args.steps
    value = samples_per_frame * args.steps / (ms_total * 1e-3)
    # a longer back-to-back run when K is small (sustained clocks / power): reported beside the K-step number, never instead of it
    sustained = None
    if args.steps < 100 and not args.no_extra:
        n_sus = max(100, min(400, int(3000.0 / max(ms_per_step, 0.1))))
        ms_sus = job.timed(n_sus) / n_sus
        sustained = {"steps": n_sus, "ms_per_step": ms_sus, "value": samples_per_frame / (ms_sus * 1e-3)}
    # dominant-kernel duration, measured live with CUDA events on the launching stream (the library brackets the
    # lic_sample launch of the first depth window with its own event pair every frame)
    kernel_ms = []
    for _ in range(min(args.steps, 10)):
        job.frame()
        torch.cuda.synchronize()
        kernel_ms.append(r.lastKernelMs())
    job.barrier()

    # ---- end-to-end through the C ABI with host buffers: camera/params in, RGBA8 frame out, every step ----
    host = torch.empty((scene.height, scene.width, 4), dtype=torch.uint8).pin_memory()
    host_np = host.numpy()
    cam = dict(scene.camera)
    lp = scene.lic_params()
    job.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        r.setCamera(**cam)               # host -> device: the frame's parameter block (kernel arguments)
        r.setLICParams(lp)
        job.frame()
        if rank == 0:
            r.readRGBA8(host_np)         # device -> host: the stored RGBA8 frame (what saveTexture writes)
    job.barrier()
    e2e_s = job.all_max(time.perf_counter() - t0)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    e2e_value = samples_per_frame * args.steps / e2e_s
    h2d = 1024          # sizeof(DevParams) kernel-argument block (< 1 KiB) per frame
    d2h = scene.width * scene.height * 4

    # the frame that was timed, for the parity stamp (rank 0 holds the assembled frame in every exchange mode)
    gpu_img = gpu_cnt = None
    if rank == 0 and not args.no_cpu:
        gpu_img = r.readRGBA32F()
        if world == 1:
            r.setOption(__import__("vectorvisualization_b200").OPT_SAMPLE_MAP, 1)
            r.render(True)
            gpu_cnt = r.readSampleMap()
    exchange = job.exchange
    job.close()

    extra = {}
    if args.config == "cfg3" and not args.no_extra:
        try:
            extra["cfg4"] = time_extra_cfg4(args)
        except Exception as ex:   # reported, never required for the headline number
            extra["cfg4"] = {"error": repr(ex)[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src, sm_mhz = measured_peaks()
    grad = scene.with_gradients and "ILLUM_GRADIENT" in scene.defines
    B = B_GRAD if grad else B_SCALAR
    k_ms = float(np.median(kernel_ms)) if kernel_ms else ms_per_step
    requested = B * local_samples / (k_ms * 1e-3) / 1e9            # GB/s asked for by the lanes of THIS rank's launch
    n = scene.field.shape
    field_bytes_per_voxel = 32 if field_layout and field_layout.startswith("xy-quad") else 16
    compulsory = n[0] * n[1] * n[2] * field_bytes_per_voxel + scene.noise.size * (16 if scene.with_gradients else 8) + scene.width * scene.height * 16
    prof = profile_record(scene.name)
    ncu = prof.get("ncu") or {}
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    issue_peak = sms * 4 * sm_mhz * 1e6 / 1e9                       # G warp instructions / s: one per sub-partition per clock
    roofline = {
        # What bounds lic_sample_kernel is instruction issue / the FMA pipe, not bytes (the same frame over a 32^3 volume takes
        # the same time, profiles/r02/): `achieved` = executed warp instructions (per ray sample from the committed ncu
        # capture of this build x this launch's ray samples) / the live kernel duration; `peak` = 4 issue slots per SM per clock.
        "bound": "issue/fma",
        "achieved": None, "peak": issue_peak, "unit": "G warp-instr/s", "frac": None,
        "kernel": "lic_sample_kernel", "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
        "traffic": prof.get("bytes_per_launch"),
        "hbm": {"requested_GBps": requested, "peak_GBps": peak, "peak_source": peak_src, "requested_bytes_frac": requested / peak,
                "algorithmic_bytes_per_ray_sample": B, "compulsory_bytes_per_frame": compulsory,
                "compulsory_hbm_frac": compulsory / (k_ms * 1e-3) / 1e9 / peak,
                "note": "SURVEY 8(d) figure: requested gather bytes are L1/L2-served (reuse makes the fraction > 1 legitimately)"},
    }
    ipr = ncu.get("warp_instructions_per_ray_sample")
    if ipr:
        roofline["achieved"] = ipr * local_samples / (k_ms * 1e-3) / 1e9
        roofline["frac"] = roofline["achieved"] / issue_peak
        roofline["warp_instructions_per_ray_sample"] = ipr
    gp = prof.get("gather_peaks") or profile_record("gather_peaks")
    if gp:
        roofline["l1_gather_peak_GBps"] = gp.get("l1_trilinear_GBps")
        roofline["l2_gather_peak_GBps"] = gp.get("l2_trilinear_GBps")
        roofline["hbm_gather_peak_GBps"] = gp.get("hbm_trilinear_GBps")
        if gp.get("l1_trilinear_GBps"):
            roofline["frac_of_l1_gather_peak"] = requested / gp["l1_trilinear_GBps"]
        if gp.get("l2_trilinear_GBps"):
            roofline["frac_of_l2_gather_peak"] = requested / gp["l2_trilinear_GBps"]
        roofline["gather_peaks_source"] = gp.get("source")
    if ncu:
        roofline["ncu"] = ncu

    cpu = parity = None
    if world == 1 and not args.no_cpu:
        try:
            runner, kind, cores = make_cpu_runner(scene)
            c = cpu_rate(scene, runner, kind, cores, 15.0)
            cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"], "sample": c["sample"]}
            parity = parity_stamp(c["rect"], c["image"], c["counts"], gpu_img, gpu_cnt)
            parity["against"] = c["kind"]
        except Exception as ex:   # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "fps": 1e3 / ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(scene),
        "run": {"ray_samples_per_frame": samples_per_frame,
                "parallelism": "sort-first 16x16 blocks over %d GPU(s), volume replicated" % world,
                "exchange": {"none": "single GPU", "p2p": "peer-to-peer tile stores over NVLink + arrival counters (vv_p2p_render)",
                             "nccl": "NCCL all_gather_into_tensor of the tile buffers"}[exchange],
                "field_layout": field_layout, "launches_per_frame": launches_per_frame},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "fps": args.steps / e2e_s, "what": "vv_set_camera + vv_set_lic_params + vv_render + vv_read_rgba8 into pinned host memory, per frame"},
        "gpu_launches": launches_per_frame * args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "clocks": sampler.summary() if sampler else None,
    }
    if sustained:
        line["sustained"] = sustained
    if extra:
        line["extra"] = extra
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------ cfg5: LIC-volume mode

def run_cfg5(args):
    """BASELINE.json config 5: per-voxel LIC volume in z-slabs over the ranks (input replicated), all-gather of the
    slabs over NCCL, then the plain ray-cast of the LIC volume (sort-first tiles as in the other configurations).
    A step = one LIC-volume update + one ray-cast frame; the metric stays the ray-cast's ray samples/s over the whole step,
    the LIC-volume stage is reported beside it (voxels/s, s per volume)."""
    import torch
    import torch.distributed as dist
    import vectorvisualization_b200 as vv
    from vectorvisualization_b200 import configs
    from vectorvisualization_b200.dist import render_distributed, update_lic_volume_distributed

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    t0 = time.perf_counter()
    scene = make_scene("cfg5", args)
    r = vv.Renderer(local_rank)
    configs.apply_scene(r, scene)
    stream = torch.cuda.current_stream()
    r.setStream(stream.cuda_stream)
    if world > 1:
        r.setPartition(rank, world)
    depth = scene.field.shape[0]
    setup_s = time.perf_counter() - t0
    keep = [None]

    def volume():
        if world > 1:
            update_lic_volume_distributed(r, depth)
        else:
            r.updateLICVolume()

    def frame():
        if world > 1:
            keep[0] = render_distributed(r)
        else:
            r.render(True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def amax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    steps, warm = max(1, args.steps), max(1, min(args.warmup, 2))
    for _ in range(warm):
        volume(); frame()
    barrier()
    samples = r.lastRaySamples()
    if world > 1:
        t = torch.tensor([samples], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        samples = int(t.item())
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    vol_ms = ray_ms = 0.0
    k_vol = []
    barrier()
    for _ in range(steps):
        ev[0].record(stream)
        volume()
        ev[1].record(stream)
        frame()
        ev[2].record(stream)
        torch.cuda.synchronize()
        vol_ms += ev[0].elapsed_time(ev[1]); ray_ms += ev[1].elapsed_time(ev[2])
    barrier()
    vol_ms = amax(vol_ms / steps); ray_ms = amax(ray_ms / steps)
    # the slab kernel alone (library events around lic_volume_kernel), rank 0's slab
    r.setLICVolumeSlab(*__import__("vectorvisualization_b200.dist", fromlist=["slab_range"]).slab_range(depth, rank, world))
    r.updateLICVolume(); torch.cuda.synchronize()
    k_vol_ms = amax(r.lastKernelMs())
    # the volume the ranks assembled from their z-slabs against the whole volume computed on one GPU, at full size: rank 0 keeps
    # a copy of the gathered volume, computes all slabs itself into the same buffer and compares on the device
    identical = None
    if world > 1:
        volume(); torch.cuda.synchronize()
        if rank == 0:
            ptr, dims = r.licVolumePtr()
            from vectorvisualization_b200.dist import device_tensor
            vol = device_tensor(ptr, (dims[2] * dims[1] * dims[0],), torch.float32)
            gathered = vol.clone()
            r.setLICVolumeSlab(0, depth)
            r.updateLICVolume(); torch.cuda.synchronize()
            identical = bool(torch.equal(vol, gathered))
            del gathered
        dist.barrier()
    if rank != 0:
        return 0
    peak, peak_src, _ = measured_peaks()
    nvox = depth * scene.field.shape[1] * scene.field.shape[2]
    step_ms = vol_ms + ray_ms
    B = B_SCALAR - 140
    line = {
        "metric": METRIC, "value": samples / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg5: %d^3 field -> %d^3 LIC volume in %d z-slab(s), all-gather, volume ray-cast %dx%d step 1/256" % (
            depth, depth, world, scene.width, scene.height)},
        "lic_volume": {"ms_per_volume": vol_ms, "voxels_per_s": nvox / (vol_ms * 1e-3), "lic_taps_per_s": nvox * 65 / (vol_ms * 1e-3),
                       "slab_kernel_ms_max_over_ranks": k_vol_ms, "voxels": nvox,
                       "gathered_slabs_bit_identical_to_one_gpu": identical,
                       "includes": "slab kernel on every rank + NCCL all-gather of the fp32 slabs" if world > 1 else "kernel only (one GPU)"},
        "volume_raycast": {"ms_per_frame": ray_ms, "ray_samples_per_frame": samples, "ray_samples_per_s": samples / (ray_ms * 1e-3)},
        "roofline": {"bound": "hbm", "kernel": "lic_volume_kernel", "achieved": B * (nvox / world) / (k_vol_ms * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": B * (nvox / world) / (k_vol_ms * 1e-3) / 1e9 / peak, "traffic": profile_record("cfg5").get("bytes_per_launch"),
                     "peak_source": peak_src, "algorithmic_bytes_per_voxel": B,
                     "note": "requested gather bytes per voxel (SURVEY 8(d): ray-cast figure minus the ray-side fetches); cache-served"},
        "gpu_launches": r.lastLaunchCount() * steps, "setup_s": round(setup_s, 1),
        "e2e": None, "cpu_baseline": None,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="cfg3", choices=["cfg1", "cfg2", "cfg3", "cfg3o", "cfg4", "cfg5"])
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra.cfg4 record and the sustained run")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="multi-GPU tile exchange (N > 1)")
    ap.add_argument("--cfg5-n", type=int, default=1024, help="cfg5: field / LIC-volume resolution")
    ap.add_argument("--cfg5-size", type=int, default=4096, help="cfg5: image size")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
