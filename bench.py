#!/usr/bin/env python
"""bench.py — output Mpix/s of the equirect -> rectilinear hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic frames: all F frames of the
workload x all its views (map on the fly + gather [+ fused blur score]).  Prints ONE JSON line.

  value     whole-job output Mpix/s with the frames already resident in HBM (kernel only)
  e2e       the same metric through the public host API (Extractor.extract on pinned host
            buffers): H2D + kernels + D2H inside the timed region
  roofline  algorithmic bytes per launch / CUDA-event time of the extraction kernel, against the
            measured HBM copy peak (MEASURED_PEAKS.json) — see DESIGN.md for the byte definition
  cpu_baseline  the reference's OpenCV CPU path (oracle/cv2_path.py: same cv2/numpy calls the
            reference makes) timed on this box's host cores on a bounded sample

N > 1: launched by torchrun, one rank per GPU; frames shard by index (every rank runs the same
per-GPU batch: weak scaling); the only collective is the all-gather of the blur-score sums when
the workload has scores.  --impl reference times the CPU path only (rank 0).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs as restated in SURVEY.md 8(d)
WORKLOADS = {
    "cfg1": dict(W=3840, H=1920, layout="cube", n=6, pitch=0, active=None, d=1024, fov=90, mode="linear", scores=False, F=32,
                 desc="3840x1920 -> cube 6x1024^2 bilinear"),
    "cfg2": dict(W=5760, H=2880, layout="ring", n=8, pitch=0, active=None, d=1600, fov=90, mode="linear", scores=False, F=32,
                 desc="5760x2880 frame batch -> ring 8 views 90deg FOV (50% overlap) 1600^2 bilinear"),
    "cfg3": dict(W=7680, H=3840, layout="fibonacci", n=20, pitch=-20, active=None, d=1920, fov=90, mode="lanczos", scores=False, F=1,
                 desc="7680x3840 -> fibonacci 20 views pitch -20 1920^2 Lanczos4"),
    "cfg4": dict(W=7680, H=3840, layout="cube", n=6, pitch=0, active=None, d=1920, fov=90, mode="linear", scores=True, F=32,
                 desc="7680x3840 clip -> cube 6x1920^2 bilinear + fused blur score"),
    "cfg4_noscore": dict(W=7680, H=3840, layout="cube", n=6, pitch=0, active=None, d=1920, fov=90, mode="linear", scores=False, F=32,
                 desc="7680x3840 clip -> cube 6x1920^2 bilinear (diagnostic: cfg4 without the blur score)"),
    "cfg4_ring": dict(W=7680, H=3840, layout="ring", n=6, pitch=0, active=None, d=1920, fov=90, mode="linear", scores=True, F=32,
                 desc="7680x3840 clip -> ring 6x1920^2 bilinear + blur score (diagnostic: cfg4 without the pole faces)"),
    "cfg5": dict(W=11520, H=5760, layout="ring", n=12, pitch=0, active=[0, 3, 6, 9], d=2048, fov=90, mode="lanczos", scores=False, F=1,
                 desc="11520x5760 -> ring 12 (active 0,3,6,9) 2048^2 Lanczos4"),
}


def synth_frames(F, H, W, distribution="smooth"):
    """SURVEY.md 8(d) clip recipe: 4 seeded base frames, frame i = roll(base[i % 4], 37 i, axis=1)."""
    bases = []
    for seed in range(min(4, F)):
        rng = np.random.default_rng(seed)
        if distribution == "noise":
            bases.append(rng.integers(0, 256, (H, W, 3), dtype=np.uint8))
        else:  # smooth: low-resolution noise, nearest x8 blocks softened by a 2-tap average (cheap, numpy only)
            lo = rng.integers(0, 256, (H // 8 + 2, W // 8 + 2, 3), dtype=np.uint8)
            up = np.repeat(np.repeat(lo, 8, axis=0), 8, axis=1).astype(np.uint16)
            sm = (up[:-4, :-4] + up[4:, 4:]) // 2
            bases.append(np.ascontiguousarray(sm[:H, :W].astype(np.uint8)))
    return [np.roll(bases[i % len(bases)], 37 * i, axis=1) for i in range(F)]


def measured_traffic(key):
    """DRAM bytes per launch of the bench kernel from the committed ncu capture (profiles/traffic.json,
    written by scripts/ncu_summary.py); None when this workload / batch / kernel has no capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        t = json.load(open(p)).get(key)
        if t:
            return t["dram_read_gb"] + t["dram_write_gb"], "profiles/" + t["source"]
    return None, None


def touched_source_pixels(pkg, ex, wl):
    """U of SURVEY.md 8(d): |union over views of source pixels addressed by any tap| for one frame,
    from the plan's own (GPU-made) float32 maps, quantised like cv2.remap does."""
    H, W = wl["H"], wl["W"]
    hit = np.zeros((H, W), dtype=bool)
    k = 8 if wl["mode"] == "lanczos" else 2
    for v in range(ex.n_views):
        mx, my = ex.make_maps(v)
        sx = np.rint(mx.astype(np.float32) * np.float32(32)).astype(np.int64)
        sy = np.rint(my.astype(np.float32) * np.float32(32)).astype(np.int64)
        ix = ((sx >> 5) - (3 if k == 8 else 0)) % W
        iy = ((sy >> 5) - (3 if k == 8 else 0)) % H
        hit[iy.ravel(), ix.ravel()] = True
    # dilate the tap origins by the k x k footprint (wrap-around on both axes)
    acc = hit.copy()
    for s in range(1, k):
        acc |= np.roll(hit, s, axis=1)
    hit = acc
    acc = hit.copy()
    for s in range(1, k):
        acc |= np.roll(hit, s, axis=0)
    return int(acc.sum())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---- CPU reference path (the oracle's cv2 call sequence) --------------------------------------
def cpu_reference_setup(wl, n_frames):
    from oracle import cv2_path
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)
    views = cv2_path.layout_views(wl["n"], wl["pitch"], wl["layout"])
    t0 = time.perf_counter()
    maps = cv2_path.build_maps(wl["H"], wl["W"], wl["d"], wl["fov"], views, wl["active"])
    map_s = time.perf_counter() - t0
    frames = synth_frames(n_frames, wl["H"], wl["W"])
    return cv2_path, views, maps, frames, map_s


def cpu_reference_step(cv2_path, wl, views, maps, frames):
    t0 = time.perf_counter()
    cv2_path.extract(frames, views, maps, wl["mode"], want_scores=wl["scores"], keep_views=False)
    return time.perf_counter() - t0


def cpu_sample_frames(wl):
    # aim for a few seconds per pass: Lanczos4 is ~12x slower than bilinear on the CPU
    return 1 if wl["mode"] == "lanczos" else 4


def run_reference_arm(args, wl, rank):
    if rank != 0:
        return
    try:
        import cv2
        n = cpu_sample_frames(wl)
        cv2_path, views, maps, frames, map_s = cpu_reference_setup(wl, n)
        V = len(maps)
        for _ in range(args.warmup):
            cpu_reference_step(cv2_path, wl, views, maps, frames)
        times = [cpu_reference_step(cv2_path, wl, views, maps, frames) for _ in range(args.steps)]
        total = sum(times)
        mpix = n * V * wl["d"] ** 2 * args.steps / total / 1e6
        cores = int(cv2.getNumThreads())
        sample = f"{n} frame(s) x {V} views per step (of the workload's {wl['F']}), maps pre-built once ({map_s:.2f} s, untimed as in processor.py:259-265)"
        line = {
            "impl": "reference", "metric": "output Mpix/s", "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, wl),
            "views_per_s": n * V * args.steps / total,
            "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count(), "cv2": cv2.__version__, "numpy": np.__version__,
                             "map_build_s": map_s},
            "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
    except Exception as e:  # cv2 missing on the box: say so rather than crash the driver
        line = {"impl": "reference", "unavailable": f"{type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)


def workload_config(args, wl):
    return {"workload": f"{args.workload}: {wl['desc']}", "frames_per_step_per_gpu": wl["F"], "src": [wl["W"], wl["H"]],
            "views": len(wl["active"]) if wl["active"] else (6 if wl["layout"] == "cube" else wl["n"]),
            "out": wl["d"], "interp": wl["mode"], "blur_scores": wl["scores"], "layout": wl["layout"],
            "l2_policy": "inputs larger than L2 (per-step input+output far above 126 MB); no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="override frames per step per GPU")
    ap.add_argument("--impl", default="x360", choices=["x360", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="skip e2e and cpu baseline (for ncu runs)")
    ap.add_argument("--path", type=int, default=0, help="x360 kernel path: 0 auto, 1 direct, 2 staged, 3 tiled")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.frames > 0:
        wl["F"] = args.frames
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the x360 arm has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("360extractor_b200")

    F, H, W, d = wl["F"], wl["H"], wl["W"], wl["d"]
    views = pkg.GeometryProcessor.generate_views(wl["n"], pitch_offset=wl["pitch"], layout_mode=wl["layout"])
    names, R = pkg.rotation_stack(views, wl["active"])
    V = len(names)
    ex = pkg.Extractor(H, W, d, wl["fov"], R, wl["mode"], device=local_rank, max_batch=int(os.environ.get("X360_BENCH_CHUNK", "0")), path=args.path)

    # ---- device-resident batch -------------------------------------------------------------
    host_frames = synth_frames(F, H, W)
    d_frames = torch.empty((F, H, W, 3), dtype=torch.uint8, device=dev)
    for i, fr in enumerate(host_frames):
        d_frames[i].copy_(torch.from_numpy(fr))
    d_out = torch.empty(ex.views_shape(F), dtype=torch.uint8, device=dev)
    d_s = torch.zeros((F, V), dtype=torch.int64, device=dev) if wl["scores"] else None
    d_s2 = torch.zeros((F, V), dtype=torch.int64, device=dev) if wl["scores"] else None
    stream = torch.cuda.current_stream()

    def step():
        ex.extract_device(d_frames, d_out, d_s, d_s2, stream=stream)
        if wl["scores"] and world > 1:
            pkg.gather_sums(d_s, d_s2, F * world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()                       # nvidia-smi needs ~100 ms to start: begin before the warm-up
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = pkg.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    barrier()
    launches = pkg.launch_count() - launches0
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    modes = ex.tile_modes()
    # keep the SAME steps running (untimed) until the sampler has seen >= 1 s of load: the timed
    # region alone is a few tens of ms, shorter than one nvidia-smi sampling period
    soak_t0 = time.perf_counter()
    while time.perf_counter() - soak_t0 < 1.2:
        for _ in range(10):
            ex.extract_device(d_frames, d_out, d_s, d_s2, stream=stream)
        torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["window"] = "warm-up + timed steps + 1.2 s soak of identical steps (100 ms sampling)"
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    out_pix_per_step = F * V * d * d * world
    mpix = out_pix_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # ---- roofline of the extraction kernel (rank 0's launches) -----------------------------
    peak, peak_src = measured_peaks()
    U = touched_source_pixels(pkg, ex, wl)
    b_alg_frame = 3 * V * d * d + 3 * U + (16 * V if wl["scores"] else 0)
    kernel_ms = statistics.mean(step_ms)            # one extraction kernel per step (+2 tiny memsets with scores)
    achieved = F * b_alg_frame / (kernel_ms * 1e-3) / 1e9
    lz = wl["mode"] == "lanczos"
    kernel_name = {3: "extract_tiled<%s,lanczos4>" if lz else "extract_tiled<%s,linear>", 2: "extract_linear_staged<%s>"}.get(
        ex.info()["path"], "extract_direct<LanczosSampler,%s>" if lz else "extract_direct<LinearSampler,%s>") \
        % ("scores" if wl["scores"] else "noscores")
    traffic, traffic_src = measured_traffic(f"{args.workload}:F{F}:{kernel_name}")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)",
                "traffic_source": traffic_src, "algorithmic_gb_per_launch": F * b_alg_frame / 1e9, "peak_source": peak_src, "algorithmic_bytes_per_frame": b_alg_frame,
                "touched_source_pixels_per_frame": U, "bytes_per_output_pixel": b_alg_frame / (V * d * d),
                "kernel": kernel_name,
                "tile_modes": modes,
                "kernel_ms_mean": kernel_ms, "kernel_ms_min": min(step_ms), "frac_of_nominal_8TBs": achieved / 8000.0}

    # ---- end to end through the public host API (pinned host buffers) ----------------------
    e2e = None
    if not (args.no_e2e or args.kernel_only):
        pin_in = pkg.PinnedBuffer((F, H, W, 3))
        pin_out = pkg.PinnedBuffer(ex.views_shape(F))
        for i, fr in enumerate(host_frames):
            pin_in.array[i] = fr
        e2e_steps = max(2, min(args.steps, 5))
        ex.extract(pin_in.array, want_scores=wl["scores"], out=pin_out.array)      # warm-up (allocates the pipeline)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            _, sc = ex.extract(pin_in.array, want_scores=wl["scores"], out=pin_out.array)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": out_pix_per_step * e2e_steps / e2e_s / 1e6, "unit": "Mpix/s",
               "h2d_bytes_per_step": F * H * W * 3, "d2h_bytes_per_step": F * V * d * d * 3 + (16 * F * V if wl["scores"] else 0),
               "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
               "api": "Extractor.extract(pinned numpy) -> x360_extract_host (chunked H2D / kernel / D2H on side streams)"}
        pin_in.free(); pin_out.free()

    # ---- CPU baseline beside it (rank 0, N=1 only) -----------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not (args.no_cpu_baseline or args.kernel_only):
        try:
            import cv2
            n = cpu_sample_frames(wl)
            cv2_path, cviews, maps, cframes, map_s = cpu_reference_setup(wl, n)
            cpu_reference_step(cv2_path, wl, cviews, maps, cframes)
            best = min(cpu_reference_step(cv2_path, wl, cviews, maps, cframes) for _ in range(3))
            cpu = {"value": n * V * d * d / best / 1e6, "unit": "Mpix/s", "cores": int(cv2.getNumThreads()), "kind": "port",
                   "sample": f"{n} frame(s) x {V} views, best of 3, maps pre-built once ({map_s:.2f} s, untimed)",
                   "host_cpus": os.cpu_count(), "cv2": cv2.__version__, "map_build_s": map_s}
        except Exception as e:
            cpu = {"value": None, "unit": "Mpix/s", "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}

    if rank == 0:
        line = {
            "metric": "output Mpix/s", "value": mpix, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, wl), "views_per_s": F * V * world * args.steps / (total_ms * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "plan": ex.info(), "lib": pkg.version(),
        }
        print(json.dumps(line), flush=True)
    ex.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
