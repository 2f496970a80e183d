#!/usr/bin/env python
"""bench.py — output Mpix/s of the equirect -> rectilinear hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic frames: all F frames of the
workload x all its views (map on the fly + gather [+ blur score]).  Prints ONE JSON line.

  value     whole-job output Mpix/s with the frames already resident in HBM (kernel only)
  e2e       the same metric through the public host API (Extractor.extract on pinned host
            buffers): H2D + kernels + D2H inside the timed region
  roofline  algorithmic bytes per launch / CUDA-event time of the extraction kernel, against the
            measured HBM copy peak (MEASURED_PEAKS.json) — see DESIGN.md for the byte definition
  cpu_baseline  the reference's OpenCV CPU path (oracle/cv2_path.py: same cv2/numpy calls the
            reference makes) timed on this box's host cores on a bounded sample (all threads, and 1 thread)
  configs   (N = 1, default workload only) the other BASELINE configurations, each with its own value,
            roofline, cpu_baseline and e2e (Lanczos4: also the integer-issue roofline)
  sharded_cfg4 / view_sharded_cfg5   (N > 1) BASELINE configs 4 and 5 on N GPUs: the 256-frame clip
            frame-sharded (strong scaling) with the NCCL all-gather of the sum pairs and the blur
            accept/reject replay INSIDE the timed region; the single 11520x5760 frame sharded by view

N > 1: launched by torchrun, one rank per GPU; the headline `value` shards frames by index (every rank
runs the same per-GPU batch: weak scaling).  --impl reference times the CPU path only (rank 0).
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
                 desc="7680x3840 clip -> cube 6x1920^2 bilinear + blur score"),
    "cfg4_noscore": dict(W=7680, H=3840, layout="cube", n=6, pitch=0, active=None, d=1920, fov=90, mode="linear", scores=False, F=32,
                 desc="7680x3840 clip -> cube 6x1920^2 bilinear (diagnostic: cfg4 without the blur score)"),
    "cfg4_ring": dict(W=7680, H=3840, layout="ring", n=6, pitch=0, active=None, d=1920, fov=90, mode="linear", scores=True, F=32,
                 desc="7680x3840 clip -> ring 6x1920^2 bilinear + blur score (diagnostic: cfg4 without the pole faces)"),
    "cfg1_h4": dict(W=3840, H=1920, layout="cube", n=6, pitch=0, active=[0, 1, 2, 3], d=1024, fov=90, mode="linear", scores=False, F=32,
                 desc="diagnostic: cfg1's four horizontal faces only"),
    "cfg1_polar": dict(W=3840, H=1920, layout="cube", n=6, pitch=0, active=[4, 5], d=1024, fov=90, mode="linear", scores=False, F=32,
                 desc="diagnostic: cfg1's Up / Down faces only"),
    "cfg5": dict(W=11520, H=5760, layout="ring", n=12, pitch=0, active=[0, 3, 6, 9], d=2048, fov=90, mode="lanczos", scores=False, F=1,
                 desc="11520x5760 -> ring 12 (active 0,3,6,9) 2048^2 Lanczos4"),
}
CLIP_FRAMES = 256            # BASELINE config 4: "synthetic 256-frame clip"


def base_frame(seed, H, W, distribution="smooth"):
    """noise = uniform bytes; smooth = the same generator at (H/8, W/8) upsampled with
    cv2.resize(INTER_CUBIC) (realistic blur scores, 1e2..1e4)."""
    rng = np.random.default_rng(seed)
    if distribution == "noise":
        return rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    lo = rng.integers(0, 256, (H // 8, W // 8, 3), dtype=np.uint8)
    import cv2
    return np.ascontiguousarray(cv2.resize(lo, (W, H), interpolation=cv2.INTER_CUBIC))


def synth_frames(F, H, W, distribution="smooth"):
    """SURVEY.md 8(d) clip recipe: 4 seeded base frames, frame i = roll(base[i % 4], 37 i, axis=1)."""
    bases = [base_frame(seed, H, W, distribution) for seed in range(min(4, F))]
    return [np.roll(bases[i % len(bases)], 37 * i, axis=1) for i in range(F)]


def synth_frames_device(torch, dev, F, H, W, first=0, distribution="smooth"):
    """The same clip built on the device (frames first .. first + F - 1): 4 base frames uploaded, rolled there."""
    bases = [torch.from_numpy(base_frame(seed, H, W, distribution)).to(dev) for seed in range(4)]
    out = torch.empty((F, H, W, 3), dtype=torch.uint8, device=dev)
    for i in range(F):
        g = first + i
        out[i].copy_(torch.roll(bases[g % 4], 37 * g, dims=1))
    return out


def measured_traffic(key):
    """DRAM bytes per launch of the bench kernel from the committed ncu capture (profiles/traffic.json,
    written by scripts/ncu_summary.py); None when this workload / batch / kernel has no capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        t = json.load(open(p)).get(key)
        if t:
            return t["dram_read_gb"] + t["dram_write_gb"], "profiles/" + t["source"]
    return None, None


def touched_source_pixels(ex, wl):
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


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


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


def cpu_baseline_block(wl, V, with_one_thread=True):
    """The reference's cv2 path on this box's host cores: all threads (best of 3) and one thread (one pass of 1 frame)."""
    try:
        import cv2
        n = cpu_sample_frames(wl)
        cv2_path, cviews, maps, cframes, map_s = cpu_reference_setup(wl, n)
        d = wl["d"]
        cpu_reference_step(cv2_path, wl, cviews, maps, cframes)
        best = min(cpu_reference_step(cv2_path, wl, cviews, maps, cframes) for _ in range(3))
        cores = int(cv2.getNumThreads())
        one = None
        if with_one_thread:
            cv2.setNumThreads(1)
            one = V * d * d / cpu_reference_step(cv2_path, wl, cviews, maps, cframes[:1]) / 1e6
            cv2.setNumThreads(os.cpu_count() or 1)
        enc = None
        try:                                            # the save step beside it: cv2.imencode of every view (what cv2.imwrite computes)
            t0 = time.perf_counter()
            cnt = 0
            for nm in list(maps)[:2]:
                mx, my = maps[nm]
                v = cv2.remap(cframes[0], mx, my, cv2.INTER_LANCZOS4 if wl["mode"] == "lanczos" else cv2.INTER_LINEAR, borderMode=cv2.BORDER_WRAP)
                cv2.imencode(".jpg", v, [cv2.IMWRITE_JPEG_QUALITY, 95])
                cnt += 1
            enc = cnt * d * d / (time.perf_counter() - t0) / 1e6
        except Exception:
            pass
        return {"value": n * V * d * d / best / 1e6, "unit": "Mpix/s", "cores": cores, "kind": "port",
                "remap_plus_jpeg_one_thread_value": enc, "remap_plus_jpeg_sample": "2 views: cv2.remap + cv2.imencode(q95), as the reference's loop + one saver thread do",
                "sample": f"{n} frame(s) x {V} views, best of 3, maps pre-built once ({map_s:.2f} s, untimed as in processor.py:259-265)",
                "one_thread_value": one, "one_thread_sample": "1 frame, 1 pass, cv2.setNumThreads(1)",
                "host_cpus": os.cpu_count(), "cpu_model": cpu_model(), "cv2": cv2.__version__, "numpy": np.__version__, "map_build_s": map_s}
    except Exception as e:
        return {"value": None, "unit": "Mpix/s", "cores": 0, "kind": "port", "sample": f"unavailable: {type(e).__name__}: {e}"}


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
        cfg = workload_config(args, wl)
        cfg["reference_frames_per_step"] = n          # the CPU arm times a bounded sample of the workload's batch; the metric is a rate
        line = {
            "impl": "reference", "metric": "output Mpix/s", "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": cfg,
            "views_per_s": n * V * args.steps / total,
            "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count(), "cpu_model": cpu_model(), "cv2": cv2.__version__, "numpy": np.__version__,
                             "map_build_s": map_s},
            "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
    except Exception as e:  # cv2 missing on the box: say so rather than crash the driver
        line = {"impl": "reference", "unavailable": f"{type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)


def workload_config(args, wl, name=None):
    return {"workload": f"{name or args.workload}: {wl['desc']}", "frames_per_step_per_gpu": wl["F"], "src": [wl["W"], wl["H"]],
            "views": len(wl["active"]) if wl["active"] else (6 if wl["layout"] == "cube" else wl["n"]),
            "out": wl["d"], "interp": wl["mode"], "blur_scores": wl["scores"], "layout": wl["layout"],
            "frames": "smooth distribution of SURVEY 8(d): rng(seed) at (H/8, W/8) upsampled by cv2.resize INTER_CUBIC; frame i = roll(base[i % 4], 37 i)",
            "l2_policy": "inputs larger than L2 (per-step input+output far above 126 MB); no explicit flush"}


# Integer-issue roofline for Lanczos4: 64 taps x 3 channels = 192 MACs = 96 dp2a per output pixel; the IDP pipe
# issues one warp instruction per 2 cycles per SM sub-partition (4 per SM): 148 SMs x 4 x 32 lanes / 2 at the SM clock.
def int_roofline(mpix, sm_mhz):
    clk = (sm_mhz or 1965.0) * 1e6
    peak_dp2a = 148 * 4 * 32 / 2.0 * clk                 # dp2a lane-instructions per second
    peak_mpix = peak_dp2a / 96.0 / 1e6
    return {"bound": "int-issue", "achieved": mpix, "peak": peak_mpix, "unit": "Mpix/s", "frac": mpix / peak_mpix,
            "model": "96 dp2a per output pixel; IDP.2A at 1 warp instruction / 2 cycles / SM sub-partition, 148 SMs at %.0f MHz" % (clk / 1e6)}


def bench_workload(pkg, torch, dist, args, name, wl, dev, local_rank, rank, world, steps, warmup, with_e2e, with_cpu, sampler=None, with_jpeg=False):
    """Kernel-only value + roofline (+ e2e, + CPU baseline) of one workload on this rank's GPU; returns a dict."""
    F, H, W, d = wl["F"], wl["H"], wl["W"], wl["d"]
    views = pkg.GeometryProcessor.generate_views(wl["n"], pitch_offset=wl["pitch"], layout_mode=wl["layout"])
    names, R = pkg.rotation_stack(views, wl["active"])
    V = len(names)
    ex = pkg.Extractor(H, W, d, wl["fov"], R, wl["mode"], device=local_rank, max_batch=int(os.environ.get("X360_BENCH_CHUNK", "0")), path=args.path,
                       flags=args.flags)
    d_frames = synth_frames_device(torch, dev, F, H, W)
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

    for _ in range(max(warmup, 3)):
        step()
    barrier()
    launches0 = pkg.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(steps):
        step()
        ev[i + 1].record(stream)
    barrier()
    launches = pkg.launch_count() - launches0
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    modes = ex.tile_modes()
    clocks = None
    if sampler is not None:
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
    mpix = out_pix_per_step * steps / (total_ms * 1e-3) / 1e6

    # ---- roofline of the extraction kernel (this rank's launches) --------------------------
    peak, peak_src = measured_peaks()
    U = touched_source_pixels(ex, wl)
    b_alg_frame = 3 * V * d * d + 3 * U + (16 * V if wl["scores"] else 0)
    kernel_ms = statistics.mean(step_ms)            # the extraction launch(es) of one step (+ tiny memsets)
    achieved = F * b_alg_frame / (kernel_ms * 1e-3) / 1e9
    lz = wl["mode"] == "lanczos"
    kernel_name = ("extract_tiled<%s,lanczos4>" if lz else "extract_tiled<%s,linear>") % ("scores" if wl["scores"] else "noscores")
    traffic, traffic_src = measured_traffic(f"{name}:F{F}:{kernel_name}")
    n_tv = max(1, sum(modes.values()))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full; a step of several launches: their sum)",
                "traffic_source": traffic_src, "algorithmic_gb_per_launch": F * b_alg_frame / 1e9, "peak_source": peak_src, "algorithmic_bytes_per_frame": b_alg_frame,
                "touched_source_pixels_per_frame": U, "bytes_per_output_pixel": b_alg_frame / (V * d * d),
                "kernel": kernel_name, "launches_per_step": launches / max(1, steps),
                "tile_modes": modes, "direct_share": (modes["direct_aligned"] + modes["direct_bytewise"]) / n_tv,
                "kernel_ms_mean": kernel_ms, "kernel_ms_min": min(step_ms), "frac_of_nominal_8TBs": achieved / 8000.0}
    res = {"config": workload_config(args, wl, name), "value": mpix, "unit": "Mpix/s", "ms_per_step": total_ms / steps,
           "views_per_s": F * V * world * steps / (total_ms * 1e-3), "roofline": roofline, "gpu_launches": int(launches), "plan": ex.info()}
    if lz:
        res["roofline_int"] = int_roofline(mpix / world, (clocks or {}).get("sm_mhz"))
    if clocks is not None:
        res["clocks"] = clocks

    # ---- end to end through the public host API (pinned host buffers) ----------------------
    if with_e2e:
        pin_in = pkg.PinnedBuffer((F, H, W, 3))
        pin_out = pkg.PinnedBuffer(ex.views_shape(F))
        for i, fr in enumerate(synth_frames(F, H, W)):
            pin_in.array[i] = fr
        e2e_steps = max(2, min(steps, 5))
        ex.extract(pin_in.array, want_scores=wl["scores"], out=pin_out.array)      # warm-up (allocates the pipeline)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ex.extract(pin_in.array, want_scores=wl["scores"], out=pin_out.array)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        res["e2e"] = {"value": out_pix_per_step * e2e_steps / e2e_s / 1e6, "unit": "Mpix/s",
                      "h2d_bytes_per_step": F * ex.upload_rows()[1] * W * 3, "d2h_bytes_per_step": F * V * d * d * 3 + (16 * F * V if wl["scores"] else 0),
                      "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                      "uploaded_rows": list(ex.upload_rows()), "src_rows": H,
                      "api": "Extractor.extract(pinned numpy) -> x360_extract_host (chunked H2D of the source rows the views read / kernel / D2H on side streams)"}
        if wl["scores"]:
            # the blur filter on (src/core/processor.py:341-376, plain threshold at the batch's median score): only the views the
            # machine keeps cross PCIe on the way back — sums first, the host replays the machine one chunk behind the GPU
            cap = pkg._capi
            _, _, s_all = ex.extract_filtered(pin_in.array, _blur_state(cap, False, 0.0), out=pin_out.array.reshape(F * V, d, d, 3))
            thr = float(np.median(s_all))
            kept = 0
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                kv, keep, _ = ex.extract_filtered(pin_in.array, _blur_state(cap, False, thr), out=pin_out.array.reshape(F * V, d, d, 3))
                kept = int(keep.sum())
            torch.cuda.synchronize()
            f_s = time.perf_counter() - t0
            res["e2e_filtered"] = {"value": out_pix_per_step * e2e_steps / f_s / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": F * ex.upload_rows()[1] * W * 3,
                                   "d2h_bytes_per_step": kept * d * d * 3 + 16 * F * V, "views_kept": kept, "views_total": F * V,
                                   "blur_filter": {"smart": False, "threshold": thr, "note": "median score of the batch: about half the views are kept"},
                                   "steps": e2e_steps, "ms_per_step": 1e3 * f_s / e2e_steps,
                                   "api": "Extractor.extract_filtered(pinned numpy, blur state) -> x360_extract_host_filtered (value counts every extracted view, kept or not, as the reference's loop does)"}
        pin_out.free()
        if with_jpeg:
            # the same call with the reference's save step on the device: frames in, the .jpg files (quality 95, byte-identical
            # to cv2.imwrite's) out — only compressed bytes cross PCIe on the way back
            jcap = min(ex.jpeg_capacity(F), 2 * F * H * W * 3 + (64 << 20))
            pin_j = pkg.PinnedBuffer((jcap,))
            off = np.zeros(F * V + 1, np.int64)
            total = ex.extract_jpeg_into(pin_in.array, pin_j.array, off, 95)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                total = ex.extract_jpeg_into(pin_in.array, pin_j.array, off, 95)
            torch.cuda.synchronize()
            j_s = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([j_s], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                j_s = float(t.item())
            res["e2e_encoded"] = {"value": out_pix_per_step * e2e_steps / j_s / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": F * ex.upload_rows()[1] * W * 3,
                                  "d2h_bytes_per_step": int(total), "jpeg_quality": 95, "bytes_per_output_pixel": total / (F * V * d * d),
                                  "steps": e2e_steps, "ms_per_step": 1e3 * j_s / e2e_steps,
                                  "api": "Extractor.extract_jpeg_into(pinned numpy) -> x360_extract_host_jpeg (H2D / extract / JPEG encode / D2H of the files)",
                                  "parity": "files byte-identical to cv2.imencode('.jpg', view, [IMWRITE_JPEG_QUALITY, 95]) (tests/test_gpu_parity.py)"}
            pin_j.free()
        pin_in.free()
    else:
        res["e2e"] = None
    res["cpu_baseline"] = cpu_baseline_block(wl, V) if with_cpu else None
    ex.close()
    del d_frames, d_out
    torch.cuda.empty_cache()
    return res


def _blur_state(cap, smart, threshold):
    st = cap.BlurState()
    cap.lib().x360_blur_state_init(st, 1, int(smart), float(threshold))
    return st


def bench_sharded_cfg4(pkg, torch, dist, args, dev, local_rank, rank, world):
    """BASELINE config 4 on N GPUs: the 256-frame clip, frame-sharded (strong scaling).  One timed step = this rank's
    contiguous block of frames through the extraction kernel(s) with the blur score, the NCCL all-gather of the int64 sum
    pairs, and the blur accept/reject replay over the full [256][6] table on every rank (processor.py:345-367)."""
    wl = dict(WORKLOADS["cfg4"])
    H, W, d = wl["H"], wl["W"], wl["d"]
    lo, hi = pkg.shard_bounds(CLIP_FRAMES, world, rank)
    views = pkg.GeometryProcessor.generate_views(wl["n"], pitch_offset=wl["pitch"], layout_mode=wl["layout"])
    names, R = pkg.rotation_stack(views, wl["active"])
    V = len(names)
    ex = pkg.Extractor(H, W, d, wl["fov"], R, wl["mode"], device=local_rank, flags=args.flags)
    d_frames = synth_frames_device(torch, dev, hi - lo, H, W, first=lo)
    d_out = torch.empty(ex.views_shape(hi - lo), dtype=torch.uint8, device=dev)
    d_s = torch.zeros((hi - lo, V), dtype=torch.int64, device=dev)
    d_s2 = torch.zeros_like(d_s)
    stream = torch.cuda.current_stream()
    keep = None

    def step():
        nonlocal keep
        ex.extract_device(d_frames, d_out, d_s, d_s2, stream=stream)
        gs, gs2 = pkg.gather_sums(d_s, d_s2, CLIP_FRAMES)                                  # NCCL all-gather (+ D2H of 2 x 12 KB)
        scores = pkg.scores_from_sums(gs, gs2, d * d)
        keep = pkg.blur_replay(scores, smart=True, threshold=100.0)                        # sequential machine (C), every rank

    for _ in range(2):
        step()
    dist.barrier(); torch.cuda.synchronize()
    n_steps = 3
    t0 = time.perf_counter()
    for _ in range(n_steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    kept = int(np.sum(keep))
    # (outside the timed region) the C replay made the decisions of the Python restatement of processor.py:345-367
    gs, gs2 = pkg.gather_sums(d_s, d_s2, CLIP_FRAMES)
    ref_keep = np.array(pkg.BlurFilter(True, True, 100.0).replay(pkg.scores_from_sums(gs, gs2, d * d)), dtype=bool)
    if not np.array_equal(ref_keep, np.asarray(keep, dtype=bool).reshape(ref_keep.shape)):
        raise RuntimeError("blur_replay (C) and BlurFilter.replay (Python) disagree on the sharded clip")
    ex.close()
    del d_frames, d_out
    torch.cuda.empty_cache()
    return {"workload": "cfg4: 7680x3840 256-frame clip -> cube 6x1920^2 bilinear + blur score, frame-sharded", "scaling": "strong",
            "frames_total": CLIP_FRAMES, "frames_per_gpu": hi - lo, "n_gpus": world, "steps": n_steps, "ms_per_step": 1e3 * dt / n_steps,
            "value": CLIP_FRAMES * V * d * d * n_steps / dt / 1e6, "unit": "Mpix/s", "views_per_s": CLIP_FRAMES * V * n_steps / dt,
            "collective": "nccl all_gather of int64 [2][F/N][6] sum pairs + blur accept/reject replay over [256][6] on every rank, inside the timed step",
            "timing": "wall clock between barrier + synchronize, max over ranks (the step ends in host code: the replay)",
            "kept_views": kept, "of": CLIP_FRAMES * V}


def bench_view_sharded_cfg5(pkg, torch, dist, args, dev, local_rank, rank, world):
    """BASELINE config 5 on N GPUs: ONE 11520x5760 frame, ring 12 with 4 active views, Lanczos4 — sharded by view
    (every rank holds the same frame and its own block of the view list; ranks beyond the 4th view have no work)."""
    wl = dict(WORKLOADS["cfg5"])
    H, W, d = wl["H"], wl["W"], wl["d"]
    views = pkg.GeometryProcessor.generate_views(wl["n"], pitch_offset=wl["pitch"], layout_mode=wl["layout"])
    names, R = pkg.rotation_stack(views, wl["active"])
    V = len(names)
    lo, hi = pkg.shard_bounds(V, world, rank)
    d_frames = synth_frames_device(torch, dev, 1, H, W)
    stream = torch.cuda.current_stream()
    ex = d_out = None
    if hi > lo:
        ex = pkg.Extractor(H, W, d, wl["fov"], R[lo:hi], wl["mode"], device=local_rank, flags=args.flags)
        d_out = torch.empty(ex.views_shape(1), dtype=torch.uint8, device=dev)
        for _ in range(3):
            ex.extract_device(d_frames, d_out, stream=stream)
    dist.barrier(); torch.cuda.synchronize()
    n_steps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if ex is not None:
        for _ in range(n_steps):
            ex.extract_device(d_frames, d_out, stream=stream)
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if ex is not None:
        ex.close()
    return {"workload": "cfg5: 11520x5760 single frame -> ring 12 (active 0,3,6,9) 2048^2 Lanczos4, view-sharded", "scaling": "strong",
            "n_gpus": world, "views_total": V, "views_per_gpu": [pkg.shard_bounds(V, world, r)[1] - pkg.shard_bounds(V, world, r)[0] for r in range(world)],
            "steps": n_steps, "ms_per_step": ms / n_steps, "value": V * d * d * n_steps / (ms * 1e-3) / 1e6, "unit": "Mpix/s",
            "collective": "none (no blur score in this config)", "timing": "CUDA events on the launch stream, max over ranks"}


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
    ap.add_argument("--no-extra", action="store_true", help="skip the other-config blocks (N = 1) / the sharded blocks (N > 1)")
    ap.add_argument("--kernel-only", action="store_true", help="skip e2e, cpu baseline and the extra blocks (for ncu runs)")
    ap.add_argument("--path", type=int, default=0, help="x360 kernel path: 0 auto (tiled); 1 / 2 exist only in -DX360_LAB builds")
    ap.add_argument("--flags", type=int, default=0, help="x360_params.flags (X360_FLAG_*): plan overrides")
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

    sampler = ClockSampler(local_rank)
    sampler.start()                       # nvidia-smi needs ~100 ms to start: begin before the warm-up
    full = not args.kernel_only
    main_res = bench_workload(pkg, torch, dist, args, args.workload, wl, dev, local_rank, rank, world, args.steps, args.warmup,
                              with_e2e=full and not args.no_e2e, with_cpu=full and not args.no_cpu_baseline and rank == 0 and world == 1, sampler=sampler,
                              with_jpeg=full and not args.no_e2e and wl["mode"] != "none")

    extra = {}
    if full and not args.no_extra and args.workload == "cfg2" and args.frames == 0:
        if world == 1:
            # the other BASELINE configurations, each to the same bar (value, roofline, e2e, CPU path beside it)
            for name in ("cfg1", "cfg4", "cfg3", "cfg5"):
                try:
                    extra[name] = bench_workload(pkg, torch, dist, args, name, dict(WORKLOADS[name]), dev, local_rank, rank, world, 10, 3,
                                                 with_e2e=not args.no_e2e, with_cpu=not args.no_cpu_baseline)
                except Exception as e:
                    extra[name] = {"error": f"{type(e).__name__}: {e}"}
        else:
            for key, fn in (("sharded_cfg4", bench_sharded_cfg4), ("view_sharded_cfg5", bench_view_sharded_cfg5)):
                try:
                    extra[key] = fn(pkg, torch, dist, args, dev, local_rank, rank, world)
                except Exception as e:
                    extra[key] = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        line = {
            "metric": "output Mpix/s", "value": main_res["value"], "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": main_res["config"], "views_per_s": main_res["views_per_s"],
            "roofline": main_res["roofline"], "cpu_baseline": main_res["cpu_baseline"], "e2e": main_res["e2e"],
            "e2e_encoded": main_res.get("e2e_encoded"),
            "gpu_launches": main_res["gpu_launches"], "clocks": main_res.get("clocks"), "plan": main_res["plan"], "lib": pkg.version(),
        }
        if "roofline_int" in main_res:
            line["roofline_int"] = main_res["roofline_int"]
        if "e2e_filtered" in main_res:
            line["e2e_filtered"] = main_res["e2e_filtered"]
        if world == 1 and extra:
            line["configs"] = extra
        else:
            line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
