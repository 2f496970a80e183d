"""PCIe ceiling of the box for the pipelined host path: pinned H2D, D2H and both at once, on 1 .. N GPUs concurrently.

    python scripts/pcie_probe.py [--gpus 1,2,4,8] [--out profiles/r02_pcie_probe.json]

One process per GPU (CUDA_VISIBLE_DEVICES), all started on a common wall-clock tick, each moving 1 GiB per copy for ~2 s per
phase; the aggregate is the sum of the per-process rates over the common window.  This is the platform number the end-to-end
rate of bench.py (`e2e`) is to be read against: x360_extract_host moves h2d_bytes + d2h_bytes per step through the same path."""
import argparse
import json
import os
import subprocess
import sys
import time


def child(start_at):
    import torch
    dev = torch.device("cuda", 0)
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d(); d2h()

    for fn in (h2d, d2h):
        fn()
    torch.cuda.synchronize()
    out = {}
    for k, (name, fn, mult) in enumerate((("h2d", h2d, 1), ("d2h", d2h, 1), ("both", both, 2))):
        t_phase = start_at + 4.0 * k
        while time.time() < t_phase:
            time.sleep(0.001)
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 2.0:
            fn()
            torch.cuda.synchronize()
            reps += 1
        dt = time.perf_counter() - t0
        out[name] = mult * reps * n / dt / 1e9
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--out", default="")
    ap.add_argument("--child", type=float, default=0.0)
    a = ap.parse_args()
    if a.child:
        return child(a.child)
    results = {}
    for n in [int(x) for x in a.gpus.split(",")]:
        start_at = time.time() + 25.0 + 3.0 * n                    # import torch + pinning 2 GiB per process takes a while
        procs = []
        for g in range(n):
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(g))
            procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--child", repr(start_at)], env=env,
                                          stdout=subprocess.PIPE, text=True))
        per = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
        results[str(n)] = {"per_gpu": per, "aggregate_gbs": {k: sum(r[k] for r in per) for k in ("h2d", "d2h", "both")}}
        print(n, "GPU(s):", {k: round(v, 1) for k, v in results[str(n)]["aggregate_gbs"].items()}, "GB/s aggregate (both = sum of the two directions)", flush=True)
    if a.out:
        json.dump({"what": "pinned cudaMemcpyAsync ceiling, 1 GiB copies, one process per GPU, concurrent", "results": results,
                   "host_cpus": os.cpu_count()}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
