"""Summarise an ncu report (--set full) for profiles/: the metrics DESIGN.md and bench.py cite.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01x_ncu_summary.csv "note" [traffic_key] [launch_index]

Writes one CSV (metric, unit, value) for one kernel launch of the report (the first by default), the per-region shared-memory
wavefront breakdown and the top stall sites from the source page, and — when traffic_key is given —
records dram bytes per launch in profiles/traffic.json for bench.py's roofline.traffic."""
import csv
import json
import os
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep, dst, note = sys.argv[1], sys.argv[2], sys.argv[3]
    key = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "-" else None
    idx = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2 + idx]
    rows = [("# " + note, "", ""), ("kernel", "", vals[hdr.index("Kernel Name")])]
    got = {}
    for i, h in enumerate(hdr):
        if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            rows.append((h, units[i], vals[i]))
            got[h] = (units[i], vals[i])
    try:
        src = page(rep, "source") if idx == 0 else [[], []]
        if len(src) >= 3:
            ix0 = {h: i for i, h in enumerate(src[1])}
            _ = [int(r[ix0["# Samples"]]) for r in src[2:]]          # multi-kernel reports print ragged source pages: skip them
    except Exception:
        src = [[], []]
    if len(src) < 3:
        with open(dst, "w", newline="") as f:
            csv.writer(f).writerows(rows)
        print("wrote", dst, "(no source page for this launch index)")
        return
    sh, data = src[1], src[2:]
    ix = {h: i for i, h in enumerate(sh)}
    tot_s = sum(int(r[ix["# Samples"]]) for r in data) or 1
    tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
    rows.append(("source: warp instructions executed", "inst", str(tot_i)))
    by_op = {}
    for r in data:
        w = int(r[ix["L1 Wavefronts Shared"]])
        if w:
            s = r[ix["Source"]].split()
            op = s[1] if s[0].startswith("@") else s[0]
            by_op[op] = by_op.get(op, 0) + w
    for op, w in sorted(by_op.items(), key=lambda kv: -kv[1]):
        rows.append((f"source: shared wavefronts {op}", "wavefronts", str(w)))
    top = sorted(((int(r[ix["# Samples"]]), r[ix["Source"]].strip()) for r in data), reverse=True)[:10]
    for n, s in top:
        rows.append((f"source: stall samples {100.0 * n / tot_s:.2f}%", "", s))
    with open(dst, "w", newline="") as f:
        csv.writer(f).writerows(rows)
    if key:
        def gb(name):
            u, v = got[name]
            scale = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}[u]
            return float(v) * scale
        path = os.path.join(os.path.dirname(os.path.abspath(dst)), "traffic.json")
        table = json.load(open(path)) if os.path.exists(path) else {}
        table[key] = {"dram_read_gb": gb("dram__bytes_read.sum"), "dram_write_gb": gb("dram__bytes_write.sum"),
                      "source": os.path.basename(dst)}
        json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", dst)


if __name__ == "__main__":
    main()
