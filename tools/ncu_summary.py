#!/usr/bin/env python
"""Turn `ncu --set full` reports into the markdown tables kept under profiles/.

usage: tools/ncu_summary.py report.ncu-rep|raw.csv [more ...] [--peak-gbs 6547.8] [--alg name=bytes ...]

One row per captured launch: device time, DRAM bytes read + written (dram__bytes_read.sum + dram__bytes_write.sum),
the DRAM throughput they imply and its fraction of the measured HBM peak (MEASURED_PEAKS.json), registers, achieved
occupancy, issue-slot and FP64-pipe utilisation, executed warp instructions, the largest stall reasons and -- for the
cross-GPU kernels -- NVLink bytes.  `--alg substring=bytes` adds the algorithmic bytes of kernels whose name contains
the substring, so the table shows achieved (algorithmic) GB/s next to the DRAM traffic.
Runs here (no GPU needed): ncu only parses the report.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STALLS = ["barrier", "branch_resolving", "dispatch_stall", "drain", "lg_throttle", "long_scoreboard", "math_pipe_throttle",
          "membar", "mio_throttle", "no_instruction", "short_scoreboard", "sleeping", "tex_throttle", "wait"]


def rows_of(report):
    if report.endswith(".csv"):        # `ncu -i report.ncu-rep --page raw --csv` saved on the GPU box
        out = open(report).read()
    else:
        out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    rows = []
    for r in rd[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            if u in scale and h not in ("ID",):
                try:
                    v = repr(float(v.replace(",", "")) * scale[u])     # bytes / microseconds
                except ValueError:
                    pass
            d[h] = v
        rows.append(d)
    return rows


def f(row, key, default=0.0):
    v = row.get(key, "")
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return default


def short_name(n):
    n = n.replace("smc::<unnamed>::", "").replace("smc::", "").replace("(int)", "").replace("(bool)", "").replace("unnamed>::", "")
    if n.startswith("void "):
        n = n[5:]
    cut = n.find("(")
    return n[:cut] if cut > 0 else n


def json_summary(pairs, out_path):
    """pairs: [(tag, report)]: per-kernel numbers of the LAST launch of every kernel in each report (the stage forced to
    resample) -> {"kernels": {"<kernel>_<tag>": {...}}}; bench.py reads k_mutate_c2 for roofline.traffic."""
    out = {"source": "ncu --set full --clock-control none, tools/profile_kernels.py, one launch per kernel", "kernels": {}}
    for tag, rep in pairs:
        for r in rows_of(rep):
            name = short_name(r["Kernel Name"])
            key = name.split("<")[0].replace("unnamed>::", "") + "_" + tag
            out["kernels"][key] = {
                "kernel": name, "time_us": f(r, "gpu__time_duration.sum"), "dram_bytes_read": f(r, "dram__bytes_read.sum"),
                "dram_bytes_write": f(r, "dram__bytes_write.sum"), "registers": int(f(r, "launch__registers_per_thread")),
                "warps_active_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "issue_active_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "fp64_pipe_pct": f(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                "warp_instructions": f(r, "smsp__inst_executed.sum"), "grid": int(f(r, "launch__grid_size")),
                "block": int(f(r, "launch__block_size"))}
    json.dump(out, open(out_path, "w"), indent=1)


def main():
    args = sys.argv[1:]
    peak = None
    alg = {}
    reports = []
    i = 0
    if args and args[0] == "--json":          # --json out.json tag=report [tag=report ...]
        json_summary([tuple(a.split("=", 1)) for a in args[2:]], args[1])
        return
    while i < len(args):
        if args[i] == "--peak-gbs":
            peak = float(args[i + 1]); i += 2
        elif args[i] == "--alg":
            k, v = args[i + 1].split("="); alg[k] = float(v); i += 2
        else:
            reports.append(args[i]); i += 1
    if peak is None:
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            peak = 6650.0
    print("| kernel | time µs | DRAM rd+wr MB | DRAM GB/s (%% of %.0f) | algorithmic MB -> GB/s (%%) | regs | warps active %% | issue active %% | "
          "FP64 pipe %% | warp inst | top stalls (warps per issue) | NVLink rx+tx MB |" % peak)
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for rep in reports:
        for r in rows_of(rep):
            name = short_name(r["Kernel Name"])
            t_us = f(r, "gpu__time_duration.sum")
            dram = f(r, "dram__bytes_read.sum") + f(r, "dram__bytes_write.sum")
            gbs = dram / (t_us * 1e-6) / 1e9 if t_us else 0.0
            a = next((v for k, v in alg.items() if k in name), None)
            a_txt = "-"
            if a:
                ag = a / (t_us * 1e-6) / 1e9
                a_txt = "%.1f -> %.0f (%.1f %%)" % (a / 1e6, ag, 100 * ag / peak)
            st = sorted(((f(r, "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s), s) for s in STALLS), reverse=True)[:3]
            nvl = f(r, "nvlrx__bytes.sum") + f(r, "nvltx__bytes.sum")
            print("| `%s` | %.1f | %.1f | %.0f (%.1f %%) | %s | %d | %.1f | %.1f | %.1f | %.3g | %s | %s |" % (
                name, t_us, dram / 1e6, gbs, 100 * gbs / peak, a_txt, int(f(r, "launch__registers_per_thread")),
                f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                f(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), f(r, "smsp__inst_executed.sum"),
                ", ".join("%s %.2f" % (s, v) for v, s in st), ("%.2f" % (nvl / 1e6)) if nvl else "-"))


if __name__ == "__main__":
    main()
