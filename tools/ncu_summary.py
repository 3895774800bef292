#!/usr/bin/env python
"""Condense ncu exports brought back in gpurun_out/ into small text/JSON files under profiles/.

    python tools/ncu_summary.py <tag> [--kernel chan256] [--samples-per-launch N]

Reads gpurun_out/{launches.csv,raw.csv,src.csv} (written by tools/gpu_round.sh) and writes
profiles/<tag>_launches.txt   per-kernel launch counts, total and share of device time
profiles/<tag>_<kernel>_ncu.txt   the --set full metrics that matter + SASS opcode histogram
profiles/chan_fm_traffic.json     dram bytes per input sample (bench.py's roofline.traffic source)
"""
from __future__ import annotations

import argparse
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__inst_executed_pipe_(alu|fma|lsu|xu|uniform|fp64|tma)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_(fma|alu|fp64)_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_ld|_op_st)?\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|"
    r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_static|shared_mem_per_block_dynamic|occupancy_limit_\w+|waves_per_multiprocessor)|"
    r"sm__cycles_elapsed\.max|lts__t_bytes\.sum|lts__t_sector_hit_rate\.pct|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio)$")


def short_kernel(name: str) -> str:
    m = re.search(r"(wc::)?(\w+)(<[^>]*>)?\(", name)
    if "at::" in name or "elementwise" in name:
        return "torch:" + (re.search(r"(\w+_kernel\w*)", name).group(1) if re.search(r"(\w+_kernel\w*)", name) else "kernel")
    return (m.group(2) + (m.group(3) or "")) if m else name[:60]


def launches(tag: str, csv_name: str = "launches.csv", title: str | None = None) -> None:
    path = os.path.join(OUT, csv_name)
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        k = short_kernel(r[4])
        a = agg.setdefault(k, [0, 0.0, r[8], r[7]])
        a[0] += 1
        a[1] += float(r[-1])
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, f"{tag}_launches.txt"), "w") as f:
        if title:
            f.write(f"# {title}\n")
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# {len(rows)} launches, total {tot / 1e6:.3f} ms\n")
        f.write(f"{'kernel':44s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}  grid block\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:44s} {a[0]:8d} {a[1] / 1e3:12.1f} {a[1] / 1e3 / a[0]:10.1f} {100 * a[1] / tot:6.1f}%  {a[2]} {a[3]}\n")


def full(tag: str, kernel: str, samples: int | None) -> None:
    raw = os.path.join(OUT, "raw.csv")
    if not os.path.exists(raw):
        return
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    body = [r for r in rows[2:] if kernel in dict(zip(hdr, r)).get("Kernel Name", "")]
    if not body:
        return
    d = dict(zip(hdr, body[-1]))
    lines = [f"# ncu --set full --clock-control none, kernel {d['Kernel Name']} (last of {len(body)} captured launches)"]
    for k in hdr:
        if KEEP.match(k):
            lines.append(f"{k:90s} {d[k]:>16s} {units[hdr.index(k)]}")
    def num(k):
        v, u = float(d[k]), units[hdr.index(k)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    lines.append(f"dram bytes per launch (read+write) = {dram:.0f}")
    if samples:
        lines.append(f"input samples per launch = {samples}; dram bytes per input sample = {dram / samples:.3f} "
                     f"(algorithmic 16.0)")
        with open(os.path.join(PROF, "chan_fm_traffic.json"), "w") as f:
            json.dump({"kernel": d["Kernel Name"], "dram_bytes_per_launch": dram, "samples_per_launch": samples,
                       "dram_bytes_per_sample": dram / samples, "source": f"profiles/{tag}_{kernel}_ncu.txt"}, f)
    src = os.path.join(OUT, "src.csv")
    if os.path.exists(src):
        rows = list(csv.reader(open(src)))
        h = rows[1]
        ia, isrc, iss = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
        ops, samp, tot = collections.Counter(), collections.Counter(), 0
        for r in rows[2:]:
            if r and r[0].startswith("Kernel"):
                break
            try:
                n = int(r[ia])
            except (ValueError, IndexError):
                continue
            toks = r[isrc].split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            ops[op] += n
            samp[op] += int(r[iss])
            tot += n
        lines.append(f"# SASS opcode histogram (warp-level instructions executed, first captured launch): total {tot}")
        for op, n in ops.most_common(24):
            lines.append(f"  {op:10s} {n:12d} {100 * n / tot:5.1f}%   stall samples {samp[op]}")
    with open(os.path.join(PROF, f"{tag}_{kernel}_ncu.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--kernel", default="chan256")
    ap.add_argument("--samples-per-launch", type=int, default=None)
    ap.add_argument("--launches-csv", default=None, help="only summarise this launch list (file name under gpurun_out/)")
    ap.add_argument("--title", default=None)
    a = ap.parse_args()
    os.makedirs(PROF, exist_ok=True)
    if a.launches_csv:
        launches(a.tag, a.launches_csv, a.title)
    else:
        launches(a.tag)
        full(a.tag, a.kernel, a.samples_per_launch)
    print(sorted(os.listdir(PROF)))
