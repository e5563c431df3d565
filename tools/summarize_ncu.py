#!/usr/bin/env python
"""Summarise ncu artefacts from gpurun_out/ into profiles/ (tracked).

  python tools/summarize_ncu.py <tag> <launches.csv> <full.ncu-rep | raw.csv>... [--batch B]

A capture may be given as the .ncu-rep or as its `ncu -i x.ncu-rep --page raw --csv` export (what the GPU box
sends back: the reports themselves exceed gpurun's 64 MiB return limit).  --batch = polynomials per NTT launch
of the profiled command (bench.py default 8192), recorded so that bench.py can scale the traffic to its batch.

Writes profiles/<tag>_launches.md (per-kernel launch times and share of the step),
profiles/<tag>_ncu.md (roofline-relevant metrics per kernel) and refreshes
profiles/traffic.json (DRAM bytes per launch of the forward transform, read by bench.py).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
batch = 8192
if "--batch" in args:
    i = args.index("--batch")
    batch = int(args[i + 1])
    del args[i:i + 2]
tag, launches = args[:2]
reps = args[2:]
sys.path.insert(0, ROOT)
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def short(name):
    name = name.split("(")[0]
    for key in ("ntt_", "elt_kernel"):
        if key in name:
            return name[name.index(key):]
    return name.replace("void ", "")


# ---- launch list
rows = list(csv.reader(open(launches)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg.setdefault(short(r[ki]), []).append(float(r[vi].replace(",", "")))
ours = {k: v for k, v in agg.items() if k.startswith(("ntt_", "elt_"))}
tot = sum(sum(v) for k, v in ours.items() if k.startswith("ntt_"))
with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n"
            "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu`\n"
            "(serialised, cold-cache per-launch times: compare SHARES, not absolutes)\n\n"
            "| kernel | launches | avg ms | share of NTT time |\n|---|---|---|---|\n")
    for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
        share = f"{100 * sum(v) / tot:.1f}%" if k.startswith("ntt_") else "-"
        f.write(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e6:.3f} | {share} |\n")

# ---- full capture
rows = []
for rep in reps:
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    part = list(csv.reader(raw.splitlines()))
    if not rows:
        rows = part
    else:  # align columns by name
        Hp = part[0]
        for r in part[2:]:
            d = dict(zip(Hp, r))
            rows.append([d.get(h, "") for h in rows[0]])
H, U = rows[0], rows[1]
want = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM busiest unit % of peak"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe inst % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: fixed-latency wait / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
]
seen = collections.OrderedDict()
for r in rows[2:]:
    k = short(r[H.index("Kernel Name")])
    if k not in seen:
        seen[k] = r
traffic = {}
with open(os.path.join(out_dir, f"{tag}_ncu.md"), "w") as f:
    f.write(f"# {tag}: ncu --set full --clock-control none (one launch per kernel)\n\n"
            "Workload: bench.py default (N=2^16, 55-bit q, batch 8192 => 4 GiB in, 4 GiB out per NTT kernel;"
            " eltwise n = 2^28, 60-bit q).\n\n")
    for k, r in seen.items():
        f.write(f"## `{k}`\n\n| metric | value |\n|---|---|\n")
        for m, label in want:
            if m in H:
                i = H.index(m)
                f.write(f"| {label} | {r[i]} {U[i]} |\n")
        f.write("\n")
        def num(m):
            return float(r[H.index(m)].replace(",", "")) if m in H and r[H.index(m)] else 0.0
        unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd = num("dram__bytes_read.sum") * unit.get(U[H.index("dram__bytes_read.sum")], 1.0)
        wr = num("dram__bytes_write.sum") * unit.get(U[H.index("dram__bytes_write.sum")], 1.0)
        traffic[k] = rd + wr
fwd = sum(v for k, v in traffic.items() if k.startswith(("ntt_row_fwd", "ntt_pipe_fwd", "ntt_fused_fwd", "ntt_dsmem_fwd"))
          or (k.startswith("ntt_col") and k.rstrip(">").endswith("1")))
import bench  # noqa: E402  (source_hash: the capture is only believed on the kernel sources it was taken on)
json.dump({"tag": tag, "source_hash": bench.source_hash(), "batch": batch, "dram_bytes_per_launch": traffic,
           "ntt_forward_bytes_per_launch": fwd, "ntt_forward_bytes_per_polynomial": fwd / batch,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum from one ncu --set full capture; "
                   "ntt_forward = all kernels of one hexl_b200_ntt_forward call"},
          open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)
print(open(os.path.join(out_dir, f"{tag}_launches.md")).read())
print({k: round(v / 1e9, 3) for k, v in traffic.items()}, "fwd", fwd / 1e9)
