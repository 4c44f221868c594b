"""Turn one `ncu --set full --import-source on` report (read here, with the ncu CLI of this container) into the
text files kept under profiles/:

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01 [workload]

  <prefix>_ncu_full_summary.txt          selected metrics of every captured launch
  <prefix>_<kernel>_source_hotspots.txt  top source lines by stall samples / executed warp instructions
  profiles/traffic.json                  dram read + write bytes per launch (bench.py's roofline.traffic), for `workload`
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg",
]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True, check=True).stdout


def short(name):
    m = re.search(r"(k_[a-z_]+)", name)
    return m.group(1) if m else name[:24]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(value.replace(",", "")) * scale


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    workload = sys.argv[3] if len(sys.argv) > 3 else "truck_4k_dof"
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    traffic = {}
    with open(prefix + "_ncu_full_summary.txt", "w") as f:
        f.write(f"# {os.path.basename(rep)}: ncu --set full --clock-control none --import-source on (cold caches, serialised launches)\n")
        for r in data:
            f.write(f"---- {r[ki][:70]}\n")
            for m in METRICS:
                if m in hdr:
                    f.write(f"  {m:84s} {r[hdr.index(m)]} {units[hdr.index(m)]}\n")
            k = short(r[ki])
            if k not in traffic:
                i0, i1 = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                traffic[k] = int(to_bytes(r[i0], units[i0]) + to_bytes(r[i1], units[i1]))
    tpath = os.path.join(os.path.dirname(prefix), "traffic.json")
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    tj[workload] = traffic
    json.dump(tj, open(tpath, "w"), indent=1)
    seen = set()
    for r in data:
        k = short(r[ki])
        if k in seen:
            continue
        seen.add(k)
        src = ncu("-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{k}",
                  "--launch-skip", "0", "--launch-count", "1")
        cur, h, agg = None, None, collections.OrderedDict()
        for row in csv.reader(io.StringIO(src)):
            if not row:
                continue
            if row[0] == "File Path":
                cur = os.path.basename(row[1]); continue
            if row[0] == "Line No":
                h = row; si, ii = h.index("# Samples"), h.index("Instructions Executed"); continue
            if h is None or len(row) <= max(si, ii) or row[2] != "-":
                continue
            try:
                key = (cur, int(row[0]), row[1].strip()[:110])
            except ValueError:
                continue
            a = agg.setdefault(key, [0, 0])
            a[0] += int(row[si] or 0); a[1] += int(row[ii] or 0)
        ts, ti = max(1, sum(v[0] for v in agg.values())), max(1, sum(v[1] for v in agg.values()))
        with open(f"{prefix}_{k}_source_hotspots.txt", "w") as f:
            f.write(f"{k}: {ti} warp instructions, {ts} stall samples; lines sorted by samples\n")
            for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
                f.write(f"{100 * v[0] / ts:5.1f}% smp {100 * v[1] / ti:5.1f}% ins  {key[0]}:{key[1]}  {key[2]}\n")
    print("kernels:", sorted(seen), "traffic:", traffic)


if __name__ == "__main__":
    main()
