"""Condense an `ncu --set full` report into the JSON summary committed under profiles/.

Usage (in the container, after bringing the .ncu-rep back in gpurun_out/):
    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv profiles/rNN_vK_ncu_summary.json
One entry per distinct kernel name (first captured launch), values with their ncu units; with --all one entry
per captured launch ("launch_index" = position in the capture): kernels such as k_copy_chunks run once per callback
with different sizes.
"""
import csv
import json
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(src, dst, keep_all=False):
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    out, seen = [], set()
    for r in rows[hdr + 2:]:
        rec = dict(zip(names, r))
        k = rec["Kernel Name"]
        if k in seen and not keep_all:
            continue
        seen.add(k)
        e = {"Kernel Name": k, "launch_index": len(out)}
        tot = 0.0
        for m in KEEP:
            if m in rec:
                u = units[names.index(m)]
                e[m] = f"{rec[m]} {u}".strip()
                if m.startswith("dram__bytes"):
                    tot += float(rec[m].replace(",", "")) * UNIT.get(u, 1.0)
        e["traffic_bytes_per_launch"] = int(tot)
        out.append(e)
    json.dump(out, open(dst, "w"), indent=1)
    for e in out:
        print(e["Kernel Name"][:70], e.get("gpu__time_duration.sum"), e["traffic_bytes_per_launch"])


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "--all"]
    main(args[0], args[1], keep_all="--all" in sys.argv)
