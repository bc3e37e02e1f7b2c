#!/usr/bin/env python
"""Condense the per-config `ncu --set full` raw CSVs (tools/profile_configs.py) into
profiles/r02_configs_ncu_summary.json: {config: {callback: fp64 pipe %}, "kernels": {config: [per-kernel metrics]}}.

    for k in 3 4 5; do ncu -i gpurun_out/r02_cfg$k.ncu-rep --page raw --csv > /tmp/cfg$k.csv; done
    python tools/ncu_configs_summary.py /tmp/cfg3.csv:3 /tmp/cfg4.csv:4 /tmp/cfg5.csv:5 profiles/r02_configs_ncu_summary.json
"""
import csv
import json
import re
import sys

from ncu_summary import KEEP, UNIT

FP64 = "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"
# VecOp template argument -> callback (kernels.cuh enum VecOp)
VEC = {0: "obj", 1: "grad!", 2: "cons!", 3: "jprod!", 4: "jtprod!", 5: "hprod!", 6: "objgrad!", 7: "objcons!"}


def callback_of(kernel):
    m = re.search(r"k_vector\w*<.*, *\(?(?:int\))?(\d+)>", kernel)
    if m:
        return VEC.get(int(m.group(1)))
    if "k_jac_coord" in kernel or "k_jac_cells_cg" in kernel:
        return "jac_coord!"
    if "k_hess_coord" in kernel or "k_hess_cells_cg" in kernel:
        return "hess_coord!"
    return None  # (k_copy_chunks serves both coordinate callbacks: listed under its own name)


def main(argv):
    out = {"kernels": {}}
    for spec in argv[:-1]:
        path, cfg = spec.rsplit(":", 1)
        rows = list(csv.reader(open(path)))
        hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        names, units = rows[hdr], rows[hdr + 1]
        ks, per_cb = [], {}
        # one record per kernel name: its LONGEST launch (a kernel may also run on a one-tile sub-model — the template
        # builder — or on the create-time self-check's sample)
        best = {}
        for r in rows[hdr + 2:]:
            rec = dict(zip(names, r))
            dur = float(rec.get("gpu__time_duration.sum", "0").replace(",", "") or 0)
            if rec["Kernel Name"] not in best or dur > best[rec["Kernel Name"]][0]:
                best[rec["Kernel Name"]] = (dur, rec)
        # (longest first: a callback's FP64 figure is that of its longest arithmetic kernel, not of the one-tile image launch)
        for k, (_, rec) in sorted(best.items(), key=lambda kv: -kv[1][0]):
            e = {"Kernel Name": k}
            tot = 0.0
            for m in KEEP:
                if m in rec:
                    u = units[names.index(m)]
                    e[m] = f"{rec[m]} {u}".strip()
                    if m.startswith("dram__bytes"):
                        tot += float(rec[m].replace(",", "")) * UNIT.get(u, 1.0)
            e["traffic_bytes_per_launch"] = int(tot)
            cb = callback_of(k)
            e["callback"] = cb
            ks.append(e)
            if cb and FP64 in rec and cb not in per_cb:
                per_cb[cb] = float(rec[FP64].replace(",", ""))
            if cb is None:  # kappa-block / fill / reduction kernels: listed under their own name
                e["callback"] = None
        out["kernels"][cfg] = ks
        out[cfg] = per_cb
    json.dump(out, open(argv[-1], "w"), indent=1)
    for cfg in out:
        if cfg != "kernels":
            print(cfg, out[cfg])


if __name__ == "__main__":
    main(sys.argv[1:])
