#!/usr/bin/env python
"""Summarise an ncu report (one kernel capture, --set full --import-source on) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_shoot_r1.ncu-rep profiles/r1_shoot_kernel.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    lines = []
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        lines.append(f"kernel: {d.get('Kernel Name')}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:72s} {d[k]:>18s} {units[hdr.index(k)]}")
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                try:
                    if float(vals[i]) >= 0.05:
                        lines.append(f"  stall {h[34:-23]:40s} {float(vals[i]):8.3f} warps per issue-active cycle")
                except ValueError:
                    pass
    # source page: opcode mix and stall samples, FP64 vs the rest
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    if len(src) > 2:
        h = src[1]
        ix = {c: i for i, c in enumerate(h)}
        if "Source" in ix and "# Samples" in ix:
            ops, smp = collections.Counter(), collections.Counter()
            for r in src[2:]:
                if len(r) < len(h):
                    continue
                op = re.sub(r"^@!?U?P\d\s+", "", r[ix["Source"]].strip()).split()[0].split(".")[0]
                ops[op] += int(r[ix["Instructions Executed"]] or 0)
                smp[op] += int(r[ix["# Samples"]] or 0)
            tot_i, tot_s = sum(ops.values()) or 1, sum(smp.values()) or 1
            lines.append("  executed warp instructions / stall samples by opcode (top 12):")
            for op, n in ops.most_common(12):
                lines.append(f"    {op:10s} {n:12d} ({100.0 * n / tot_i:5.1f} %)   samples {smp[op]:6d} ({100.0 * smp[op] / tot_s:5.1f} %)")
            fp = sum(n for o, n in ops.items() if o in ("DFMA", "DMUL", "DADD"))
            fps = sum(n for o, n in smp.items() if o in ("DFMA", "DMUL", "DADD"))
            lines.append(f"    FP64 share: {100.0 * fp / tot_i:.1f} % of instructions, {100.0 * fps / tot_s:.1f} % of samples")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    # optional: record the kernel's DRAM traffic per launch for bench.py's roofline.traffic
    #   python tools/ncu_summary.py <rep> <txt> <traffic.json> <key> <units in the captured launch> <workload text>
    if len(sys.argv) >= 7:
        import datetime
        import json
        import os
        jpath, key, units, workload = sys.argv[3], sys.argv[4], float(sys.argv[5]), sys.argv[6]
        d = dict(zip(hdr, rows[2]))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(d[k]) * scale[units_of(hdr, units_row=units_row(rows), key=k)]
        cur = json.load(open(jpath)) if os.path.exists(jpath) else {}
        cur[key] = {"bytes_per_launch": tot, "units_in_capture": units, "bytes_per_unit": tot / units, "workload": workload,
                    "file": os.path.relpath(out), "date": datetime.date.today().isoformat()}
        json.dump(cur, open(jpath, "w"), indent=1)


def units_row(rows):
    return rows[1]


def units_of(hdr, units_row, key):
    return units_row[hdr.index(key)]


if __name__ == "__main__":
    main()
