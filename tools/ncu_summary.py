#!/usr/bin/env python
"""Turns an `ncu --set full` report into the committed evidence: profiles/<name>.txt (selected raw metrics, one kernel per
block), profiles/<name>_hot_lines.txt (samples / instructions by CUDA source line) and an entry of profiles/ncu_summary.json
(per-launch DRAM traffic and issue-slot utilisation, which bench.py copies into `roofline`).
    python tools/ncu_summary.py gpurun_out/x.ncu-rep <name> [<config>_<precision>_<nenv> [kernel-substring]]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def main():
    rep, name = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    want = sys.argv[4] if len(sys.argv) > 4 else ""
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out, summary = [f"# ncu --set full --clock-control none; source {os.path.basename(rep)}; cold-cache, serialised launch (compare shares, not absolutes)"], None
    for r in rows[2:]:
        kn = r[hdr.index("Kernel Name")]
        out.append(f"## {kn[:150]}")
        vals = {}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                out.append(f"{k:95s} {r[i]:>18s} {units[i]}")
                vals[k] = (float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else None, units[i])
        if key and want in kn and summary is None:
            dr, dw = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
            summary = {"dram_bytes_per_launch": dr[0] * UNIT[dr[1]] + dw[0] * UNIT[dw[1]], "dram_read_bytes": dr[0] * UNIT[dr[1]],
                       "dram_write_bytes": dw[0] * UNIT[dw[1]],
                       "issue_active_pct": vals["smsp__issue_active.avg.pct_of_peak_sustained_active"][0],
                       "warps_active_pct": vals["sm__warps_active.avg.pct_of_peak_sustained_active"][0],
                       "duration_us_under_ncu": vals["gpu__time_duration.sum"][0], "registers": vals["launch__registers_per_thread"][0],
                       "kernel": kn[:120], "source": f"profiles/{name}.txt"}
    open(os.path.join(ROOT, "profiles", name + ".txt"), "w").write("\n".join(out) + "\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    tmp = "/tmp/_ncu_src.csv"
    open(tmp, "w").write(src)
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), tmp, "40"], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", name + "_hot_lines.txt"), "w").write(hot)
    if key and summary:
        p = os.path.join(ROOT, "profiles", "ncu_summary.json")
        d = json.load(open(p)) if os.path.exists(p) else {}
        d[key] = summary
        json.dump(d, open(p, "w"), indent=1, sort_keys=True)
        print(key, summary)
    print("\n".join(out[:40]))


if __name__ == "__main__":
    main()
