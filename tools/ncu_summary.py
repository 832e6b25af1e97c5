#!/usr/bin/env python
"""Print the judged subset of an ncu report as a markdown table (and the raw CSV rows of that subset).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--csv out.csv]

Reads `ncu -i REPORT --page raw --csv`; one column block per captured launch.
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.sum",
    "sm__inst_executed_pipe_tma.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__inst_executed_op_global_atom.sum", "smsp__inst_executed_op_shared_atom.sum",
    "sm__sass_inst_executed_op_global_ld.sum", "sm__sass_inst_executed_op_global_st.sum",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "sm__ctas_launched.sum",
]


def main():
    rep = sys.argv[1]
    out_csv = sys.argv[sys.argv.index("--csv") + 1] if "--csv" in sys.argv else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    name_i = col.get("Kernel Name")
    keep = [h for h in WANT if h in col]
    print("| metric | unit | " + " | ".join(f"`{(l[name_i] if name_i is not None else '?')[:48]}`" for l in launches) + " |")
    print("|---|---|" + "---|" * len(launches))
    for h in keep:
        print(f"| {h} | {units[col[h]]} | " + " | ".join(l[col[h]] for l in launches) + " |")
    if out_csv:
        with open(out_csv, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["Kernel Name"] + keep)
            w.writerow([""] + [units[col[h]] for h in keep])
            for l in launches:
                w.writerow([l[name_i] if name_i is not None else ""] + [l[col[h]] for h in keep])


if __name__ == "__main__":
    main()
