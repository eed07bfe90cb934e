"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.md
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_push.md "title" """
import csv, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
        "sm__sass_inst_executed_op_shared_atom.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_sample_count"]
with open(out, "w") as f:
    f.write(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none --import-source on)\n\n")
    for r in rows[2:]:
        f.write("| metric | value | unit |\n|---|---|---|\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"| {k} | {r[i]} | {units[i]} |\n")
        f.write("\n")
print(open(out).read())
