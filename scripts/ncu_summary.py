"""Summarise an .ncu-rep (ncu --set full) into the text block kept under profiles/: per kernel the duration, DRAM
bytes, L2 sectors, L1/LSU utilisation, issue utilisation, instructions, occupancy, registers, shared memory and the
warp stall reasons above 0.3 cycles per issued instruction.  Usage: python scripts/ncu_summary.py REP [title]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
idx = {h: i for i, h in enumerate(hdr)}
print(title)
for r in rows[2:]:
    print("----")
    for w in want:
        if w in idx:
            print(f"{w:90s} {r[idx[w]]} {units[idx[w]]}")
    for h in hdr:
        if "smsp__average_warp" in h and "issue_stalled" in h and h.endswith(".ratio") and "not_issued" not in h:
            v = float(r[idx[h]] or 0)
            if v > 0.3:
                print(f"stall {h.split('issue_stalled_')[1][:-len('_per_issue_active.ratio')]:40s} {v:.6f}")
