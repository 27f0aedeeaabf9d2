#!/usr/bin/env python
"""Key per-kernel metrics from an .ncu-rep (run where ncu is installed; no GPU needed).
   usage: ncu_raw_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_bytes.sum']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        vals = [r[i] for r in rows[2:]]
        if w == 'Kernel Name':
            vals = [v.replace('void pylda::', '').replace('(pylda::EParams)', '') for v in vals]
        print("%-66s %-14s %s" % (w, units[i], " | ".join(vals)))
