#!/usr/bin/env python
"""ncu report -> compact per-launch CSV of the metrics DESIGN.md / bench.py quote.  usage: ncu_summary.py report.ncu-rep > out.csv"""
import csv, io, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, data = rows[0], rows[1], rows[2:]
keep = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed']
idx = [(k, h.index(k)) for k in keep if k in h]
w = csv.writer(sys.stdout)
w.writerow([k + (f' [{u[i]}]' if u[i] else '') for k, i in idx])
for r in data:
    w.writerow([r[i][:90] for _, i in idx])
