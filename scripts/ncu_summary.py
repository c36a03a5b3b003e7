"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + top SASS instructions by stall samples."""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum',
        'smsp__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed']
print("metric,unit,value")
for i, n in enumerate(h):
    if n in keep or ('issue_stalled' in n and n.endswith('ratio')):
        print(f"{n},{u[i]},{v[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
tot = sum(int(r[isamp]) for r in data)
totex = sum(int(r[iex]) for r in data)
print(f"# total stall samples {tot}; SASS instructions {len(data)}; warp-instructions executed {totex}")
print("# top instructions by samples: index,samples,pct,executed,sass")
for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][isamp]))[:top]:
    print(f"{k},{r[isamp]},{100 * int(r[isamp]) / tot:.1f}%,{r[iex]},{r[isrc].strip()}")
# samples by opcode class
cls = {}
for r in data:
    op = re.sub(r'^@!?U?P\d\s+', '', r[isrc].strip()).split()[0].split('.')[0]
    cls[op] = cls.get(op, 0) + int(r[isamp])
print("# samples by opcode:", ", ".join(f"{k}={100 * s / tot:.1f}%" for k, s in sorted(cls.items(), key=lambda x: -x[1])[:14]))
