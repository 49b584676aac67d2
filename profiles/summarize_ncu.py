#!/usr/bin/env python
"""usage: python profiles/summarize_ncu.py <report.ncu-rep>  -- key metrics + top stall reasons per kernel launch"""
import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]; idx={h:i for i,h in enumerate(hdr)}
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','l1tex__t_bytes.sum','smsp__inst_executed.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__waves_per_multiprocessor','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu.sum','smsp__cycles_active.avg','sm__cycles_active.avg','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed']
stall=[h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print('====', r[idx['Kernel Name']][:70])
    for w in want:
        if w in idx: print(f'  {w:62s} {r[idx[w]][:20]:>20s} {units[idx[w]]}')
    vals=sorted([(float(r[idx[h]]) if r[idx[h]] else 0,h) for h in stall],reverse=True)[:6]
    print('  stalls:', ', '.join('%s=%.2f'%(h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v) for v,h in vals))
