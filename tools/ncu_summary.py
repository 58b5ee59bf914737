#!/usr/bin/env python3
"""Summarise an `ncu --page raw --csv` dump: one block per kernel launch with the metrics the roofline needs."""
import csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))       # (an ncu --log-file starts with ==PROF== lines)
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_tex_wavefronts.sum',
        'l1tex__f_wavefronts.sum', 'l1tex__texin_requests.sum', 'l1tex__t_requests_pipe_tex.sum', 'l1tex__t_sectors_pipe_tex.sum',
        'sm__inst_executed_pipe_tex.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max']
sel = sys.argv[2] if len(sys.argv) > 2 else ''
seen = {}
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    if sel and sel not in name: continue
    seen[name] = seen.get(name, 0) + 1
    if seen[name] > int(sys.argv[3]) if len(sys.argv) > 3 else seen[name] > 1: continue
    print('====', name, '#', seen[name])
    for w in want:
        if w in idx and r[idx[w]] != '': print(f"  {w:72s} {r[idx[w]]:>18s} {units[idx[w]]}")
    stall = sorted([(float(r[idx[h]] or 0), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for h in hdr
                    if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h], reverse=True)[:6]
    print('  stalls (warps per issue):', ', '.join(f"{n}={v:.2f}" for v, n in stall))
