"""usage: ncu_sum.py report.ncu-rep  -- key metrics + stall ratios for every kernel in the report"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); h = rows[0]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for r in rows[2:]:
    d = dict(zip(h, r))
    print("==", d.get('Kernel Name', '')[:60])
    for k in want:
        if k in d: print("  ", k, d[k])
    st = []
    for k, v in d.items():
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
            try:
                if float(v) > 0.2: st.append((float(v), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError: pass
    print("   stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)))
