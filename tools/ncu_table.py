"""Compact per-kernel table from an `ncu --page raw --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = [('gpu__time_duration.sum', 'ms'), ('dram__bytes_read.sum', 'dR_MB'), ('dram__bytes_write.sum', 'dW_MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('launch__registers_per_thread', 'regs'),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu%'), ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu%'),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu%'), ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'), ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'), ('smsp__inst_executed.sum', 'inst'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem_wf'), ('smsp__average_warp_latency_issue_stalled_short_scoreboard_per_warp_active.pct','sb_short'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','st_long'),('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','st_short'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','st_math'),('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','st_bar'),
        ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','st_mio'),('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','st_lg'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','st_wait'),('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','st_nsel')]
want = [(w, s) for w, s in want if w in hdr]
idx = [hdr.index(w) for w, _ in want]; ki = hdr.index('Kernel Name')
print(f"{'kernel':20s}" + "".join(f"{s:>9s}" for _, s in want))
for r in rows[2:]:
    def f(v):
        try: return float(v.replace(',', ''))
        except ValueError: return float('nan')
    v = [f(r[i]) for i in idx]
    print(f"{r[ki].split('(')[0][:20]:20s}" + "".join(f"{x:9.2f}" if abs(x) < 1e6 else f"{x:9.2e}" for x in v))
