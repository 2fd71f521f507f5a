"""Prints the handful of metrics the round notes quote from an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv, subprocess, sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
STALL = 'smsp__warp_issue_stalled_'
for f in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h = r[0]
    print('==', f)
    for row in r[2:]:
        for w in WANT:
            if w in h:
                print('  %-70s %s %s' % (w, row[h.index(w)], r[1][h.index(w)]))
        st = [(float(row[i].replace(',', '')), h[i][len(STALL):].replace('_per_warp_active.pct', '')) for i in range(len(h))
              if h[i].startswith(STALL) and h[i].endswith('_per_warp_active.pct') and row[i]]
        for v, n in sorted(st, reverse=True)[:8]:
            print('  stall %-40s %.1f' % (n, v))
