"""Print a handful of ncu raw-page metrics of a report: python tools/ncu_keys.py report.ncu-rep [more substrings]"""
import csv, subprocess, sys
rep = sys.argv[1]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'sm__inst_executed_pipe_fp64.avg.pct',
        'stalled_long_scoreboard_per', 'stalled_math_pipe', 'stalled_short_scoreboard_per', 'stalled_barrier_per', 'stalled_mio',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct', 'registers_per_thread ', 'stalled_wait_per', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct', 'smsp__issue_active.avg.pct', 'stalled_lg_throttle_per', 'stalled_not_selected_per',
        'sm__throughput.avg.pct', 'smsp__inst_executed.sum ', 'stalled_dispatch', 'stalled_no_instruction', 'stalled_branch',
        'sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct', 'dram__throughput.avg.pct'] + sys.argv[2:]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if any(k in h + ' ' for k in keep):
            print("  %-90s %-10s %s" % (h, u, v))
