"""Print selected metrics from an .ncu-rep (via `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys, re
rep = sys.argv[1]
pats = sys.argv[2:] or ['gpu__time_duration.sum', 'dram__bytes_read.sum$', 'dram__bytes_write.sum$', 'dram__throughput.avg.pct', 'gpu__dram_throughput',
                        'sm__warps_active.avg.pct', 'launch__registers_per_thread', 'launch__occupancy_limit', 'sm__throughput.avg.pct', 'smsp__issue_active.avg.pct',
                        'smsp__inst_executed.sum$', 'lts__t_sector_hit_rate', 'launch__waves', 'achieved_occupancy', 'sm__pipe_tensor', 'smsp__cycles_active.avg$', 'gpc__cycles_elapsed.max', 'stall', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum$']
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('==', r[4][:90], r[7], r[8])
    for i, h in enumerate(hdr):
        if any(re.search(p, h) for p in pats):
            print(f'  {h} [{units[i]}] = {r[i]}')
