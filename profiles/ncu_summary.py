#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs."""
import csv, io, re, subprocess, sys
KEYS = r'gpu__time_duration\.sum$|dram__bytes_(read|write)\.sum$|gpu__dram_throughput\.avg\.pct|sm__pipe_tensor.*cycles_active.*pct|sm__inst_executed_pipe_(fp64|tensor|uniform|alu|fma|fmaheavy|xu|lsu|tmem)[a-z_]*\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.pct|launch__registers_per_thread$|launch__grid_size|launch__block_size|sm__throughput\.avg\.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct|smsp__cycles_active\.avg$|sm__cycles_elapsed\.max|lts__t_bytes\.sum$|lts__throughput\.avg\.pct|l1tex__throughput\.avg\.pct|smsp__warp_issue_stalled.*_per_warp_active\.pct|sm__memory_throughput|smsp__average_warp.*stall|sm__cycles_active\.avg$|launch__shared_mem_per_block_dynamic|tmem|tensor'
def main(path, pat=KEYS):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('## kernel:', d.get('Kernel Name', '')[:100], 'grid', d.get('Grid Size'), 'block', d.get('Block Size'))
        for k, u in zip(hdr, units):
            if re.search(pat, k) and d[k] not in ('', 'n/a'):
                print(f'{k} [{u}] = {d[k]}')
if __name__ == '__main__':
    main(*sys.argv[1:])
