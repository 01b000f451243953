#!/usr/bin/env python
"""DRAM traffic per pilot of the hot kernels from an ncu report (`ncu --set full ... -o X` on the GPU box, read here with
`ncu -i X.ncu-rep --page raw --csv`): writes / updates profiles/r02_traffic.json, which bench.py reads for `roofline.traffic`.

    python tools/ncu_traffic.py gpurun_out/r02_dense.ncu-rep --pilots 1048576 --kernel dense_tc_kernel
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles', 'r02_traffic.json')


def to_bytes(value, unit):
    v = float(value.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[unit]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--pilots', type=int, required=True, help='pilots one captured launch processed')
    ap.add_argument('--kernel', required=True, help='substring of the kernel name; the LAST matching launch of the report is used')
    ap.add_argument('--algorithmic-bytes-per-pilot', type=float, default=None)
    a = ap.parse_args()
    raw = subprocess.run(['ncu', '-i', a.report, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(head)}
    pick = None
    for r in rows[2:]:
        if a.kernel in r[col['Kernel Name']]:
            pick = r
    if pick is None:
        raise SystemExit(f'no launch of {a.kernel} in {a.report}')

    def metric(name):
        return to_bytes(pick[col[name]], units[col[name]])
    rd, wr = metric('dram__bytes_read.sum'), metric('dram__bytes_write.sum')
    dur = float(pick[col['gpu__time_duration.sum']].replace(',', ''))
    dur_ms = dur * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(units[col['gpu__time_duration.sum']], 1e-6)
    table = json.load(open(OUT)) if os.path.exists(OUT) else {}
    table[a.kernel] = {'dram_bytes_read': rd, 'dram_bytes_write': wr, 'pilots': a.pilots, 'dram_bytes_per_pilot': (rd + wr) / a.pilots,
                       'algorithmic_bytes_per_pilot': a.algorithmic_bytes_per_pilot, 'duration_ms_under_ncu': dur_ms,
                       'kernel_name': pick[col['Kernel Name']][:160],
                       'source': f'ncu --set full capture {os.path.basename(a.report)} (dram__bytes_read.sum + dram__bytes_write.sum of one launch)'}
    json.dump(table, open(OUT, 'w'), indent=1)
    print(json.dumps(table[a.kernel]))


if __name__ == '__main__':
    main()
