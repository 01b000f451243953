#!/usr/bin/env python
"""Per kernel of libqce_b200.so: counts of the SASS mnemonics that show which hardware paths it uses (cuobjdump -sass).
    python profiles/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'quantized_channel_estimation_b200', 'libqce_b200.so')
KEYS = (('UTCHMMA', r'\bUTCHMMA'), ('UTCBAR', r'\bUTCBAR'), ('LDTM', r'\bLDTM'), ('STTM', r'\bSTTM'), ('UBLKCP', r'\bUBLKCP'),
        ('HMMA', r'\bHMMA'), ('STG256', r'\bSTG\.E\.[A-Z0-9.]*256'), ('LDG256', r'\bLDG\.E\.[A-Z0-9.]*256'))


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        for key, pat in KEYS:
            if re.search(pat, line):
                cur[key] += 1
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print('# SASS evidence, round 2 (final tree): per kernel of libqce_b200.so the number of tcgen05 MMA (UTCHMMA), tcgen05 commit (UTCBAR), TMEM load / store')
    print('# (LDTM / STTM), bulk-TMA copy (UBLKCP), legacy mma.sync (HMMA) and 256-bit global store / load (STG256 / LDG256) instructions.')
    print('# cuobjdump -sass quantized_channel_estimation_b200/libqce_b200.so  (profiles/sass_evidence.py)')
    print(f'# totals over {len(per)} kernels: ' + ' '.join(f'{k}={tot[k]}' for k, _ in KEYS))
    for name in sorted(per):
        print(name + ' ' + ' '.join(f'{k}={per[name][k]}' for k, _ in KEYS))


if __name__ == '__main__':
    main()
