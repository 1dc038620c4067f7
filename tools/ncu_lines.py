"""Join the SASS page of an ncu report with nvdisasm line info: instructions executed and stall samples per SOURCE LINE.
usage: python tools/ncu_lines.py report.ncu-rep lib.so kernel_substring [min_pct]"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, so, pat = sys.argv[1:4]
min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
line_of = {}
for cub in glob.glob(os.path.join(tmp, '*.cubin')):
    dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout
    cur_fn, cur_line, on = None, None, False
    for ln in dis.splitlines():
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
        if m:
            on = pat in m.group(1)
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            line_of[int(m.group(1), 16)] = (cur_line, m.group(2))
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg, hdr, on = {}, None, False
base = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        on = pat_h = any(pat_part in r[1] for pat_part in [sys.argv[5]] ) if len(sys.argv) > 5 else True
        base = None
        continue
    if r and r[0] == 'Address':
        hdr = r
        continue
    if not on or hdr is None or len(r) < len(hdr):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    key = line_of.get(addr - base, (None, r[1]))[0]
    a = agg.setdefault(key, [0, 0, {}])
    a[0] += int(r[hdr.index('Instructions Executed')] or 0)
    a[1] += int(r[hdr.index('# Samples')] or 0)
    for name in ('stall_wait', 'stall_short_sb', 'stall_long_sb', 'stall_math', 'stall_not_selected', 'stall_branch_resolving', 'stall_no_inst', 'stall_dispatch', 'stall_mio'):
        a[2][name] = a[2].get(name, 0) + int(r[hdr.index(name)] or 0)
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print('total instructions %d, samples %d' % (ti, ts))
for key, a in sorted(agg.items(), key=lambda kv: (kv[0] is None, kv[0])):
    if 100.0 * a[0] / ti >= min_pct or 100.0 * a[1] / ts >= min_pct:
        top = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
        print('%-28s inst %5.2f%%  samples %5.2f%%  %s' % (key, 100.0 * a[0] / ti, 100.0 * a[1] / ts,
                                                          ' '.join('%s=%d' % (k[6:], v) for k, v in top if v)))
