"""SASS mnemonic census of the shipped library: cuobjdump -sass arrayfire_b200/libafb200.so | python tools/sass_census.py"""
import collections
import re
import sys

PREF = ['UBLKCP', 'SYNCS', 'UCGABAR', 'REDUX', 'LDG.E.128', 'STG.E.128', 'ATOMG', 'MEMBAR.ALL.GPU', 'MEMBAR.SC.GPU', 'MEMBAR.ALL.SYS',
        'MEMBAR.SC.SYS', 'UTMALDG', 'UTMASTG', 'UTCHMMA', 'UTCQMMA', 'HMMA', 'IMMA', 'LDGSTS', 'LDS.128', 'STS.128']
WANT = ['pipe_scan_kernelILi0EiiLb1Ej', 'pipe_scan_kernelILi0EffLb0Ei', 'pipe_where_kernelIhE', 'span_kernelINS_9ValPolicyILi0EffEEE',
        'dim_kernelINS_9ValPolicyILi0EffEELi4E', 'dim_scan_kernelILi0Eff', 'rbk_tile_kernelILi0EiiE', 'rbk_count_kernel', 'expr_reduce_kernelIfE']
c = collections.Counter()
nk = total = 0
per = {}
cur = None
rx = re.compile(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)')
for line in sys.stdin:
    if 'Function :' in line:
        nk += 1
        name = line.split('Function :')[1].strip()
        cur = None
        for w in WANT:
            if w in name and w not in per:
                cur = w
                per[w] = collections.Counter()
        continue
    m = rx.match(line)
    if not m:
        continue
    op = m.group(1)
    total += 1
    if cur:
        per[cur]['instructions'] += 1
    for p in PREF:
        if op.startswith(p):
            c[p] += 1
            if cur:
                per[cur][p] += 1
print("kernels", nk, "instructions", total)
for p in PREF:
    print("| `%s*` | %d |" % (p, c[p]))
for w, v in per.items():
    print(w, dict(v))
