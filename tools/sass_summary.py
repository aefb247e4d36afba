"""SASS evidence for profiles/: per kernel of libsnfft_b200.so the counts of the instructions that
matter for this path (DFMA/DMUL/DADD fp64 pipe, DMMA fp64 tensor pipe, LDGSTS = cp.async,
UBLKCP = cp.async.bulk (1-D TMA), UBLKPF = bulk L2 prefetch, SYNCS = mbarrier, LDS/STS, BAR, SHFL),
registers and spills (cuobjdump -res-usage).  python tools/sass_summary.py > profiles/sass_summary_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'snfft_b200', '_lib', 'libsnfft_b200.so')
OPS = ['DFMA', 'DMUL', 'DADD', 'DMMA', 'LDGSTS', 'UBLKCP', 'UBLKPF', 'SYNCS', 'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'BAR']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
    return [re.sub(r'\(anonymous namespace\)::|snfft::|\(.*$', '', o) for o in out]


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and cur:
            op = m.group(1)
            counts[cur]['total'] += 1
            for key in OPS:
                if op == key or (key in ('LDG', 'STG', 'LDS', 'STS') and op.startswith(key) and not op.startswith('LDGSTS')
                                 and not op.startswith('LDSM') and not op.startswith('STSM')):
                    if op == key or op[len(key):len(key) + 1] in ('', '.'):
                        counts[cur][key] += 1
    usage = {}
    fn = None
    for line in res.split('\n'):
        m = re.search(r'Function (\S+):', line)
        if m:
            fn = m.group(1)
        m = re.search(r'REG:(\d+).*?SHARED:(\d+)', line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)))
    names = demangle(order)
    print('kernel | SASS instr | ' + ' | '.join(OPS) + ' | regs | static smem')
    print('---|---|' + '---|' * (len(OPS) + 2))
    for fn, nice in sorted(zip(order, names), key=lambda t: t[1]):
        c = counts[fn]
        reg, sh = usage.get(fn, ('?', '?'))
        print(f'{nice} | {c["total"]} | ' + ' | '.join(str(c[k]) for k in OPS) + f' | {reg} | {sh}')


if __name__ == '__main__':
    main()
