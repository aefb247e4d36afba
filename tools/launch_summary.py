"""Per-launch summary of an ncu --csv launch list (time, DRAM bytes, optional L2 sectors / warp
instructions): python tools/launch_summary.py FILE.csv [LAST_N] [NAME_REGEX]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
last = int(sys.argv[2]) if len(sys.argv) > 2 else 40
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
d = collections.OrderedDict()
for x in csv.DictReader(lines):
    k = x['ID']
    d.setdefault(k, {'name': re.sub(r'\(.*', '', x['Kernel Name'])[-60:], 'grid': x['Grid Size']})
    d[k][x['Metric Name']] = (float(x['Metric Value'].replace(',', '')), x['Metric Unit'])
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'sector': 1, 'inst': 1}
keys = [k for k in d if pat is None or pat.search(d[k]['name'])]
for k in keys[-last:]:
    v = d[k]
    get = lambda m: v[m][0] * scale.get(v[m][1], 1) if m in v else None
    t = get('gpu__time_duration.sum')
    rd, wr = get('dram__bytes_read.sum'), get('dram__bytes_write.sum')
    out = [k, v['name'], v['grid'], '%.3f ms' % t]
    if rd is not None:
        out.append('dram rd %.0f MB wr %.0f MB (%.2f TB/s)' % (rd / 1e6, wr / 1e6, (rd + wr) / t / 1e9))
    if get('lts__t_sectors.sum') is not None:
        out.append('L2 %.0f MB' % (get('lts__t_sectors.sum') * 32 / 1e6))
    if get('smsp__inst_executed.sum') is not None:
        out.append('inst %.1fM' % (get('smsp__inst_executed.sum') / 1e6))
    print(*out)
