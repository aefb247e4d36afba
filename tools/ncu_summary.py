"""Key metrics of an `ncu --set full` report, one block per distinct kernel:
python tools/ncu_summary.py REPORT.ncu-rep [TITLE] > profiles/rN/name.txt  (needs ncu on PATH)."""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sectors.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
if len(sys.argv) > 2:
    print(sys.argv[2])
seen = OrderedDict()
for r in rows[2:]:
    name = r[idx['Kernel Name']]
    if name[:70] in seen:
        continue
    seen[name[:70]] = True
    print('kernel:', name[:110])
    for w in WANT:
        if w in idx:
            print('  %s = %s %s' % (w, r[idx[w]], units[idx[w]]))
    stalls = [(h, float(r[i])) for h, i in idx.items()
              if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')
              and r[i] not in ('', 'n/a')]
    stalls.sort(key=lambda x: -x[1])
    print('  stall cycles per issued instruction (top): ' + ', '.join(
        '%s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v)
        for h, v in stalls[:6]))
