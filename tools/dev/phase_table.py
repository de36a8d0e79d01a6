"""Per-phase durations from ncu launch lists (gpu__time_duration.sum csv): python tools/dev/phase_table.py a.csv [b.csv ...]"""
import csv, sys
NAMES = ['PRO', 'F1', 'F2', 'F3', 'BNF', 'F4', 'F5', 'F6', 'B6', 'B5', 'B4', 'BNB', 'B3', 'B2', 'B1', 'ADAM1', 'F1b', 'F2b', 'F3b',
         'BNFb', 'F4b', 'F5b', 'F6b', 'B6b', 'B5b', 'B4b', 'BNBb', 'B3b', 'B2b', 'B1b', 'ADAM2', 'T1', 'T2', 'T3', 'BNT', 'T4', 'T5', 'T6', 'EPI',
         'SWD', 'SWG', 'SWS']
cols = []
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    h, data = rows[0], rows[1:]
    iv = h.index('Metric Value')
    cols.append([float(r[iv].replace(',', '')) / 1000 for r in data][:42])
print('%-6s' % 'phase' + ''.join('%12s' % f.split('/')[-1][:11] for f in sys.argv[1:]))
for i, n in enumerate(NAMES):
    print('%-6s' % n + ''.join('%12.1f' % c[i] for c in cols if i < len(c)))
print('%-6s' % 'sum' + ''.join('%12.1f' % sum(c) for c in cols))
gemm = [i for i, n in enumerate(NAMES) if n[0] in 'FBT' and not n.startswith('BN')]
print('%-6s' % 'gemm' + ''.join('%12.1f' % sum(c[i] for i in gemm) for c in cols))
