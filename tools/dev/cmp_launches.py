"""Side-by-side per-launch durations of two ncu launch lists (first iteration), aligned on the kernels both contain."""
import csv, sys
def load(f):
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    h = rows[0]; iv = h.index('Metric Value'); ik = h.index('Kernel Name')
    return [(r[ik].replace('void ', '').split('(')[0], float(r[iv].replace(',', '')) / 1000) for r in rows[1:]]
a, b = load(sys.argv[1]), load(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 62
extra = ('bae_wconv_kernel',)
ia = ib = 0; ta = tb = ga = gb = 0.0
while ia < len(a) and ib < len(b) and ib < n:
    if a[ia][0] in extra and b[ib][0] not in extra: print("%-40s %9.1f %9s" % (a[ia][0], a[ia][1], '')); ta += a[ia][1]; ia += 1; continue
    if b[ib][0] in extra and a[ia][0] not in extra: print("%-40s %9s %9.1f" % (b[ib][0], '', b[ib][1])); tb += b[ib][1]; ib += 1; continue
    print("%-40s %9.1f %9.1f %6.3f" % (a[ia][0], a[ia][1], b[ib][1], b[ib][1] / a[ia][1]))
    ta += a[ia][1]; tb += b[ib][1]
    if 'phase_kernel_tc' in a[ia][0]: ga += a[ia][1]; gb += b[ib][1]
    ia += 1; ib += 1
print("total %.1f %.1f  gemm %.1f %.1f (%.3f)" % (ta, tb, ga, gb, gb / ga))
