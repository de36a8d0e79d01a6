#!/usr/bin/env python3
"""SASS evidence of the shipped library: instruction counts per kernel from `cuobjdump -sass` -> profiles/<prefix>_sass_summary.md
python tools/sass_summary.py [prefix] [label]"""
import collections, re, subprocess, sys
prefix = sys.argv[1] if len(sys.argv) > 1 else 'r2_v8'
label = sys.argv[2] if len(sys.argv) > 2 else 'Round 2, step program v8 (3xFP16 split, pre-split operand images)'
LIB = 'bayesianautoencoders_b200/lib/libbae_b200.so'
sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
filt = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
names = iter(filt)
COLS = ['UTCHMMA', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'LDTM', 'UBLKCP', 'USETMAXREG', 'SYNCS', 'UCGABAR', 'F2FP', 'HADD2.F32', 'REDUX',
        'LDS.128', 'STS.128', 'STS.64', 'LDL', 'STL', 'MUFU', 'FFMA', 'HMMA', 'IMMA']
per, variants, cur = collections.OrderedDict(), {}, None
for line in sass.split('\n'):
    if 'Function : ' in line:
        full = next(names)
        cur = re.sub(r'\(int\)', '', full.replace('void ', '').replace('bae::', '')).split('(')[0]
        per[cur] = collections.Counter(); variants[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if not m or cur is None: continue
    op = m.group(1)
    per[cur]['instructions'] += 1
    for c in COLS:
        if op == c or op.startswith(c + '.') or op.startswith(c + '_') or (c in ('LDS.128', 'STS.128', 'STS.64') and op.startswith(c.split('.')[0]) and op.endswith(c.split('.')[1])):
            per[cur][c] += 1
    if op.split('.')[0] in ('UTCHMMA', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'LDTM', 'UBLKCP', 'UCGABAR_ARV', 'UCGABAR_WAIT', 'USETMAXREG'):
        variants[cur][op] += 1
rows = [f"| `{k}` | " + " | ".join(str(v[c]) for c in COLS) + f" | {v['instructions']} |" for k, v in sorted(per.items()) if 'bae_' in k]
show = [k for k in per if k.startswith('bae_phase_kernel_tc<0, -1, 3>') or k.startswith('bae_phase_kernel_tc<3, 5, 1>')]
md = f"""# {label} - SASS evidence of the shipped library

`cuobjdump -sass {LIB}` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a`), instruction counts per kernel (`tools/sass_summary.py`).
UTCHMMA = tcgen05.mma (`.2CTA`: cta_group::2), UTCBAR = tcgen05.commit (multicast to both CTAs of the pair), UTMALDG / UTMASTG / UTMAREDG = TMA tensor load /
store / reduce-add, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (operand-image tiles), USETMAXREG = setmaxnreg, SYNCS = mbarrier operations, UCGABAR = cluster barrier,
F2FP.F16.F32.PACK_AB / HADD2.F32 = the FP16 hi / lo split of the converter warps, REDUX = warp-wide integer max of the magnitude bounds.
`bae_phase_kernel_tc<kindA, kindB, mode>`: mode 0 = generic front half, 1 = B operand images, 2 = both operands images (no conversion code),
3 = B image + converted A tiles stored to an activation image (the extra UTMASTG).

| kernel | """ + " | ".join(COLS) + """ | instructions |
|---|""" + "---:|" * (len(COLS) + 1) + "\n" + "\n".join(rows) + "\n"
for k in show:
    md += f"\nInstruction variants in `{k}`:\n\n" + "\n".join(f"* `{op}` x {n}" for op, n in sorted(variants[k].items())) + "\n"
open(f'profiles/{prefix}_sass_summary.md', 'w').write(md)
print("written", len(rows), "kernels")
