#!/usr/bin/env python3
"""Turns the ncu exports of a profiling session into the committed summaries under profiles/.

    ncu -i gpurun_out/prof_program_r1.ncu-rep --page raw --csv > /tmp/raw_prog.csv
    python tools/make_profiles.py gpurun_out/launches_r1_final.csv /tmp/raw_prog.csv [gpurun_out/phases_final.log]
"""
import csv
import json
import sys

launch_csv, raw_csv = sys.argv[1], sys.argv[2]
phases_log = sys.argv[3] if len(sys.argv) > 3 else None
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
h, data = rows[0], rows[1:]
iv, ik = h.index('Metric Value'), h.index('Kernel Name')
vals = [float(r[iv].replace(',', '')) / 1000 for r in data]   # ns -> us
names = ['PRO', 'F1', 'F2', 'F3', 'BNF', 'F4', 'F5', 'F6', 'B6 (dW6|dH5)', 'B5', 'B4', 'BNB', 'B3', 'B2', 'B1', 'ADAM1', 'F1b', 'F2b', 'F3b',
         'BNFb', 'F4b', 'F5b', 'F6b', 'B6b', 'B5b', 'B4b', 'BNBb', 'B3b', 'B2b', 'B1b', 'ADAM2', 'T1', 'T2', 'T3', 'BNT', 'T4', 'T5', 'T6', 'EPI']
with open('profiles/r1_v4_tc_phase_launches.csv', 'w') as f:
    f.write("launch,iteration,phase,kernel,duration_us\n")
    for i, (r, v) in enumerate(zip(data, vals)):
        ph = names[i % 39] if i < 195 else ['SWAP_DECIDE', 'SWAP_GATHER', 'SWAP_SCATTER'][i - 195]
        f.write(f"{i},{min(i // 39, 4)},{ph},{r[ik].split('(')[0]},{v:.3f}\n")
F = {'F1': 2 * 2000 * 450 * 500, 'F2': 2 * 2000 * 400 * 450, 'F3': 2 * 2000 * 300 * 400, 'F4': 2 * 2000 * 400 * 300, 'F5': 2 * 2000 * 450 * 400,
     'F6': 2 * 2000 * 500 * 450}
fl = dict(F)
for k, v in F.items():
    fl[k.replace('F', 'T')] = v * 0.9
fl.update({'B6': 2 * 2000 * 500 * 450 * 2, 'B5': 2 * 2000 * 450 * 400 * 2, 'B4': 2 * 2000 * 400 * 300 * 2, 'B3': 2 * 2000 * 300 * 400 * 2,
           'B2': 2 * 2000 * 400 * 450 * 2, 'B1': 2 * 2000 * 450 * 500})
it0 = vals[:39]
tot = sum(it0)
lines, g, b, a, gf = [], 0.0, 0.0, 0.0, 0.0
for n, v in zip(names, it0):
    base = n.split(' ')[0].rstrip('b')
    tf = f"{fl[base] * 64 / (v * 1e-6) / 1e12:.0f}" if base in fl else ""
    if base in fl:
        g += v; gf += fl[base] * 64
    if base.startswith('BN'):
        b += v
    if base.startswith('ADAM'):
        a += v
    lines.append(f"| {n} | {v:.1f} | {100 * v / tot:.1f}% | {tf} |")
inkernel = ""
if phases_log:
    txt = open(phases_log).read()
    inkernel = "\nIn-kernel phase times of the same build (`tests/dev_phases.py`, `%globaltimer` after every grid barrier of the persistent kernel, us):\n\n```\n" + \
        "\n".join(l for l in txt.splitlines() if l.startswith(('launch ms', '{', 'sum us'))) + "\n```\n"
md = f"""# Round 1, engine v4 (CTA pairs) - tcgen05 3xTF32 path, per-phase launch list (ncu gpu__time_duration.sum, --clock-control none)

Command: `BAE_PHASE_LAUNCH=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python tools/profile_step.py --chains 64 --rounds 2 --steps-per-round 5 --profile-round 1`
(Madelon shape, 64 chains, the 2nd PT round: 5 iterations + 1 exchange sweep; every phase of the persistent step program launched as its
own `bae_phase_kernel_tc` so that ncu can list them; production = ONE cooperative launch per round). Raw list: `r1_v4_tc_phase_launches.csv`.
Per-launch times are cold-cache and serialised; compare SHARES. Sum per iteration: {', '.join(f'{sum(vals[i * 39:(i + 1) * 39]) / 1000:.2f}' for i in range(5))} ms;
exchange sweep (decide / gather / scatter): {vals[195]:.0f} / {vals[196]:.0f} / {vals[197]:.0f} us.

| phase | iteration 0 (us) | share | algorithmic TFLOP/s (64 chains) |
|---|---:|---:|---:|
""" + "\n".join(lines) + f"""

GEMM phases: {100 * g / tot:.1f}% of the iteration ({gf / (g * 1e-6) / 1e12:.0f} TFLOP/s algorithmic over the GEMM phases; the sustained-clock ceiling of a
3-way TF32 split is 1381.9 / 6 = 230 TFLOP/s); BatchNorm {100 * b / tot:.1f}%; Adam + noise {100 * a / tot:.1f}%.
{inkernel}
The shares agree for the GEMM phases (same time inside the persistent kernel as stand-alone). The BatchNorm and Adam phases are slower inside the
persistent kernel than as stand-alone launches: its register allocation spills in these phases and the 227 KB shared-memory carve-out leaves no L1
behind local memory (the Adam inputs are already streamed through shared memory by the bulk-copy engine; without that they took 1.8 / 1.4 ms).
"""
open('profiles/r1_v4_tc_phase_launches.md', 'w').write(md)

r = list(csv.reader(open(raw_csv)))
d = {c: (u, v) for c, u, v in zip(r[0], r[1], r[2])}
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'sm__cycles_elapsed.max']
tab = "\n".join(f"| {k} | {d[k][0]} | {d[k][1]} |" for k in keys if k in d)


def to_bytes(k):
    u, v = d[k]
    return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]


rd, wr, ms = to_bytes('dram__bytes_read.sum'), to_bytes('dram__bytes_write.sum'), float(d['gpu__time_duration.sum'][1])
md2 = f"""# Round 1, engine v4 (CTA pairs) - `bae_program_kernel_tc` (persistent step kernel, tcgen05 3xTF32 path), one `ncu --set full` capture

Command: `ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bae_program_kernel_tc -c 1 python tools/profile_step.py --chains 8 --rounds 2 --steps-per-round 1 --profile-round 1`
(Madelon shape, 8 chains, ONE iteration = one cooperative launch: 148 CTAs = 74 CTA pairs (cluster dimension 2) x 512 threads, 1 CTA/SM; the report itself stays in `gpurun_out/`.)

Kernel duration {ms:.3f} ms under ncu. Algorithmic FLOPs of this launch: 8 chains x 27.18 GFLOP = 217 GFLOP -> {217 / ms:.0f} TFLOP/s at 8 chains
(production runs 64 chains per GPU: 121 TFLOP/s, `bench.py`).
DRAM traffic per launch: read {rd / 1e6:.0f} MB + write {wr / 1e6:.0f} MB = {(rd + wr) / 8 / 1e6:.0f} MB per chain-iteration. Algorithmic weight-state traffic is
22 x 4 B x 1.05 M = 93 MB per chain-iteration (SURVEY 8d); the rest is activations and deltas (41 MB per chain, written once and re-read by the backward
pass and the weight-gradient GEMMs) plus the four K-split copies of the gradient (4 x 4.2 MB written by the dW GEMMs, read by each Adam phase).
SASS evidence (`cuobjdump -sass libbae_b200.so`): UTCHMMA (tcgen05.mma kind::tf32), UTMALDG (TMA), UBLKCP (cp.async.bulk, Adam staging), LDTM (tcgen05.ld),
UTCBAR (tcgen05.commit), USETMAXREG (setmaxnreg), SYNCS (mbarrier).

| metric | unit | value |
|---|---|---|
{tab}

Reading: tensor pipe active {d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'][1][:5]}% of the cycles over the WHOLE step (GEMM phases are ~70% of it, and each 3xTF32 K block needs 6 MMAs);
DRAM {d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][1][:4]}% of peak - the step is bound by the operand pipeline of the tile loop (TMA -> hi/lo conversion -> MMA, see
`r1_tc_microbench.md`) and, in the BatchNorm / Adam phases, by the 256 threads per SM that run them. Local-memory instructions
({d['sass__inst_executed_local_loads'][1][:7]} loads) are register spills of the persistent kernel; they go to L2 (no L1 with a 227 KB carve-out) - the next thing to remove.
"""
open('profiles/r1_v4_program_kernel_ncu_full.md', 'w').write(md2)
json.dump({"shape": "madelon", "gemm_path": 2, "chains": 8, "iterations": 1, "dram_bytes_read": rd, "dram_bytes_write": wr, "kernel_ms_under_ncu": ms,
           "source": "ncu --set full, bae_program_kernel_tc, 8 chains x 1 iteration (profiles/r1_v4_program_kernel_ncu_full.md)"},
          open('profiles/r1_program_kernel_traffic.json', 'w'), indent=1)
print("profiles written")
