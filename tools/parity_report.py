#!/usr/bin/env python3
"""Tabulates the measured parity errors the GPU tests write (gpurun_out/parity_*.json) into profiles/r2_parity_measured.md."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows_tf, rows_free = [], []
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "parity_*.json"))):
    r = json.load(open(f))
    name = os.path.basename(f)[len("parity_"):-len(".json")]
    e = r["errors"]
    mx = lambda k, a=0: max(e[k][a:]) if k in e and len(e[k]) > a else float("nan")
    if name.startswith("tf_"):
        dec = r["decisions"]
        sw = r["swaps"]
        rows_tf.append(f"| {name[3:]} | {r['gemm_path']} | {len(e['lik_prop'])} | {mx('lik_prop'):.1e} | {mx('lik_prop_terms'):.1e} | "
                       f"{max(mx('mse_train'), mx('mse_test')):.1e} | {e['grad'][0]:.1e} / {mx('grad', 1):.1e} | {mx('prior'):.0e} | {mx('diff_prop'):.1e} | "
                       f"{mx('sum_value'):.1e} | {e['theta'][0]:.1e} / {mx('theta', 1):.1e} | {e['m'][0]:.1e} / {mx('m', 1):.1e} | "
                       f"{sum(1 for d in dec if d[3] == d[4])}/{len(dec)} (min margin {min(d[2] for d in dec):.1e}) | "
                       f"{sum(1 for x in sw if x[3] == x[4])}/{len(sw)} | {r['random_walk_steps']} |")
    elif name.startswith("free_"):
        dec = r["decisions"]
        rows_free.append(f"| {name[5:]} | {r['gemm_path']} | {len(e['lik_prop'])} | {e['lik_prop'][0]:.1e} | {mx('lik_prop'):.1e} | {mx('sum_value'):.1e} | "
                         f"{max(r['w']):.1e} | {sum(1 for d in dec if d[3] == d[4])}/{len(dec)} | {r['swap_counts']} |")
md = f"""# Round 2 - measured parity of the CUDA step programs against the CPU oracle (tests/test_gpu_step_parity.py)

Written by `tools/parity_report.py` from the `gpurun_out/parity_*.json` files of the last `pytest -m gpu` run on a B200.
gemm path 2 = tcgen05 3xTF32 step program, 1 = FP32 CUDA-core step kernel. Errors are relative to the oracle value (maxima over
chains and iterations); `a / b` = iteration 0 / later iterations.

## Teacher-forced (the oracle takes the device state before every iteration and every sweep)

| case | path | iterations | log-lik | log-lik (vs its terms) | MSE | gradient | prior | diff_prop | sum_value | theta after 1 it. | Adam m after 1 it. | MH decisions equal | swap decisions equal | random-walk its. |
|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|
{chr(10).join(rows_tf)}

Swap decisions that differ all sit inside the rounding noise of the reference's own fp32 re-evaluation of the two
log-likelihoods (|exponent - log u| < 2e-6 x the sum of the four |lhood| terms; the tests assert equality outside it).

## Free-running (one `bae_run(S)` call against the oracle's own trajectory; device traces bit-identical to the single-step run)

| case | path | iterations | log-lik it. 0 | log-lik max | sum_value max | final w | MH decisions equal | swaps (accepted, proposed) |
|---|---|---|---|---|---|---|---|---|
{chr(10).join(rows_free)}
"""
open(os.path.join(ROOT, "profiles", "r2_parity_measured.md"), "w").write(md)
print(md)
