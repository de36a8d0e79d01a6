"""Operand images (BAE_TC_PRECONV) against the converter path: traces and final weights must be BIT-identical; step time at 64 chains.
python tools/dev/exp_preconv.py [shape]"""
import os, subprocess, sys, numpy as np
sys.path.insert(0, '.')
shape = sys.argv[1] if len(sys.argv) > 1 else "madelon"
if len(sys.argv) > 2 and sys.argv[2] == "child":
    from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
    from bench import SHAPES, synth
    dims4, ntr, nte = SHAPES[shape]
    train, test = synth(dims4, ntr, nte)
    C = int(os.environ.get("EXP_CHAINS", "8"))
    s = DeviceSampler(dims4, train, test, C, 8, 100, np.tile(2.0 ** (np.arange(8) / 7), C // 8), swap_interval=int(os.environ.get("EXP_SI", "3")), gemm_path=2)
    rng = np.random.default_rng(0)
    for j in range(C): s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(s.w_size))
    s.run(7); s.synchronize()
    s.run(5); s.synchronize()
    ms = []
    for _ in range(3):
        s.run(5); s.synchronize(); ms.append(s.last_ms())
    out = {}
    for j in range(C):
        out[f"tr{j}"] = np.frombuffer(s.traces(j, 0, 27).tobytes(), dtype=np.uint8)
        st = s.get_state(j)
        for k in ("theta", "m", "v"): out[f"{k}{j}"] = st[k]
    np.savez(sys.argv[3], ms=np.array(ms), **out)
    print("preconv", os.environ.get("BAE_TC_PRECONV", "1"), "chains", C, "ms per 5 iterations", [round(x, 2) for x in ms])
    s.close()
    sys.exit(0)
res = []
for pc in ("0", "1"):
    env = dict(os.environ, BAE_TC_PRECONV=pc)
    f = f"/tmp/preconv_{pc}.npz"
    subprocess.run([sys.executable, __file__, shape, "child", f], env=env, check=True)
    res.append(np.load(f))
bad = 0
for k in res[0].files:
    if k == "ms": continue
    a, b = res[0][k], res[1][k]
    same = a.tobytes() == b.tobytes()
    if not same:
        bad += 1
        print("DIFF", k, np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))))
print("bit-identical" if bad == 0 else f"{bad} arrays differ")
