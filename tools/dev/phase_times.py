"""Per-phase milliseconds of back-to-back launches (BAE_PHASE_TIMES=1: an event after every phase), Madelon, 64 chains, averaged
over the iterations of a few bae_run(5) calls. python tools/dev/phase_times.py [lib.so ...]  (libraries under build/; none = in-tree)"""
import os, subprocess, sys
import numpy as np
NAMES = ['PRO', 'WAMAX', 'F1', 'F2', 'F3', 'BNF', 'F4', 'F5', 'F6', 'B6', 'B5', 'B4', 'BNB', 'B3', 'B2', 'B1', 'ADAM1', 'F1b', 'F2b', 'F3b',
         'BNFb', 'F4b', 'F5b', 'F6b', 'B6b', 'B5b', 'B4b', 'BNBb', 'B3b', 'B2b', 'B1b', 'ADAM2', 'T1', 'T2', 'T3', 'BNT', 'T4', 'T5', 'T6', 'EPI',
         'SWD', 'SWG', 'SWS']
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, '.')
    import ctypes as C
    from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
    from bayesianautoencoders_b200 import _native as N
    from bench import SHAPES, synth
    dims4, ntr, nte = SHAPES[os.environ.get("EXP_SHAPE", "madelon")]
    train, test = synth(dims4, ntr, nte)
    Cn = int(os.environ.get("EXP_CHAINS", "64"))
    s = DeviceSampler(dims4, train, test, Cn, 8, 100, np.tile(2.0 ** (np.arange(8) / 7), Cn // 8), swap_interval=5, gemm_path=2)
    rng = np.random.default_rng(0)
    for j in range(Cn): s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(s.w_size))
    ms = np.zeros(64); n = np.zeros(64, np.int32)
    s.run(5); s.synchronize()
    N.check(s.lib.bae_debug_phase_times(s._h, ms.ctypes.data, n.ctypes.data))
    tot = []
    for _ in range(3): s.run(5); s.synchronize(); tot.append(s.last_ms())
    N.check(s.lib.bae_debug_phase_times(s._h, ms.ctypes.data, n.ctypes.data))
    np.save(sys.argv[2], np.concatenate([ms / np.maximum(n, 1), [np.mean(tot)]]))
    s.close(); sys.exit(0)
libs = sys.argv[1:] or [None]   # a library under build/, or VAR=value (an environment switch of the in-tree library)
cols = []
for l in libs:
    env = dict(os.environ, BAE_PHASE_TIMES="1")
    for part in (l.split(":") if l else []):   # lib.so, VAR=value, or lib.so:VAR=value
        if "=" in part: env[part.split("=")[0]] = part.split("=")[1]
        else: env["BAE_B200_LIB"] = os.path.join(os.getcwd(), "build", part)
    subprocess.run([sys.executable, __file__, "child", "/tmp/pt.npy"], env=env, check=True, stderr=subprocess.DEVNULL)
    cols.append(np.load("/tmp/pt.npy"))
print("%-6s" % "phase" + "".join("%14s" % (l or "in-tree")[:13] for l in libs))
for i, nme in enumerate(NAMES):
    print("%-6s" % nme + "".join("%14.1f" % (c[i] * 1000) for c in cols))
g = [i for i, nme in enumerate(NAMES) if nme[0] in "FBT" and not nme.startswith("BN")]
print("%-6s" % "gemm" + "".join("%14.1f" % (sum(c[i] for i in g) * 1000) for c in cols))
print("%-6s" % "iter" + "".join("%14.1f" % (sum(c[:40]) * 1000) for c in cols))
print("%-6s" % "run5" + "".join("%14.2f" % c[64] for c in cols) + "  ms per bae_run(5) (direct launches)")
