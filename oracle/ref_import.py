"""Import the UNMODIFIED reference script as a module (TEST INFRASTRUCTURE ONLY).

This file is checker code. It exists only in the build container, where the
read-only reference tree is mounted at /root/reference; it is used by
`oracle/make_golden.py` to freeze golden vectors under `tests/golden/` and by
the `-m "not gpu"` tests that pin `oracle/bae_oracle.py` against the reference
when the tree is present. Nothing in the product package may import it.

The reference (`Bayesian/bayes_autoenc_langevincheck.py`, "MAIN") is a script,
not a package. Importing it needs four stand-ins:

1. `matplotlib` / `mpl_toolkits` are not installed   -> MagicMock modules (MAIN:26,30).
2. The module body downloads Madelon at import time   -> `urlopen` returns synthetic text (MAIN:90-94).
3. It rewrites `Experiment No.txt` in the cwd         -> import from a scratch cwd (MAIN:46-51).
4. Shapes are module globals read at call time        -> set with `set_shape()` after import (MAIN:75-102).
"""
from __future__ import annotations

import importlib.util
import io
import os
import sys
import tempfile
from unittest import mock

REF_MAIN = "/root/reference/Bayesian/bayes_autoenc_langevincheck.py"

SHAPES = {
    # name: (in_shape, in_one, in_two, enc_shape)   MAIN:76-79, 84-87, 97-100
    "swiss": (3, 10, 5, 2),
    "coil": (85, 70, 60, 50),
    "madelon": (500, 450, 400, 300),
}

_cached = None


def available() -> bool:
    return os.path.exists(REF_MAIN)


def load_reference():
    """Return the reference module object (imported once per process)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference tree not mounted at /root/reference")
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "mpl_toolkits.mplot3d.axes3d"):
        sys.modules.setdefault(name, mock.MagicMock())
    import urllib.request

    def fake_urlopen(url, *a, **k):  # 4 rows of Madelon-format text; never used afterwards
        return io.StringIO("\n".join(" ".join(["1"] * 500) for _ in range(4)))

    scratch = tempfile.mkdtemp(prefix="bae_ref_")
    with open(os.path.join(scratch, "Experiment No.txt"), "w") as f:
        f.write("1")
    old_cwd = os.getcwd()
    os.chdir(scratch)
    try:
        with mock.patch.object(urllib.request, "urlopen", fake_urlopen):
            spec = importlib.util.spec_from_file_location("bae_reference_main", REF_MAIN)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
    finally:
        os.chdir(old_cwd)
    mod.print = lambda *a, **k: None  # silence per-accept prints (MAIN:539)
    mod._scratch = scratch
    _cached = mod
    return mod


def set_shape(mod, dims4, lrate=0.01, step_size=0.005):
    """Select the autoencoder widths (read by `Model.__init__`, MAIN:151-178)."""
    mod.in_shape, mod.in_one, mod.in_two, mod.enc_shape = [int(x) for x in dims4]
    mod.lrate = lrate
    mod.step_size = step_size
    mod.use_dataset = 3  # keeps the post-loop classification tail on its cheapest branch (MAIN:709-818)
