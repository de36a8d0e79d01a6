"""Accuracy experiment (GPU): teacher-forced Madelon case, gradient / likelihood errors of one library build + env."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.parity_util import run_teacher_forced, synth_data
from tests.test_gpu_step_parity import _inputs, MADELON
dims4, ntr, nte = MADELON
train, test = synth_data(dims4, ntr, nte)
th0, w0 = _inputs(dims4, 2, 17)
gp = int(os.environ.get("EXP_GEMM_PATH", "2"))
rep, _ = run_teacher_forced(dims4, train, test, [1.0, 2.0], 6, 5, th0, w0, n_rungs=2, seed=6, gemm_path=gp)
e = rep["errors"]
print(json.dumps({"tag": os.environ.get("EXP_TAG", ""), "grad": ["%.2e" % x for x in e["grad"]], "lik": ["%.2e" % x for x in e["lik_prop"]],
                  "mse_test": ["%.2e" % x for x in e["mse_test"]], "sum_value": ["%.2e" % x for x in e["sum_value"]]}))
