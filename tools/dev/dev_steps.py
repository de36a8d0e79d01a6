import sys, numpy as np
sys.path.insert(0, '.')
from oracle import bae_oracle as bo
from tests.golden_util import GoldenCase
from tests.test_gpu_parity import _run_both
np.set_printoptions(linewidth=200)
name = sys.argv[1] if len(sys.argv) > 1 else "swiss_1chain"
swaps = len(sys.argv) > 2
g = GoldenCase(name)
dev, records, swap_log, counts, reps, final, init_states = _run_both(
    g.dims4, g.train, g.test, list(g.temps), g.S, g.swap_interval,
    [g.theta_init(j) for j in range(g.n)], [g.rnd(j)["w0_randn"] for j in range(g.n)], n_rungs=(g.n if swaps else 1), do_swaps=swaps)
print('init', init_states[0])
print('swap_log', swap_log, counts)
for j in range(g.n):
    for i, rec in enumerate(records[j]):
        d = dev[j][i]
        print(j, i, 'L' if rec.langevin else 'R', int(d['langevin']), 'acc', int(rec.accepted), int(d['accepted']),
              'likp %.5f %.5f' % (rec.likelihood_proposal, d['likelihood_proposal']),
              'lik %.5f %.5f' % (rec.likelihood, d['likelihood']),
              'prior %.4f %.4f' % (rec.prior_prop, d['prior_proposal']),
              'dprop %.3f %.3f' % (rec.diff_prop, d['diff_prop']),
              'sum %.3f %.3f' % (rec.sum_value, d['sum_value']))
