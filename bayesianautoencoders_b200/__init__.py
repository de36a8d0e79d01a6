"""bayesianautoencoders_b200 -- B200-native parallel-tempering Langevin sampler over autoencoder weights.

Host-side mirror of the reference's interface for this path (`Model`, `ptReplica`,
`ParallelTempering`, MAIN:145-1090) over a C-ABI CUDA library (include/bae_b200.h).
"""
from .sampler import DeviceSampler, Hyper, default_theta_init, param_offsets, param_shapes, SORTED_KEYS  # noqa: F401
from ._native import BaeError  # noqa: F401
from .pt import Model, ParallelTempering, ptReplica  # noqa: F401
from .predict import PosteriorPredictive  # noqa: F401
from .results import gelman_rubin, gelman_rubin_reference, write_replica_files  # noqa: F401

__version__ = "0.1.0"
