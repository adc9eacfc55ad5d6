"""bipartitesbm-mcmc_b200 -- B200-native Metropolis-Hastings sweep of the bipartite SBM.

Only the hot path of junipertcy/bipartiteSBM-MCMC lives here: csrc/ (CUDA kernels + the C ABI
of libbisbm.so, include/bisbm.h) and host.py (the Python mirror of the reference interface).
The directory name has a hyphen; import it with importlib.import_module("bipartitesbm-mcmc_b200").
"""
from . import build, dist, host  # noqa: F401
from .host import (ChainPool, Graph, BisbmError, blockmodel_t, metropolis_hasting, mt19937,  # noqa: F401
                   edge_to_adj, load_edge_list, load_memberships, memberships_from_block_sizes, load_library,
                   exponential_schedule, linear_schedule, logarithmic_schedule, constant_schedule, abrupt_cool_schedule)
