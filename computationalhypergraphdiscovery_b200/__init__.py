"""computationalhypergraphdiscovery_b200 -- B200-native Computational Hypergraph Discovery hot path.

Drop-in for the reference package `ComputationalHypergraphDiscovery` (its __init__.py:26-29 exports
GraphDiscovery, Modes, decision, util): same classes and call signatures, FP64 throughout, with the
per-node Gaussian-process ancestor discovery running in hand-written sm_100a CUDA kernels behind the
C ABI of include/chd_b200.h (libchd_b200.so).  There is no CPU fallback.

    import computationalhypergraphdiscovery_b200 as CHD
    gd = CHD.GraphDiscovery(X, names)      # or CHD.GraphDiscovery.from_dataframe(df)
    gd.fit()
    gd.G                                   # networkx.DiGraph with the discovered hyperedges
"""
from ._GraphDiscoveryMain import GraphDiscovery
from . import Modes
from . import decision
from . import util
from . import random

__version__ = "0.1.0"
__all__ = ["GraphDiscovery", "Modes", "decision", "util", "random"]
