"""Import shim: `import ComputationalHypergraphDiscovery as CHD` (the reference's package name, its
__init__.py:26-29) resolves to the B200-native implementation, so the reference's notebooks and scripts run
unmodified.  Everything lives in `computationalhypergraphdiscovery_b200`; this module only aliases it
(including the sub-modules `Modes`, `Modes.kernels`, `Modes.container`, `decision`, `util`, `random`)."""
import importlib
import sys

import computationalhypergraphdiscovery_b200 as _impl
from computationalhypergraphdiscovery_b200 import GraphDiscovery, Modes, decision, util, random  # noqa: F401

__version__ = _impl.__version__
__all__ = list(_impl.__all__)

for _sub in ("Modes", "Modes.kernels", "Modes.container", "decision", "util", "random", "_GraphDiscoveryMain"):
    sys.modules[__name__ + "." + _sub] = importlib.import_module(_impl.__name__ + "." + _sub)
