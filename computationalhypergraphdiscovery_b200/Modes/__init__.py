"""Mode kernels and the cluster container (mirror of the reference's
`ComputationalHypergraphDiscovery.Modes`, Modes/__init__.py:13-14)."""
from .kernels import ModeKernel, LinearMode, QuadraticMode, GaussianMode
from .container import ModeContainer
