"""CPU oracle for the CHD hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy FP64 restatement of the reference's JAX
implementation (TheoBourdais/ComputationalHypergraphDiscovery, v2.1.0) of the
per-node ancestor-discovery hot path.  It exists to *check* the CUDA product
(`computationalhypergraphdiscovery_b200`), never to serve it:

  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
    `--impl reference` legs may import it;
  * the product package never imports it and has no CPU fallback.

Parity status: PINNED.  The reference cannot be imported here (jax/jaxlib 0.4.28
are not installed, no network), so the restatement is pinned against the
reference's own stored notebook outputs (the only result-pinning artefacts the
reference has; it ships no tests): 81 stage-1 noise values at 16 digits + gamma
(3 digits) + Z bounds (2 decimals) + final ancestor lists + post-prune
gamma/noise of `examples/Algebraic-equations/algebraic_equations_graph_discovery.ipynb`
and the 18 stage-1 values of `predictions-example1.ipynb`.  The harvested
vectors live in `tests/golden/` together with the script that made them
(`tests/golden/make_golden.py`); `tests/test_oracle_golden.py` enforces them.

Third-party arithmetic restated here (absent from /root/reference): jax /
jaxlib 0.4.28 (pin: requirements.txt:1-2): threefry2x32 PRNG, random.split /
random.normal, jax.scipy.optimize.minimize(method="BFGS") + its strong-Wolfe
line search, nanmedian / argmin / sort semantics.

Every function cites the reference file:line it follows.  Abbreviations:
  MAIN = src/ComputationalHypergraphDiscovery/_GraphDiscoveryMain.py
  REG  = src/ComputationalHypergraphDiscovery/interpolatory.py (== non_interpolatory.py)
  HELP = src/ComputationalHypergraphDiscovery/helper_functions.py
  KERN = src/ComputationalHypergraphDiscovery/Modes/kernels.py
  CONT = src/ComputationalHypergraphDiscovery/Modes/container.py
  DEC  = src/ComputationalHypergraphDiscovery/decision.py
"""
from . import jaxprng, bfgs, kernels, regression, chd  # noqa: F401
