"""The BASELINE.json configurations that fit one `fit()` call, as synthetic data of the named shapes (SURVEY.md 8d;
the reference's own datasets are git-LFS pointers).  Used by bench.py (`fit_wall_s`), tools/fit_configs.py and the
multi-GPU equality check.  Each builder returns (GraphDiscovery kwargs, fit kwargs, description); the oracle-side
equivalents (`oracle_kwargs`) are plain tuples so that only tests / the CPU-baseline leg ever import the oracle."""
import networkx as nx
import numpy as np

from .synthetic import make_data


def _possible_edges(names, seed, keep=0.8):
    rng = np.random.default_rng(seed)
    pe = nx.DiGraph()
    pe.add_nodes_from(names)
    for a in names:
        for b in names:
            if a != b and rng.random() < keep:
                pe.add_edge(a, b)
    return pe


def c1():
    """configs[0]: Sachs 11 proteins, 500-sample subsample, LinearMode + QuadraticMode."""
    from . import Modes
    X, names = make_data(500, 11, seed=1)
    return (dict(X=X, names=names, kernels=[Modes.LinearMode(), Modes.QuadraticMode()]), {},
            "C1 Sachs-500 shape: n=500, N=11, Linear+Quadratic",
            dict(kernels=[("linear", 2.0, None), ("quadratic", 1.0, None)]))


def c2(n=7466):
    """configs[1]: Sachs full dataset shape, three kernels scaled by 0.1, 4 clusters, possible_edges."""
    from . import Modes
    X, names = make_data(n, 11, seed=2)
    clusters = [names[0:3], names[3:6], names[6:8], names[8:11]]
    kernels = [0.1 * Modes.LinearMode(), 0.1 * Modes.QuadraticMode(), 0.1 * Modes.GaussianMode(l=1)]
    return (dict(X=X, names=names, kernels=kernels, clusters=clusters, possible_edges=_possible_edges(names, 3)), {},
            f"C2 Sachs-full shape: n={n}, N=11, 4 clusters, possible_edges, 0.1*(Linear,Quadratic,Gaussian)", None)


def c3(n=600, ns=1122, b=32, targets=32):
    """configs[2]: Chemistry-Big shape (experiment.py:39-63): ns state columns, b derivative columns that are sparse
    quadratic forms of 8 states, bipartite possible_edges, gamma_min=2e-6, ThresholdModeChooser(0.025)."""
    from . import Modes, decision
    rng = np.random.default_rng(5)
    S = rng.random((n, ns))
    D = np.empty((n, b))
    for j in range(b):
        idx = rng.choice(ns, size=8, replace=False)
        D[:, j] = (S[:, idx[0]] * S[:, idx[1]] - S[:, idx[2]] * S[:, idx[3]] + S[:, idx[4]] * S[:, idx[5]]
                   - 0.5 * S[:, idx[6]] + 0.01 * rng.standard_normal(n))
    X = np.concatenate([S, D], axis=1)
    names = [f"s{i}" for i in range(ns)] + [f"partial_{j}" for j in range(b)]
    pe = nx.DiGraph()
    pe.add_edges_from([e for f in names[:ns] for g in names[ns:] for e in ((f, g), (g, f))])
    return (dict(X=X, names=names, kernels=[Modes.QuadraticMode()], possible_edges=pe, gamma_min=2e-6),
            dict(targets=names[ns:ns + targets], mode_chooser=decision.ThresholdModeChooser(threshold=0.025)),
            f"C3 Chemistry-Big shape: n={n}, N={ns}+{b}, bipartite possible_edges, Quadratic, {targets} targets, "
            f"gamma_min=2e-6", None)


BUILDERS = {"c1": c1, "c2": c2, "c3": c3}
