"""Seeded synthetic structural-equation data of the shapes BASELINE.json names:
roots W ~ N(0,1) of shape (n, N/2); every non-root x_i = f_i(1-3 random roots) + 0.1 N(0,1) with f_i cycled
from {sum, product, square+1, w sin(w'), cube} (the families of the reference's algebraic examples);
columns shuffled with the same generator.  Used by bench.py and the parity tests (same array is fed
to the CUDA path and to the oracle)."""
import numpy as np


def make_data(n, N, seed=0, noise=0.1):
    rng = np.random.default_rng(seed)
    n_roots = max(1, N // 2)
    W = rng.standard_normal((n, n_roots))
    cols = [W[:, j] for j in range(n_roots)]
    for i in range(N - n_roots):
        k = int(rng.integers(1, 4))
        pa = rng.choice(n_roots, size=min(k, n_roots), replace=False)
        a = W[:, pa[0]]
        b = W[:, pa[1 % len(pa)]]
        fam = i % 5
        if fam == 0:
            x = W[:, pa].sum(axis=1)
        elif fam == 1:
            x = a * b
        elif fam == 2:
            x = a ** 2 + 1
        elif fam == 3:
            x = a * np.sin(b)
        else:
            x = a ** 3
        cols.append(x + noise * rng.standard_normal(n))
    X = np.stack(cols, axis=1)
    perm = rng.permutation(N)
    X = X[:, perm]
    names = [f"v{j}" for j in range(N)]
    return X, names
