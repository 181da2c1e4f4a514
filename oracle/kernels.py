"""Oracle (test infrastructure): the three CHD mode kernels restated in NumPy.

Follows KERN (src/ComputationalHypergraphDiscovery/Modes/kernels.py):
  linear    K = 1 + s_l (X*w) Y^T                                   KERN:99-102
  quadratic K = s_q (alpha + (X*w) Y^T)^2 + 1 - alpha^2 s_q          KERN:169-173
            alpha = 0.5 * scales["linear"] / s_q                    KERN:197-201
  gaussian  K = s_g prod_d(1 + w_d E_d) + K_quad - 1                 KERN:251-257
            E_d[i,j] = exp(-(x_id - y_jd)^2 / (2 l^2))               KERN:247-249
and the per-cluster "individual influence" matrices KERN:104-106, 175-185,
270-278.  The Gaussian product is evaluated dimension by dimension (never
materialising the reference's n x n x N `exps`, KERN:307-309); multiplying by
(1 + 0*E) = 1 for inactive dims is exact so skipping them is bit-identical.
"""
import numpy as np


class KernelSpec:
    """Plain description of one mode kernel as the reference configures it in
    `setup()` (KERN:108-123, 187-209, 293-317)."""

    def __init__(self, name, scale, alpha=None, quad_scale=None, l=None, fn=None):
        self.name = name          # "linear" | "quadratic" | "gaussian" | a user ModeKernel's name
        self.scale = float(scale)
        self.alpha = alpha        # quadratic shift (quadratic & gaussian)
        self.quad_scale = quad_scale  # scale of the gaussian's internal quadratic part
        self.l = l
        self.fn = fn              # user ModeKernel: kernel(X, Y, which_dim) (KERN:7-68)

    def __repr__(self):
        return self.name


def make_specs(kernels):
    """kernels: list of (name, scale, l) -> list[KernelSpec], following
    prepare_functions MAIN:235-237 and the `setup` methods (incl. the shared,
    mutated `scales` dict, KERN:197-200, and KeyError for gaussian w/o quadratic,
    KERN:302)."""
    scales = {name: scale for (name, scale, _l) in kernels}
    specs = []
    for (name, scale, l) in kernels:
        if callable(l):           # user ModeKernel (name, scale, kernel function)
            specs.append(KernelSpec(name, scale, fn=l))
        elif name == "linear":
            specs.append(KernelSpec("linear", scale))
        elif name == "quadratic":
            if "linear" not in scales:
                scales["linear"] = 2.0
            specs.append(KernelSpec("quadratic", scale, alpha=0.5 * scales["linear"] / scale))
        elif name == "gaussian":
            qs = scales["quadratic"]  # KeyError mirrors KERN:302
            if "linear" not in scales:
                scales["linear"] = 2.0
            specs.append(KernelSpec("gaussian", scale, alpha=0.5 * scales["linear"] / qs,
                                    quad_scale=qs, l=float(l)))
        else:
            raise ValueError(name)
    return specs


def linear_core(X, Y, w):
    return (X * w[None, :]) @ Y.T


def quad_from_core(L, scale, alpha):
    return scale * (alpha + L) ** 2 + (1 - alpha ** 2 * scale)


def make_exps(X, l):
    """GaussianMode.setup's cached tensor exps[j, i, d] (KERN:307-309; symmetric in i, j for X = Y)."""
    diff = X[:, None, :] - X[None, :, :]
    return np.exp(-(diff ** 2) / (2 * l ** 2))


def gaussian_product(X, Y, w, l, exps=None):
    """prod_d (1 + w_d * exp(-(x_d - y_d)^2 / (2 l^2)))  (KERN:253-255).  With the reference's cached
    `exps` the product runs over all N dims (mask by multiply) exactly as the reference does."""
    if exps is not None:
        return np.prod(1.0 + w[None, None, :] * exps, axis=2)
    P = np.ones((X.shape[0], Y.shape[0]))
    inv = 1.0 / (2 * l ** 2)
    for d in np.nonzero(w)[0]:
        diff = X[:, d][:, None] - Y[:, d][None, :]
        P *= 1.0 + w[d] * np.exp(-(diff ** 2) * inv)
    return P


def gram(spec, X, Y, w, exps=None):
    """kernel(X, Y, which_dim)  (ModeKernel.__call__ KERN:58-59).  `exps`: optional cache (X is Y)."""
    w = np.asarray(w, dtype=np.float64)
    if spec.fn is not None:
        return np.asarray(spec.fn(X, Y, w), dtype=np.float64)
    L = linear_core(X, Y, w)
    if spec.name == "linear":
        return 1 + spec.scale * L
    if spec.name == "quadratic":
        return quad_from_core(L, spec.scale, spec.alpha)
    if spec.name == "gaussian":
        return (spec.scale * gaussian_product(X, Y, w, spec.l, exps)
                + quad_from_core(L, spec.quad_scale, spec.alpha) - 1)
    raise ValueError(spec.name)


def _quad_influence(X, Y, w, w_only, scale, alpha):
    lin_only = linear_core(X, Y, w_only)
    rest = w * (1 - w_only)
    lin_rest = linear_core(X, Y, rest)
    return scale * (alpha + lin_only) ** 2 + 2 * scale * lin_only * lin_rest - scale * alpha ** 2


def individual_influence(spec, X, Y, w, w_only, exps=None):
    """individual_influence(X, Y, which_dim, which_dim_only) KERN:125-141,
    211-228, 319-334."""
    w = np.asarray(w, dtype=np.float64)
    w_only = np.asarray(w_only, dtype=np.float64)
    if spec.fn is not None:
        own = getattr(spec.fn, "individual_influence", None)     # a user subclass may override the hook
        if own is not None:
            return np.asarray(own(X, Y, w, w_only), dtype=np.float64)
        return gram(spec, X, Y, w) - gram(spec, X, Y, w * (1 - w_only))   # ModeKernel default, KERN:39-56
    if spec.name == "linear":
        return gram(spec, X, Y, w_only) - 1                      # KERN:104-106
    if spec.name == "quadratic":
        return _quad_influence(X, Y, w, w_only, spec.scale, spec.alpha)  # KERN:175-185
    if spec.name == "gaussian":                                  # KERN:270-278
        rest = w * (1 - w_only)
        only_part = gaussian_product(X, Y, w_only, spec.l, exps) - 1
        rest_part = gaussian_product(X, Y, rest, spec.l, exps)
        return (_quad_influence(X, Y, w, w_only, spec.quad_scale, spec.alpha)
                + spec.scale * only_part * rest_part)
    raise ValueError(spec.name)
