"""Mode-kernel *descriptors*.

Same classes, constructor arguments, attributes and error behaviour as the reference's
Modes/kernels.py (ModeKernel KERN:7-68, LinearMode :71-141, QuadraticMode :144-228,
GaussianMode :231-343), but they hold no JAX closures: `setup()` only resolves the scales into
the chd_kernel_desc the CUDA library consumes (include/chd_b200.h).  The Gram matrices and the
per-cluster influence terms are computed by csrc/gram.cu and csrc/activation.cu.
"""
from .. import _native


class ModeKernel:
    def __init__(self) -> None:
        self._scale = 1.0

    @property
    def is_interpolatory(self):
        res = self._is_interpolatory
        assert isinstance(res, bool)
        return res

    @property
    def memory_efficient_required(self):
        return self._memory_efficient_required

    @property
    def scale(self):
        return self._scale

    def __repr__(self) -> str:
        return self.name

    def __mul__(self, other):
        # KERN:61-65: only a float multiplier is accepted (anything else returns None)
        if isinstance(other, float):
            assert other > 0
            self._scale = other
            return self

    def __rmul__(self, other):
        return self.__mul__(other)

    # --- extension hook (KERN:7-68): a user subclass provides `name`, `_scale`, `_is_interpolatory`,
    # `_memory_efficient_required`, `setup(X, scales)` and `kernel(X, Y, which_dim)` (NumPy in, NumPy out);
    # it is served by the generic path: host-provided Gram -> chd_regress, activations through
    # `individual_influence` (default: whole kernel minus the kernel without the cluster, KERN:39-56).
    def individual_influence(self, X, Y, which_dim, which_dim_only):
        whole_k = self(X, Y, which_dim)
        rest = which_dim * (1 - which_dim_only)
        rest_k = self(X, Y, rest)
        return whole_k - rest_k

    def __call__(self, X, Y, which_dim):
        return self.kernel(X, Y, which_dim)

    def setup(self, X, scales):
        raise NotImplementedError

    @property
    def desc(self):
        """chd_kernel_desc for the CUDA library (built-in modes, valid after setup()); None for user kernels."""
        return getattr(self, "_desc", None)


class LinearMode(ModeKernel):
    def __init__(self, memory_efficient_required=False) -> None:
        super().__init__()
        self.hyperparameters = {}
        self._is_interpolatory = False
        self._memory_efficient_required = memory_efficient_required
        self.name = "linear"
        self._scale = 2.0   # KERN:97

    def setup(self, X, scales):
        assert self.scale == scales[self.name]
        self._desc = _native.make_desc(_native.KERNEL_LINEAR, self.scale)


class QuadraticMode(ModeKernel):
    def __init__(self, memory_efficient_required=False) -> None:
        super().__init__()
        self._is_interpolatory = False
        self.name = "quadratic"
        self._memory_efficient_required = memory_efficient_required

    def setup(self, X, scales):
        assert self.scale == scales[self.name]
        if "linear" not in scales:          # KERN:197-200 (mutates the shared dict)
            scales["linear"] = 2.0
        self.alpha = 0.5 * scales["linear"] / self.scale
        self._desc = _native.make_desc(_native.KERNEL_QUADRATIC, self.scale, alpha=self.alpha)


class GaussianMode(ModeKernel):
    def __init__(self, l, memory_efficient_required=True) -> None:
        super().__init__()
        self._l = l
        self._is_interpolatory = True
        self._memory_efficient_required = memory_efficient_required
        self.name = "gaussian"
        self.quadratic_part = QuadraticMode()

    def setup(self, X, scales):
        assert self.scale == scales[self.name]
        self.quadratic_part._scale = scales["quadratic"]   # KeyError without a QuadraticMode, KERN:302
        self.quadratic_part.setup(X, scales)
        self._desc = _native.make_desc(_native.KERNEL_GAUSSIAN, self.scale, alpha=self.quadratic_part.alpha,
                                       quad_scale=self.quadratic_part.scale, l=float(self._l))

    @property
    def l(self):
        return self._l

    @l.setter
    def l(self, value):
        raise AttributeError("Length scale cannot be changed after construction")

    def __getstate__(self):
        return self.__dict__.copy()
