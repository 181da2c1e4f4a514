"""Kernel / mode choosers: same classes, call signatures and decisions as the reference's
decision.py (KernelChooser DEC:8-31, MinNoiseKernelChooser :34-63, ThresholdKernelChooser :66-100,
ModeChooser :103-139, MaxIncrementModeChooser :142-155, ThresholdModeChooser :158-192,
EarlyStopping* :195-236).  They run per node on host NumPy arrays (the reference jit/vmaps them,
MAIN:305, which changes nothing about the result).  jnp.argmin/argmax semantics are kept:
first occurrence, NaN wins."""
from typing import Any

import numpy as np


def _argmin(a):
    a = np.asarray(a)
    nan = np.isnan(a)
    return int(np.argmax(nan)) if nan.any() else int(np.argmin(a))


def _argmax(a):
    a = np.asarray(a)
    nan = np.isnan(a)
    return int(np.argmax(nan)) if nan.any() else int(np.argmax(a))


class KernelChooser:
    def __init__(self):
        pass

    def is_a_kernel_chooser(self):
        return True


class MinNoiseKernelChooser(KernelChooser):
    """Kernel with the least noise among those whose noise is below the Z-test lower bound."""

    def __init__(self):
        pass

    def __call__(self, kernel_perfs):
        ybs, noises, Z_lows, Z_highs, gammas, keys = kernel_perfs
        valid_indices = noises < Z_lows
        is_valid = bool(np.any(valid_indices))
        selected_index = _argmin(np.where(valid_indices, noises, 2))
        return (
            selected_index * is_valid - 1 * (1 - is_valid),
            ybs[selected_index],
            noises[selected_index],
            Z_lows[selected_index],
            Z_highs[selected_index],
            gammas[selected_index],
            keys[selected_index],
        )


class ThresholdKernelChooser(KernelChooser):
    """First kernel whose noise is below the threshold."""

    def __init__(self, threshold):
        self.threshold = threshold

    def __call__(self, kernel_perfs):
        ybs, noises, Z_lows, Z_highs, gammas, keys = kernel_perfs
        valid_indices = noises < self.threshold
        indices = np.where(valid_indices, np.arange(noises.shape[0]), noises.shape[0])
        selected_index = int(np.min(indices))
        is_valid = selected_index < noises.shape[0]
        gather = min(selected_index, noises.shape[0] - 1)   # JAX clamps out-of-bounds gathers
        return (
            selected_index * is_valid - 1 * (1 - is_valid),
            ybs[gather],
            noises[gather],
            Z_lows[gather],
            Z_highs[gather],
            gammas[gather],
            keys[gather],
        )


class ModeChooser:
    def __init__(self):
        pass

    def is_a_mode_chooser(self):
        return True

    def __call__(self, *args: Any, **kwds: Any) -> Any:
        Exception("Not implemented")


class MaxIncrementModeChooser(ModeChooser):
    """Chain index with the largest noise increment."""

    def __init__(self):
        pass

    def __call__(self, list_of_modes, list_of_noises, list_of_Zs):
        increments = list_of_noises[1:] - list_of_noises[:-1]
        increments = np.append(increments, 1 - list_of_noises[-1])
        return _argmax(increments)


class ThresholdModeChooser(ModeChooser):
    """Last chain index whose noise is below the threshold (argmin of the noise if none)."""

    def __init__(self, threshold):
        self.threshold = threshold

    def __call__(self, list_of_modes, list_of_noises, list_of_Zs):
        is_under_threshold = list_of_noises < self.threshold
        valid_indices = np.where(is_under_threshold, np.arange(list_of_noises.shape[0]), -1)
        if np.any(is_under_threshold):
            return int(np.max(valid_indices))
        return _argmin(list_of_noises)


class EarlyStopping:
    def __init__(self):
        pass

    def is_an_earlyStopping(self):
        return True


class NoEarlyStopping(EarlyStopping):
    def __init__(self):
        pass

    def __call__(self, list_of_modes, list_of_noises, list_of_Zs):
        return False


class CustomEarlyStopping(EarlyStopping):
    def __init__(self, custom_function):
        self.choice_function = custom_function

    def __call__(self, list_of_modes, list_of_noises, list_of_Zs):
        res = self.choice_function(list_of_modes, list_of_noises, list_of_Zs)
        assert res in [True, False], f"The result must be a boolean, got : {res}"
        return res
