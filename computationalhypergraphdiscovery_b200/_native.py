"""ctypes binding of libchd_b200.so (C ABI: include/chd_b200.h).

PyTorch is used for device buffers and streams only; every numeric step of the hot path is a
call into the library.  There is no CPU fallback: importing this module without the built
library, or calling it without a B200, raises.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libchd_b200.so")

KERNEL_LINEAR, KERNEL_QUADRATIC, KERNEL_GAUSSIAN = 0, 1, 2
Z_SAMPLES = 100
GRID_POINTS = 10000
BAND, BAND_LD = 32, 34
ST_BFGS_SUCCESS, ST_USED_MEDIAN, ST_GRID_EDGE = 1, 2, 4


class KernelDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("reserved", ctypes.c_int32), ("scale", ctypes.c_double),
                ("alpha", ctypes.c_double), ("quad_scale", ctypes.c_double), ("l", ctypes.c_double)]


class NativeError(RuntimeError):
    pass


_lib = None

_vp, _i, _d, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t
_SIGNATURES = {
    "chd_version": (ctypes.c_char_p, []),
    "chd_last_error": (ctypes.c_char_p, []),
    "chd_launch_count": (ctypes.c_ulonglong, []),
    "chd_plan_create": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i, _i]),
    "chd_plan_destroy": (None, [_vp]),
    "chd_workspace_bytes": (_sz, [_vp, _i]),
    "chd_plan_info": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "chd_plan_tuning": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "chd_normal": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "chd_split": (_i, [_vp, _i, _vp, _vp, _vp]),
    "chd_gram": (_i, [_vp, ctypes.POINTER(KernelDesc), _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _sz, _vp]),
    "chd_regress": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                         _sz, _vp]),
    "chd_activations": (_i, [_vp, ctypes.POINTER(KernelDesc), _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp,
                             _sz, _vp]),
    "chd_prune_chain": (_i, [_vp, ctypes.POINTER(KernelDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i,
                             _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_i), _vp, _vp, _sz, _vp]),
    "chd_predict": (_i, [_vp, ctypes.POINTER(KernelDesc), _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "chd_tridiag": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "chd_dmma_peak": (_i, [_i, ctypes.POINTER(_d)]),
    "chd_fp64_probe": (_i, [_i, _i, ctypes.POINTER(_d)]),
    "chd_dmma_occupancy_probe": (_i, [_i, _i, _i, ctypes.POINTER(_d)]),
    "chd_band_workspace_bytes": (_sz, [_vp, _i]),
    "chd_band_reduce": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "chd_band_backtransform": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "chd_band_chase": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "chd_band_solve": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "chd_plan_set_two_stage": (_i, [_vp, _i]),
    "chd_profile_enable": (_i, [_i]),
    "chd_profile_read": (_i, [ctypes.POINTER(_d), ctypes.POINTER(_i)]),
    "chd_profile_read_cat": (_i, [_i, ctypes.POINTER(_d), ctypes.POINTER(_i)]),
    "chd_plan_two_stage": (_i, [_vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load():
    """Load the shared library (idempotent). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C computationalhypergraphdiscovery_b200/csrc`). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(rc, what):
    if rc != 0:
        msg = load().chd_last_error().decode()
        raise NativeError(f"{what} failed with code {rc}: {msg}")


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device, contiguous tensors only"
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(load().chd_launch_count())


def version():
    return load().chd_version().decode()


def gamma_grid_host():
    """jnp.logspace(-6, 6, 10000) (interpolatory.py:129) evaluated on the host in FP64."""
    return 10.0 ** np.linspace(-6.0, 6.0, GRID_POINTS)


def make_desc(kind, scale, alpha=0.0, quad_scale=0.0, l=1.0):
    return KernelDesc(int(kind), 0, float(scale), float(alpha), float(quad_scale), float(l))


class Plan:
    """Problem sizes + tuning for one device (host object, chd_plan)."""

    def __init__(self, n, N, C, device=None, ctas_per_matrix=0):
        if not torch.cuda.is_available():
            raise NativeError("no CUDA device: the CHD hot path only runs on a B200 (no CPU fallback)")
        lib = load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.n, self.N, self.C = int(n), int(N), int(C)
        h = ctypes.c_void_p()
        _check(lib.chd_plan_create(ctypes.byref(h), self.n, self.N, self.C, self.device.index, int(ctas_per_matrix)),
               "chd_plan_create")
        self._h = h
        g, w, s = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _check(lib.chd_plan_info(h, ctypes.byref(g), ctypes.byref(w), ctypes.byref(s)), "chd_plan_info")
        self.ctas_per_matrix, self.matrices_per_wave, self.sm_count = g.value, w.value, s.value
        _check(lib.chd_plan_tuning(h, ctypes.byref(g), ctypes.byref(w), ctypes.byref(s)), "chd_plan_tuning")
        self.panel_width, self.threads_per_cta, self.bulk_symv = g.value, w.value, s.value
        self.two_stage = bool(lib.chd_plan_two_stage(h))
        self.ldk = (self.n + 1) // 2 * 2
        self.grid = torch.from_numpy(gamma_grid_host()).to(self.device)
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                load().chd_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def workspace_bytes(self, T):
        return int(load().chd_workspace_bytes(self._h, int(T)))

    def workspace(self, T):
        need = self.workspace_bytes(T)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    # ------------------------------------------------------------------ building blocks
    def normal(self, keys, rows, cols):
        T = keys.shape[0]
        out = torch.empty((T, rows, cols), dtype=torch.float64, device=self.device)
        _check(load().chd_normal(_ptr(keys), T, rows, cols, _ptr(out), cols, _stream()), "chd_normal")
        return out

    def split(self, keys):
        T = keys.shape[0]
        nk, sk = torch.empty_like(keys), torch.empty_like(keys)
        _check(load().chd_split(_ptr(keys), T, _ptr(nk), _ptr(sk), _stream()), "chd_split")
        return nk, sk

    def gram(self, desc, X, w, Xq=None, out=None):
        """K[t] = kernel(Xq, X, w[t]); w: (T, N) uint8."""
        T = w.shape[0]
        if Xq is None:
            Xq, nq = X, self.n
        else:
            nq = Xq.shape[0]
        if out is None:
            out = torch.zeros((T, nq, self.ldk), dtype=torch.float64, device=self.device)
        ws = self.workspace(T)
        _check(load().chd_gram(self._h, ctypes.byref(desc), _ptr(X), _ptr(Xq), nq, _ptr(w), T, _ptr(out), self.ldk,
                               _ptr(ws), ws.numel(), _stream()), "chd_gram")
        return out

    def regress(self, K, ga, keys, gamma_min, mode=0, base=1e-9):
        """K: (T, n, ldk) destroyed. Returns dict of device tensors."""
        T = K.shape[0]
        dev, f64 = self.device, torch.float64
        out = dict(yb=torch.empty((T, self.n), dtype=f64, device=dev),
                   noise=torch.empty(T, dtype=f64, device=dev), z_low=torch.empty(T, dtype=f64, device=dev),
                   z_high=torch.empty(T, dtype=f64, device=dev), gamma=torch.empty(T, dtype=f64, device=dev),
                   keys=torch.empty_like(keys), lambda_min=torch.empty(T, dtype=f64, device=dev),
                   status=torch.zeros(T, dtype=torch.int32, device=dev), gamma_min=gamma_min)
        ws = self.workspace(T)
        _check(load().chd_regress(self._h, _ptr(K), self.ldk, _ptr(ga), T, _ptr(self.grid), int(mode), float(base),
                                  _ptr(gamma_min), _ptr(keys), _ptr(out["yb"]), _ptr(out["noise"]),
                                  _ptr(out["z_low"]), _ptr(out["z_high"]), _ptr(out["gamma"]), _ptr(out["keys"]),
                                  _ptr(out["lambda_min"]), _ptr(out["status"]), _ptr(ws), ws.numel(), _stream()),
               "chd_regress")
        return out

    # ---- two-stage building blocks
    def set_two_stage(self, on):
        _check(load().chd_plan_set_two_stage(self._h, int(bool(on))), "chd_plan_set_two_stage")
        self.two_stage = bool(on)
        self._ws = None

    def band_workspace(self, T):
        need = int(load().chd_band_workspace_bytes(self._h, int(T)))
        if getattr(self, "_bws", None) is None or self._bws.numel() < need:
            self._bws = None
            self._bws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._bws

    def band_reduce(self, K, ga, B=None):
        """K (T, n, ldk) destroyed (holds the block reflectors afterwards); returns band (T, n, BAND_LD), beta0 (T).
        B (T, n, nz) is replaced by Q^T B."""
        T = K.shape[0]
        band = torch.empty((T, self.n, BAND_LD), dtype=torch.float64, device=self.device)
        beta0 = torch.empty(T, dtype=torch.float64, device=self.device)
        ws = self.band_workspace(T)
        ldb = B.shape[2] if B is not None else 0
        _check(load().chd_band_reduce(self._h, _ptr(K), self.ldk, _ptr(ga), T, _ptr(B), ldb, ldb, _ptr(band),
                                      _ptr(beta0), _ptr(ws), ws.numel(), _stream()), "chd_band_reduce")
        return band, beta0

    def band_backtransform(self, K, x, beta0):
        T = K.shape[0]
        out = torch.empty((T, self.n), dtype=torch.float64, device=self.device)
        ws = self.band_workspace(T)
        _check(load().chd_band_backtransform(self._h, _ptr(K), self.ldk, _ptr(x), _ptr(beta0), _ptr(out), T, _ptr(ws),
                                             ws.numel(), _stream()), "chd_band_backtransform")
        return out

    def band_chase(self, band):
        T = band.shape[0]
        d = torch.empty((T, self.n), dtype=torch.float64, device=self.device)
        e = torch.empty((T, self.n), dtype=torch.float64, device=self.device)
        ws = self.band_workspace(T)
        _check(load().chd_band_chase(self._h, _ptr(band), T, _ptr(d), _ptr(e), _ptr(ws), ws.numel(), _stream()),
               "chd_band_chase")
        return d, e

    def band_solve(self, band, gamma_used, Bq=None):
        T = band.shape[0]
        dev, f64 = self.device, torch.float64
        out = dict(x=torch.empty((T, self.n), dtype=f64, device=dev), noise=torch.empty(T, dtype=f64, device=dev),
                   z_low=torch.empty(T, dtype=f64, device=dev), z_high=torch.empty(T, dtype=f64, device=dev))
        ws = self.band_workspace(T)
        nz = Bq.shape[2] if Bq is not None else 0
        _check(load().chd_band_solve(self._h, _ptr(band), _ptr(gamma_used), _ptr(Bq), nz, nz, T, _ptr(out["x"]),
                                     _ptr(out["noise"]), _ptr(out["z_low"]), _ptr(out["z_high"]), _ptr(ws),
                                     ws.numel(), _stream()), "chd_band_solve")
        return out

    def tridiag(self, K, ga, B=None):
        T = K.shape[0]
        dev, f64 = self.device, torch.float64
        d = torch.empty((T, self.n), dtype=f64, device=dev)
        e = torch.empty((T, self.n), dtype=f64, device=dev)
        beta0 = torch.empty(T, dtype=f64, device=dev)
        ws = self.workspace(T)
        ldb = B.shape[2] if B is not None else 0
        _check(load().chd_tridiag(self._h, _ptr(K), self.ldk, _ptr(ga), T, _ptr(B), ldb, ldb, _ptr(d), _ptr(e),
                                  _ptr(beta0), _ptr(ws), ws.numel(), _stream()), "chd_tridiag")
        return d, e, beta0

    def activations(self, desc, X, cluster_of, w0, alive, yb, ga, prune=False):
        T = w0.shape[0]
        act = torch.empty((T, self.C), dtype=torch.float64, device=self.device)
        pruned = torch.empty(T, dtype=torch.int32, device=self.device)
        ws = self.workspace(T)
        _check(load().chd_activations(self._h, ctypes.byref(desc), _ptr(X), _ptr(cluster_of), _ptr(w0), _ptr(alive),
                                      _ptr(yb), _ptr(ga), T, _ptr(act), int(bool(prune)), _ptr(pruned), _ptr(ws),
                                      ws.numel(), _stream()), "chd_activations")
        return act, pruned

    def prune_chain(self, desc, X, cluster_of, w0, alive, ga, yb, keys, lambda_min, gamma_min, first_step, steps,
                    keep_yb=True, p_cache=None):
        """Runs `steps` prune steps device-resident; state tensors are updated in place.
        p_cache: optional dict {"P": (T, n, ldk) tensor, "valid": 0/1} carrying the Gaussian product map."""
        T = w0.shape[0]
        dev, f64 = self.device, torch.float64
        S = max(int(steps), 0)          # steps == 0: zero-length chain outputs, nothing launched
        out = dict(pruned=torch.full((T, S), -1, dtype=torch.int32, device=dev),
                   noise=torch.empty((T, S), dtype=f64, device=dev), z_low=torch.empty((T, S), dtype=f64, device=dev),
                   z_high=torch.empty((T, S), dtype=f64, device=dev), gamma=torch.empty((T, S), dtype=f64, device=dev),
                   act=torch.empty((T, S, self.C), dtype=f64, device=dev),
                   yb=torch.empty((T, S, self.n), dtype=f64, device=dev) if keep_yb else None,
                   status=torch.zeros(T, dtype=torch.int32, device=dev))
        if steps <= 0:
            return out
        ws = self.workspace(T)
        pv = ctypes.c_int(int(p_cache["valid"])) if p_cache else None
        _check(load().chd_prune_chain(self._h, ctypes.byref(desc), _ptr(X), _ptr(cluster_of), _ptr(w0), _ptr(alive),
                                      _ptr(ga), _ptr(yb), _ptr(keys), _ptr(self.grid), _ptr(lambda_min),
                                      _ptr(gamma_min), T, int(first_step), int(steps), _ptr(out["pruned"]),
                                      _ptr(out["noise"]), _ptr(out["z_low"]), _ptr(out["z_high"]), _ptr(out["gamma"]),
                                      _ptr(out["act"]), _ptr(out["yb"]),
                                      _ptr(p_cache["P"]) if p_cache else None, ctypes.byref(pv) if p_cache else None,
                                      _ptr(out["status"]), _ptr(ws), ws.numel(), _stream()), "chd_prune_chain")
        if p_cache:
            p_cache["valid"] = pv.value
        return out

    def predict(self, desc, X, Xq, w, yb):
        T, nq = w.shape[0], Xq.shape[0]
        out = torch.empty((T, nq), dtype=torch.float64, device=self.device)
        need = self.workspace_bytes(T) + T * nq * self.ldk * 8 + nq * 8 * ((self.N + 31) // 32 * 32) * T + 4096
        ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        _check(load().chd_predict(self._h, ctypes.byref(desc), _ptr(X), _ptr(Xq), nq, _ptr(w), _ptr(yb), T, _ptr(out),
                                  _ptr(ws), ws.numel(), _stream()), "chd_predict")
        return out


def profile_enable(on=True):
    _check(load().chd_profile_enable(int(bool(on))), "chd_profile_enable")


def profile_read():
    """(total ms, launches) of the tridiagonalisation kernel since profile_enable(True)."""
    ms, cnt = ctypes.c_double(), ctypes.c_int()
    _check(load().chd_profile_read(ctypes.byref(ms), ctypes.byref(cnt)), "chd_profile_read")
    return ms.value, cnt.value


PROFILE_CATEGORIES = {"tridiag": 0, "symm": 1, "syr2k": 2, "panel_qr": 3, "chase": 4, "wform": 5, "band_solve": 6,
                      "backtransform": 7}


def profile_read_all():
    """{kernel: (total ms, launches)} of the regression kernels since profile_enable(True)."""
    out = {}
    for name, cat in PROFILE_CATEGORIES.items():
        ms, cnt = ctypes.c_double(), ctypes.c_int()
        _check(load().chd_profile_read_cat(cat, ctypes.byref(ms), ctypes.byref(cnt)), "chd_profile_read_cat")
        if cnt.value:
            out[name] = (ms.value, cnt.value)
    return out


def dmma_peak_tflops(iters=20000):
    v = ctypes.c_double()
    _check(load().chd_dmma_peak(int(iters), ctypes.byref(v)), "chd_dmma_peak")
    return v.value


def dmma_occupancy_probe_tflops(warps_per_sm, chains, iters=20000):
    """DMMA rate with one CTA of `warps_per_sm` warps per SM and `chains` independent accumulators per warp."""
    v = ctypes.c_double()
    _check(load().chd_dmma_occupancy_probe(int(iters), int(warps_per_sm), int(chains), ctypes.byref(v)),
           "chd_dmma_occupancy_probe")
    return v.value


def fp64_probe_tflops(mode, iters=20000):
    """mode 0: DMMA only, 1: DFMA only, 2: both interleaved."""
    v = ctypes.c_double()
    _check(load().chd_fp64_probe(int(iters), int(mode), ctypes.byref(v)), "chd_fp64_probe")
    return v.value
