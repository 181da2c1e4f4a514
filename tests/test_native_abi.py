"""CPU-side checks of the boundary: the library builds/loads, exports every symbol the header
declares, and the host mirror of the reference API behaves like the reference (no compute calls)."""
import ctypes
import os
import pickle
import re

import networkx as nx
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from computationalhypergraphdiscovery_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _native.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "chd_b200.h")).read()
    declared = set(re.findall(r"\b(chd_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), name
    from computationalhypergraphdiscovery_b200 import _native
    assert set(_native.EXPORTED_SYMBOLS) == declared


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.chd_version()
    assert isinstance(lib.chd_last_error(), bytes)
    assert lib.chd_launch_count() == 0 or lib.chd_launch_count() > 0


def test_kernel_desc_layout_matches_header():
    from computationalhypergraphdiscovery_b200 import _native
    assert ctypes.sizeof(_native.KernelDesc) == 40


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import computationalhypergraphdiscovery_b200 as CHD
    X = np.random.default_rng(0).standard_normal((40, 4))
    gd = CHD.GraphDiscovery(X, list("abcd"), verbose=False)
    with pytest.raises(Exception, match="no CPU fallback"):
        gd.fit()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "computationalhypergraphdiscovery_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_constructor_validation_and_modes():
    import computationalhypergraphdiscovery_b200 as CHD
    rng = np.random.default_rng(0)
    X = rng.standard_normal((30, 3))
    with pytest.raises(AssertionError):
        CHD.GraphDiscovery(X, ["a", "b"], verbose=False)
    Xc = X.copy()
    Xc[:, 1] = 1.0
    with pytest.raises(ValueError):
        CHD.GraphDiscovery(Xc, list("abc"), verbose=False)
    with pytest.raises(AssertionError):
        CHD.GraphDiscovery(X, list("abc"), kernels=[CHD.Modes.LinearMode(), CHD.Modes.LinearMode()], verbose=False)
    with pytest.raises(KeyError):   # GaussianMode needs a QuadraticMode (reference KERN:302)
        CHD.GraphDiscovery(X, list("abc"), kernels=[CHD.Modes.GaussianMode(l=1.0)], verbose=False)
    gd = CHD.GraphDiscovery(X, list("abc"), verbose=False)
    np.testing.assert_allclose(gd.X.mean(0), 0, atol=1e-12)
    np.testing.assert_allclose(gd.X.std(0), 1, atol=1e-12)
    with pytest.raises(ValueError):
        gd.kernels = []
    k = 0.1 * CHD.Modes.QuadraticMode()
    assert k.scale == 0.1 and (2 * CHD.Modes.LinearMode()) is None   # int multiplier returns None (KERN:61-68)
    g = CHD.Modes.GaussianMode(l=0.5)
    with pytest.raises(AttributeError):
        g.l = 2.0
    gd2 = CHD.GraphDiscovery(X, list("abc"), kernels=[0.1 * CHD.Modes.LinearMode(), 0.5 * CHD.Modes.QuadraticMode(),
                                                    CHD.Modes.GaussianMode(l=2.0)], verbose=False)
    d = gd2.kernels[2].desc
    assert (d.kind, d.scale, d.quad_scale, d.l) == (2, 1.0, 0.5, 2.0) and d.alpha == 0.5 * 0.1 / 0.5


def test_masks_clusters_and_possible_edges():
    import computationalhypergraphdiscovery_b200 as CHD
    from oracle import chd as ochd
    rng = np.random.default_rng(1)
    names = list("abcde")
    X = rng.standard_normal((30, 5))
    pe = nx.DiGraph()
    pe.add_edges_from([("a", "b"), ("c", "b"), ("b", "b"), ("d", "e"), ("a", "e")])
    clusters = [["a", "c"], ["b"], ["d", "e"]]
    gd = CHD.GraphDiscovery(X, names, clusters=clusters, possible_edges=pe, verbose=False)
    assert not pe.has_edge("b", "b")          # self loops removed on the caller's graph (MAIN:112)
    idx, w0 = gd._prepare_modes_for_ancestor_finding(["b", "e"])
    im = ochd.make_index_matrix(names, clusters)
    np.testing.assert_array_equal(gd.modes.index_matrix, im)
    ref = ochd.initial_masks(im, idx, gd.possible_edges_adjacency)
    np.testing.assert_array_equal(gd._active_modes(w0), ref)
    assert gd.modes.cluster_of.tolist() == [0, 1, 0, 2, 2]
    with pytest.raises(AssertionError):
        CHD.GraphDiscovery(X, names, clusters=[["a"], ["b"]], verbose=False)


def test_pickle_roundtrip_and_cluster_graph():
    import computationalhypergraphdiscovery_b200 as CHD
    rng = np.random.default_rng(2)
    names = list("abcd")
    gd = CHD.GraphDiscovery(rng.standard_normal((20, 4)), names, verbose=False)
    gd.G.add_edge("a", "c", signal=0.5)
    gd.G.add_edge("c", "a", signal=0.5)
    gd.G.add_edge("a", "b", signal=0.1)
    gd2 = pickle.loads(pickle.dumps(gd))
    assert list(gd2.G.edges) == list(gd.G.edges) and gd2.kernels[1].desc.alpha == 1.0
    new = gd.prepare_new_graph_with_clusters([["a", "b"], ["c", "d"]])
    # edges from an earlier cluster to a later one are removed, the rest inherited (MAIN:210-222)
    assert not new.G.has_edge("a", "c") and new.G.has_edge("c", "a") and new.G.has_edge("a", "b")


def test_choosers_match_oracle():
    import computationalhypergraphdiscovery_b200 as CHD
    from oracle import chd as ochd
    rng = np.random.default_rng(3)
    for _ in range(200):
        nk = 3
        noises = rng.random(nk)
        zl = rng.random(nk)
        perf = [np.zeros((nk, 4)), noises, zl, zl + 0.1, rng.random(nk), np.zeros((nk, 2), np.uint32)]
        got = CHD.decision.MinNoiseKernelChooser()(perf)[0]
        assert got == ochd.min_noise_kernel_chooser(noises, zl)[0]
        thr = rng.random()
        assert CHD.decision.ThresholdKernelChooser(thr)(perf)[0] == ochd.threshold_kernel_chooser(noises, thr)[0]
        chain = rng.random(7)
        assert CHD.decision.MaxIncrementModeChooser()(None, chain, None) == ochd.max_increment_mode_chooser(chain)
        assert CHD.decision.ThresholdModeChooser(thr)(None, chain, None) == ochd.threshold_mode_chooser(chain, thr)


def test_anomaly_clamp_matches_oracle(capsys):
    import computationalhypergraphdiscovery_b200 as CHD
    from oracle import chd as ochd
    a = np.array([[0.1, 0.2, 1.2, 0.3], [0.1, 0.2, 0.3, 0.4], [-0.2, 0.5, 0.5, 0.5]])
    n, zl, zh = CHD.GraphDiscovery._anomaly_clamp(a, a * 0.5, a * 0.9)
    np.testing.assert_array_equal(n, ochd.anomaly_clamp(a))
    np.testing.assert_array_equal(zl, ochd.anomaly_clamp(a * 0.5))
    assert "Anomaly" in capsys.readouterr().out
