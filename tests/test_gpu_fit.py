"""End-to-end parity of GraphDiscovery.fit() (CUDA, through the C ABI) against the oracle and the
reference's golden notebook outputs."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import chd as ochd  # noqa: E402
from algebraic_data import algebraic_examples, predictions_example1  # noqa: E402

KN = ["linear", "quadratic", "gaussian"]
DUP = [{"$x_1$": "$w_1$", "$x_2$": "$w_2$"}, {"$x_1$": "$w_1$", "$x_3$": "$w_3$"}, {}, {"$x_1$": "$w_1$"}]


def _canon(names, dup):
    return sorted(dict(dup).get(nm, nm) for nm in names)


def _fit_cuda(X, names, kernels=None, **kw):
    import computationalhypergraphdiscovery_b200 as CHD
    gd = CHD.GraphDiscovery(X, names, kernels=kernels, verbose=False, **kw)
    gd.fit()
    return gd


def _compare_with_oracle(gd, res, names, dup, check_chain=True):
    st = gd.last_fit["stage1"]
    np.testing.assert_allclose(st[1], res["stage1"]["noises"], rtol=1e-8)
    np.testing.assert_allclose(st[2], res["stage1"]["Z_lows"], rtol=1e-8)
    np.testing.assert_allclose(st[3], res["stage1"]["Z_highs"], rtol=1e-8)
    np.testing.assert_allclose(st[4], res["stage1"]["gammas"], rtol=1e-6)
    assert (st[5] == res["stage1"]["keys"]).all()
    assert (gd.last_fit["chosen_kernel"] == res["chosen_kernel"]).all()
    for name in names:
        node = res["nodes"].get(name)
        data = gd.G.nodes[name]
        if node is None:
            assert "kernel_index" not in data and gd.G.in_degree(name) == 0
            continue
        assert data["kernel_index"] == node["kernel_index"] and data["type"] == node["type"]
        got = [a for a, _ in gd.G.in_edges(name)]
        assert _canon(got, dup) == _canon(node["ancestors"], dup), name
    if check_chain and not dup:
        for k, ch in res["chains"].items():
            mine = gd.last_fit["chains"][k]
            assert mine["targets"] == ch["targets"]
            assert (mine["active"] == ch["active"]).all()
            np.testing.assert_allclose(mine["noises"], ch["noises"], rtol=1e-8)
            np.testing.assert_allclose(mine["Z_lows"], ch["Z_lows"], rtol=1e-8)
            np.testing.assert_allclose(mine["gammas"], ch["gammas"], rtol=1e-6)
            m = ch["activations"] != 2.0
            np.testing.assert_allclose(mine["activations"][m], ch["activations"][m], rtol=1e-8)
            assert (mine["chosen_modes"] == ch["chosen_modes"]).all()
            for i, name in enumerate(ch["targets"]):
                node = res["nodes"][name]
                yb = gd.G.nodes[name]["coeff"]
                assert np.linalg.norm(yb - node["coeff"]) <= 1e-7 * np.linalg.norm(node["coeff"])
                assert abs(gd.G.nodes[name]["noise"] - node["noise"]) <= 1e-8 * abs(node["noise"])
                for a, _b, sig in gd.G.in_edges(name, data="signal"):
                    assert abs(sig - node["signals"][a]) <= 1e-8 * abs(node["signals"][a])


def test_synthetic_fit_matches_oracle():
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, names = make_data(300, 8, seed=0)
    gd = _fit_cuda(X, names)
    res = ochd.fit(X, names)
    _compare_with_oracle(gd, res, names, {})


def test_sachs_shaped_linear_quadratic():
    """Config 1 shape: n=500, N=11, LinearMode + QuadraticMode."""
    import computationalhypergraphdiscovery_b200 as CHD
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, names = make_data(500, 11, seed=1)
    gd = _fit_cuda(X, names, kernels=[CHD.Modes.LinearMode(), CHD.Modes.QuadraticMode()])
    res = ochd.fit(X, names, kernels=[("linear", 2.0, None), ("quadratic", 1.0, None)])
    _compare_with_oracle(gd, res, names, {})


def test_predictions_example1_golden(golden_dir):
    import computationalhypergraphdiscovery_b200 as CHD
    gold = json.load(open(os.path.join(golden_dir, "predictions_example1.json")))["fit"]
    X, names = predictions_example1()
    gd = _fit_cuda(X, names, kernels=[CHD.Modes.LinearMode(), CHD.Modes.QuadraticMode(), CHD.Modes.GaussianMode(l=0.5)])
    st = gd.last_fit["stage1"]
    for t, name in enumerate(names):
        g = gold["nodes"][name]
        for k, kn in enumerate(KN):
            ref = g["stage1"][kn]
            assert abs(st[1][t, k] - ref["noise"]) <= 1e-8 * abs(ref["noise"]), (name, kn)
            assert f"{st[2][t, k]:.2f}" == ref["Z_low_2dp"] and f"{st[3][t, k]:.2f}" == ref["Z_high_2dp"]
            assert f"{max(1e-9, st[4][t, k]):.2e}" == ref["gamma_print"]
        got = [a for a, _ in gd.G.in_edges(name)]
        assert _canon(got, DUP[0]) == _canon(g["ancestors"], DUP[0]), name
    assert f"{gd.G.nodes['$x_1$']['gamma']:.3g}" == "0.0905"
    assert f"{gd.G.nodes['$x_2$']['noise']:.3g}" == "0.000991"


@pytest.mark.parametrize("ex", [0, 2])
def test_algebraic_golden(golden_dir, ex):
    gold = json.load(open(os.path.join(golden_dir, "algebraic.json")))["examples"][ex]
    X, names = algebraic_examples()[ex]
    gd = _fit_cuda(X, names)
    st = gd.last_fit["stage1"]
    for t, name in enumerate(names):
        g = gold["fit"]["nodes"][name]
        for k, kn in enumerate(KN):
            ref = g["stage1"][kn]
            assert abs(st[1][t, k] - ref["noise"]) <= 1e-8 * abs(ref["noise"]), (name, kn, st[1][t, k], ref["noise"])
            assert f"{max(1e-9, st[4][t, k]):.2e}" == ref["gamma_print"], (name, kn)
        got = [a for a, _ in gd.G.in_edges(name)]
        assert _canon(got, DUP[ex]) == _canon(g["ancestors"], DUP[ex]), name
        if g["kernel"] is not None:
            assert gd.G.nodes[name]["type"] == g["kernel"]
    for short, d in gold["dependencies"].items():
        node = gd.G.nodes["$" + short + "$"]
        assert f"{node['gamma']:.3g}" == d["gamma_3sf"], (short, node["gamma"])
        assert f"{node['noise']:.3g}" == d["noise_3sf"], (short, node["noise"])
