"""Pin the oracle against the reference's stored notebook outputs
(tests/golden/*.json, harvested by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import chd
from algebraic_data import algebraic_examples, predictions_example1

KN = ["linear", "quadratic", "gaussian"]
# exact-duplicate columns (x_1 == w_1 in ex1/ex2/ex4, x_3 == w_3 in ex2) tie to the last bit;
# the reference itself is not reproducible across BLAS builds there.
DUP = [{"$x_1$": "$w_1$", "$x_2$": "$w_2$"}, {"$x_1$": "$w_1$", "$x_3$": "$w_3$"}, {}, {"$x_1$": "$w_1$"}]


def _canon(names, dup):
    m = dict(dup)
    return sorted(m.get(nm, nm) for nm in names)


def _check_fit(res, gold, names, dup, rel_tol):
    worst = 0.0
    for t, name in enumerate(names):
        g = gold["nodes"][name]
        for k, kn in enumerate(KN):
            ref = g["stage1"][kn]
            v = res["stage1"]["noises"][t, k]
            worst = max(worst, abs(v - ref["noise"]) / abs(ref["noise"]))
            assert f"{res['stage1']['Z_lows'][t, k]:.2f}" == ref["Z_low_2dp"], (name, kn)
            assert f"{res['stage1']['Z_highs'][t, k]:.2f}" == ref["Z_high_2dp"], (name, kn)
            assert f"{max(1e-9, res['stage1']['gammas'][t, k]):.2e}" == ref["gamma_print"], (name, kn)
        node = res["nodes"].get(name)
        if g["kernel"] is None:
            assert node is None, name
        else:
            assert node is not None and node["type"] == g["kernel"], name
            assert _canon(node["ancestors"], dup) == _canon(g["ancestors"], dup), name
            assert f"{node['noise']:.2f}" == g["noise_after_2dp"], name
    assert worst < rel_tol, worst


def _check_deps(res, deps, dup):
    for short, d in deps.items():
        node = res["nodes"]["$" + short + "$"]
        assert node["type"] == d["kernel"]
        assert f"{node['gamma']:.3g}" == d["gamma_3sf"], (short, node["gamma"])
        assert f"{node['noise']:.3g}" == d["noise_3sf"], (short, node["noise"])
        got = _canon(node["ancestors"], dup)
        want = _canon(["$" + a + "$" for a in d["ancestors"]], dup)
        assert got == want, short


def test_predictions_example1(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "predictions_example1.json")))["fit"]
    X, names = predictions_example1()
    res = chd.fit(X, names, kernels=[("linear", 2.0, None), ("quadratic", 1.0, None), ("gaussian", 1.0, 0.5)])
    _check_fit(res, gold, names, DUP[0], 1e-12)
    assert f"{res['nodes']['$x_1$']['gamma']:.3g}" == "0.0905"
    assert f"{res['nodes']['$x_2$']['noise']:.3g}" == "0.000991"


@pytest.mark.parametrize("ex", [0, 1, 2, 3])
def test_algebraic_example(golden_dir, ex):
    gold = json.load(open(os.path.join(golden_dir, "algebraic.json")))["examples"][ex]
    X, names = algebraic_examples()[ex]
    res = chd.fit(X, names)
    _check_fit(res, gold["fit"], names, DUP[ex], 1e-9)
    _check_deps(res, gold["dependencies"], DUP[ex])
