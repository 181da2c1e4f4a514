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


# Bookkeeping of the BFGS `success`-flag escape hatch: how often it fired, over how many
# compared chains.  conftest.py writes it to gpurun_out/parity_notes.json at the end of a GPU session; bench.py
# quotes the committed copy (profiles/) as `parity_notes`.
FLIP_STATS = {"fits_compared": 0, "chains_compared": 0, "chain_steps_compared": 0, "flag_flips": 0,
              "chain_steps_excluded": 0, "flips": []}


def flip_report():
    f = FLIP_STATS
    return (f"BFGS success-flag flips so far: {f['flag_flips']} in {f['chains_compared']} chains / "
            f"{f['chain_steps_compared']} chain steps ({f['chain_steps_excluded']} steps excluded after a flip)")


def _force_chooser(k):
    """A user KernelChooser (DEC:8-31 interface) that always picks kernel k."""
    import computationalhypergraphdiscovery_b200 as CHD

    class Force(CHD.decision.KernelChooser):
        def __call__(self, perfs):
            ybs, noises, zl, zh, gammas, keys = perfs
            return k, ybs[k], noises[k], zl[k], zh[k], gammas[k], keys[k]

    return Force()


def _median_count_alternatives(K, ga, gamma_min):
    """When find_gamma falls back to the median eigenvalue (interpolatory.py:118-122, 134) it first drops the
    eigenvalues below gamma_min.  For a numerically rank-deficient Gram matrix those are rounding noise of size
    n eps |K|: how many of them fall below gamma_min = 1e-9 differs between eigvalsh and eigh of the
    SAME LAPACK, and decides between neighbouring medians (a 1e-3 effect on the noise ratio).  Returns the
    (gamma, noise) pairs of every count that rounding can produce; an independent eigensolver is accepted if it
    lands on one of them."""
    lam, V = np.linalg.eigh(K)
    n = lam.size
    band = 10 * n * np.finfo(float).eps * float(np.max(np.abs(lam)))
    nb = int((lam < gamma_min).sum())
    out = []
    for c in range(0, n - 1):
        moved = lam[min(c, nb):max(c, nb)]            # eigenvalues whose side of gamma_min changes
        if moved.size and np.max(np.abs(moved - gamma_min)) > band:
            continue
        med = float(np.median(lam[c:]))
        gu = max(med, gamma_min)
        Pga = V.T @ ga
        cf = gu / (lam + gu)
        out.append((med, float(np.dot(Pga * cf, Pga * cf) / np.dot(Pga * cf, Pga)), c - nb))
    return out


def _flag_flips(gd, res):
    """{target name: first chain step} where gamma differs from the oracle's because the BFGS `success` flag of
    interpolatory.py:133-134 came out the other way.  Near convergence that flag is decided by rounding noise of
    the loss, so the reference's own outcome is one of two candidates -- the
    optimiser's x or the median eigenvalue -- and which one is not reproducible by any independent implementation.
    A flip is only accepted if our gamma IS the reference's other candidate; everything after it is incomparable."""
    flips = {}
    for k, ch in res["chains"].items():
        mine = gd.last_fit["chains"][k]
        tol = np.clip(1e-8 * 1e-3 / np.maximum(np.abs(ch["gammas"]), 1e-300), 1e-8, 1e-5)
        bad = np.abs(mine["gammas"] - ch["gammas"]) > np.maximum(1e-6, 100 * tol) * np.abs(ch["gammas"])
        for i, name in enumerate(ch["targets"]):
            if bad[i].any():
                s0 = int(np.argmax(bad[i]))
                alt = ch["alt_gammas"][i, s0]
                assert abs(mine["gammas"][i, s0] - alt) <= 1e-5 * abs(alt), (name, s0, mine["gammas"][i, s0], alt)
                flips[name] = s0
                FLIP_STATS["flag_flips"] += 1
                FLIP_STATS["chain_steps_excluded"] += int(ch["gammas"].shape[1] - s0)
                FLIP_STATS["flips"].append({"target": str(name), "kernel": int(k), "step": s0,
                                            "n": int(gd.X.shape[0]), "N": int(gd.X.shape[1]),
                                            "gamma_oracle": float(ch["gammas"][i, s0]),
                                            "gamma_cuda": float(mine["gammas"][i, s0])})
        FLIP_STATS["chains_compared"] += len(ch["targets"])
        FLIP_STATS["chain_steps_compared"] += int(ch["gammas"].size)
    FLIP_STATS["fits_compared"] += 1
    if flips:
        print(f"\n[parity] {len(flips)} BFGS success-flag flip(s) in this fit: {flips}; {flip_report()}")
    return flips


def _compare_with_oracle(gd, res, names, dup, check_chain=True):
    st = gd.last_fit["stage1"]
    np.testing.assert_allclose(st[1], res["stage1"]["noises"], rtol=1e-8)
    np.testing.assert_allclose(st[2], res["stage1"]["Z_lows"], rtol=1e-8)
    np.testing.assert_allclose(st[3], res["stage1"]["Z_highs"], rtol=1e-8)
    np.testing.assert_allclose(st[4], res["stage1"]["gammas"], rtol=1e-6)
    assert (st[5] == res["stage1"]["keys"]).all()
    assert (gd.last_fit["chosen_kernel"] == res["chosen_kernel"]).all()
    flips = _flag_flips(gd, res) if (check_chain and not dup) else {}
    assert len(flips) <= max(1, len(names) // 8), flips      # rare by construction
    for name in names:
        node = res["nodes"].get(name)
        data = gd.G.nodes[name]
        if node is None:
            assert "kernel_index" not in data and gd.G.in_degree(name) == 0
            continue
        assert data["kernel_index"] == node["kernel_index"] and data["type"] == node["type"]
        if name in flips:
            continue
        got = [a for a, _ in gd.G.in_edges(name)]
        assert _canon(got, dup) == _canon(node["ancestors"], dup), name
    if check_chain and not dup:
        for k, ch in res["chains"].items():
            mine = gd.last_fit["chains"][k]
            assert mine["targets"] == ch["targets"]
            S = ch["gammas"].shape[1]
            ok = np.ones(ch["gammas"].shape, dtype=bool)          # comparable (target, step) pairs
            for i, name in enumerate(ch["targets"]):
                if name in flips:
                    ok[i, flips[name]:] = False
            assert (mine["active"] == ch["active"])[np.broadcast_to(ok[:, :, None], ch["active"].shape)].all()
            # 1e-8 relative where the regression is well conditioned; both implementations lose accuracy like
            # eps * |K| / gamma, so the bound is scaled for gamma < 1e-3 (capped at 1e-5)
            tol = np.clip(1e-8 * 1e-3 / np.maximum(np.abs(ch["gammas"]), 1e-300), 1e-8, 1e-5)
            assert (np.abs(mine["noises"] - ch["noises"]) <= tol * np.abs(ch["noises"]))[ok].all()
            assert (np.abs(mine["Z_lows"] - ch["Z_lows"]) <= tol * np.abs(ch["Z_lows"]))[ok].all()
            assert (np.abs(mine["gammas"] - ch["gammas"]) <= np.maximum(1e-6, 100 * tol) * np.abs(ch["gammas"]))[ok].all()
            m = ch["activations"] != 2.0
            na = ch["activations"].shape[1]
            tol_a = np.broadcast_to(tol[:, :na, None], ch["activations"].shape)
            # the activations of step s use the regression of step s (flipped step itself is already incomparable)
            m = m & np.broadcast_to(ok[:, :na, None], ch["activations"].shape)
            assert (np.abs(mine["activations"] - ch["activations"])[m] <= (tol_a * np.abs(ch["activations"]))[m]).all()
            for i, name in enumerate(ch["targets"]):
                if name in flips:
                    continue
                assert (mine["chosen_modes"][i] == ch["chosen_modes"][i]).all()
                node = res["nodes"][name]
                yb = gd.G.nodes[name]["coeff"]
                t_i = float(tol[i, node["chosen_mode"]])
                assert np.linalg.norm(yb - node["coeff"]) <= 10 * t_i * np.linalg.norm(node["coeff"])
                assert abs(gd.G.nodes[name]["noise"] - node["noise"]) <= t_i * abs(node["noise"])
                for a, _b, sig in gd.G.in_edges(name, data="signal"):
                    assert abs(sig - node["signals"][a]) <= t_i * abs(node["signals"][a])


def test_synthetic_fit_matches_oracle():
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, names = make_data(300, 8, seed=0)
    gd = _fit_cuda(X, names)
    res = ochd.fit(X, names)
    _compare_with_oracle(gd, res, names, {})


def test_multi_batch_gaussian_chain_matches_oracle():
    """a kernel group larger than one launch batch (forced here: 3 targets per batch) reuses ONE Gaussian
    product-map buffer across its batches and writes the chain slices batch by batch."""
    import computationalhypergraphdiscovery_b200 as CHD
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, names = make_data(160, 8, seed=21)
    gd = CHD.GraphDiscovery(X, names, verbose=False)
    gd._force_batch_size = 3
    gd.fit(kernel_chooser=_force_chooser(2))
    res = ochd.fit(X, names, kernel_chooser=("force", 2))
    assert gd.last_fit["chains"][2]["noises"].shape == (8, 7)
    _compare_with_oracle(gd, res, names, {})


def test_group_with_empty_initial_masks_runs_one_step():
    """ADVICE r1 / MAIN:619-660: with possible_edges a target can start with no candidate ancestor at all; the
    reference still runs ONE prune step (the `all masks empty` check comes after the step).  The regression of that
    step is on the constant Gram matrix (rank-deficient: values are not comparable), so only the
    shape of the chain and the absence of garbage are checked."""
    import networkx as nx
    import computationalhypergraphdiscovery_b200 as CHD
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, names = make_data(120, 5, seed=22)
    pe = nx.DiGraph()
    pe.add_nodes_from(names)
    pe.add_edges_from((a, b) for a in names for b in names[1:] if a != b)      # nothing may point to names[0]
    gd = CHD.GraphDiscovery(X, names, possible_edges=pe, verbose=False)
    gd.fit(targets=[names[0]], kernel_chooser=_force_chooser(1))
    ch = gd.last_fit["chains"][1]
    assert ch["noises"].shape == (1, 2) and ch["activations"].shape == (1, 1, 5)
    assert (ch["activations"] == 2.0).all()
    assert (ch["active"] == 0).all()
    assert np.isfinite(ch["noises"]).all() or np.isnan(ch["noises"][0, 1])
    assert gd.G.in_degree(names[0]) == 0


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


@pytest.mark.parametrize("ex", [0, 1, 2, 3])
def test_algebraic_golden(golden_dir, ex):
    """The reference's stored notebook outputs (ALG-NB:82-188, 290-436, 555-691, 792-939, 998-1105) for all four
    algebraic examples.  A stage-1 value may differ from the notebook's only through the rounding-decided count of
    eigenvalues below gamma_min (see _median_count_alternatives); such cases are counted in FLIP_STATS."""
    from oracle import kernels as okern
    gold = json.load(open(os.path.join(golden_dir, "algebraic.json")))["examples"][ex]
    X, names = algebraic_examples()[ex]
    gd = _fit_cuda(X, names)
    st = gd.last_fit["stage1"]
    specs = okern.make_specs([("linear", 2.0, None), ("quadratic", 1.0, None), ("gaussian", 1.0, 1.0)])
    touched = set()
    for t, name in enumerate(names):
        g = gold["fit"]["nodes"][name]
        for k, kn in enumerate(KN):
            ref = g["stage1"][kn]
            if abs(st[1][t, k] - ref["noise"]) <= 1e-8 * abs(ref["noise"]):
                assert f"{max(1e-9, st[4][t, k]):.2e}" == ref["gamma_print"], (name, kn)
                continue
            w = np.ones(len(names))
            w[t] = 0
            alts = _median_count_alternatives(okern.gram(specs[k], gd.X, gd.X, w), gd.X[:, t], 1e-9)
            hit = [a for a in alts if abs(st[1][t, k] - a[1]) <= 1e-8 * abs(a[1]) and abs(st[4][t, k] - a[0]) <= 1e-8 * a[0]]
            assert hit, (name, kn, st[1][t, k], ref["noise"], alts[:5])
            FLIP_STATS["median_count_flips"] = FLIP_STATS.get("median_count_flips", 0) + 1
            FLIP_STATS["flips"].append({"example": ex, "target": name, "kernel": kn, "kind": "eigenvalue count below "
                                        "gamma_min decided by rounding", "count_shift": hit[0][2],
                                        "noise_cuda": float(st[1][t, k]), "noise_notebook": ref["noise"]})
            print(f"\n[parity] {name}/{kn}: median-eigenvalue gamma with a rounding-decided count (shift {hit[0][2]}): "
                  f"noise {st[1][t, k]} vs notebook {ref['noise']}")
            touched.add(name)
    FLIP_STATS["golden_stage1_values_compared"] = FLIP_STATS.get("golden_stage1_values_compared", 0) + 3 * len(names)
    for t, name in enumerate(names):
        g = gold["fit"]["nodes"][name]
        if name in touched:
            continue          # the node's chain starts from the other median: not comparable with the notebook
        got = [a for a, _ in gd.G.in_edges(name)]
        assert _canon(got, DUP[ex]) == _canon(g["ancestors"], DUP[ex]), name
        if g["kernel"] is not None:
            assert gd.G.nodes[name]["type"] == g["kernel"]
    for short, d in gold["dependencies"].items():
        if "$" + short + "$" in touched:
            continue
        node = gd.G.nodes["$" + short + "$"]
        assert f"{node['gamma']:.3g}" == d["gamma_3sf"], (short, node["gamma"])
        assert f"{node['noise']:.3g}" == d["noise_3sf"], (short, node["noise"])


def test_clusters_possible_edges_and_threshold_choosers():
    """Config-2 style workflow: clusters + possible_edges + scaled kernels + threshold choosers."""
    import networkx as nx
    import computationalhypergraphdiscovery_b200 as CHD
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, names = make_data(260, 9, seed=3)
    clusters = [names[0:3], names[3:5], names[5:6], names[6:9]]
    rng = np.random.default_rng(4)
    pe = nx.DiGraph()
    pe.add_nodes_from(names)
    for a in names:
        for b in names:
            if a != b and rng.random() > 0.2:
                pe.add_edge(a, b)
    kernels = [0.5 * CHD.Modes.LinearMode(), 0.5 * CHD.Modes.QuadraticMode(), 0.7 * CHD.Modes.GaussianMode(l=1.5)]
    gd = CHD.GraphDiscovery(X, names, kernels=kernels, clusters=clusters, possible_edges=pe, verbose=False,
                            gamma_min=1e-8)
    gd.fit(kernel_chooser=CHD.decision.ThresholdKernelChooser(0.9), mode_chooser=CHD.decision.ThresholdModeChooser(0.5))
    adj = np.array(nx.adjacency_matrix(pe, nodelist=names).todense())
    res = ochd.fit(X, names, kernels=[("linear", 0.5, None), ("quadratic", 0.5, None), ("gaussian", 0.7, 1.5)],
                   clusters=clusters, adjacency=adj, gamma_min=1e-8, kernel_chooser=("threshold", 0.9),
                   mode_chooser=("threshold", 0.5))
    _compare_with_oracle(gd, res, names, {})


def test_loop_number_zero_two_clusters():
    """C = 2 clusters -> no prune step, activations of the full set only (MAIN:662-691)."""
    import computationalhypergraphdiscovery_b200 as CHD
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, names = make_data(150, 4, seed=5)
    clusters = [names[:2], names[2:]]
    gd = CHD.GraphDiscovery(X, names, clusters=clusters, verbose=False)
    gd.fit()
    res = ochd.fit(X, names, clusters=clusters)
    _compare_with_oracle(gd, res, names, {})


def test_predict_and_pickle_roundtrip():
    import pickle
    import computationalhypergraphdiscovery_b200 as CHD
    from oracle import kernels as okern
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, names = make_data(220, 6, seed=6)
    gd = _fit_cuda(X, names)
    Xq, _ = make_data(57, 6, seed=7)
    pred = gd.predict(names, Xq)
    specs = okern.make_specs([("linear", 2.0, None), ("quadratic", 1.0, None), ("gaussian", 1.0, 1.0)])
    Xn = (X - X.mean(0)) / X.std(0)
    Xqn = (Xq - X.mean(0)) / X.std(0)
    for j, name in enumerate(names):
        node = gd.G.nodes[name]
        if "kernel_index" not in node:
            ref = np.zeros(57) * X.std(0)[j] + X.mean(0)[j]
        else:
            K = okern.gram(specs[node["kernel_index"]], Xqn, Xn, np.asarray(node["active_modes"], dtype=float))
            ref = (K @ (-node["coeff"])) * X.std(0)[j] + X.mean(0)[j]
        np.testing.assert_allclose(pred[:, j], ref, rtol=1e-9, atol=1e-9)
    gd2 = pickle.loads(pickle.dumps(gd))
    assert sorted(gd2.G.edges) == sorted(gd.G.edges)
    np.testing.assert_allclose(gd2.predict(names, Xq), pred, rtol=1e-12, atol=1e-12)
    new = gd2.prepare_new_graph_with_clusters([names[:3], names[3:]])
    new.print_func = lambda *a, **k: None
    new.fit()
    assert new.G.number_of_nodes() == 6


def test_targets_subset_keeps_full_run_subkeys():
    """fit(targets=subset) uses split(key, len(targets)+1) like the reference (MAIN:311)."""
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    import computationalhypergraphdiscovery_b200 as CHD
    X, names = make_data(180, 6, seed=8)
    gd = CHD.GraphDiscovery(X, names, verbose=False)
    gd.fit(targets=names[1:4], key=CHD.random.PRNGKey(7))
    from oracle import jaxprng
    res = ochd.fit(X, names, targets=names[1:4], key=jaxprng.PRNGKey(7))
    st = gd.last_fit["stage1"]
    np.testing.assert_allclose(st[1], res["stage1"]["noises"], rtol=1e-8)
    np.testing.assert_allclose(st[2], res["stage1"]["Z_lows"], rtol=1e-8)
    assert (gd.last_fit["chosen_kernel"] == res["chosen_kernel"]).all()


def test_custom_mode_kernel_extension_hook():
    """A user ModeKernel (NumPy `kernel(X, Y, which_dim)`, KERN:7-68) goes through the generic path and must
    reproduce the built-in quadratic mode it re-implements."""
    import computationalhypergraphdiscovery_b200 as CHD
    from computationalhypergraphdiscovery_b200.synthetic import make_data

    class MyQuadratic(CHD.Modes.ModeKernel):
        def __init__(self):
            super().__init__()
            self.name = "myquad"
            self._is_interpolatory = False
            self._memory_efficient_required = False

        def setup(self, X, scales):
            assert self.scale == scales[self.name]
            self.alpha = 0.5 * 2.0 / self.scale

        def kernel(self, X, Y, which_dim):
            L = (X * which_dim[None, :]) @ Y.T
            return self.scale * (self.alpha + L) ** 2 + (1 - self.alpha ** 2 * self.scale)

    X, names = make_data(200, 6, seed=9)
    ref = CHD.GraphDiscovery(X, names, kernels=[CHD.Modes.QuadraticMode()], verbose=False)
    ref.fit()
    gd = CHD.GraphDiscovery(X, names, kernels=[MyQuadratic()], verbose=False)
    gd.fit()
    assert sorted(gd.G.edges) == sorted(ref.G.edges)
    np.testing.assert_allclose(gd.last_fit["stage1"][1], ref.last_fit["stage1"][1], rtol=1e-9)
    for nm in names:
        if "type" in gd.G.nodes[nm]:
            assert gd.G.nodes[nm]["type"] == "myquad"


def _user_kernels():
    import computationalhypergraphdiscovery_b200 as CHD

    class Cubic(CHD.Modes.ModeKernel):
        """K = 1 + s (L + L^3 / 4): not one of the built-in modes, default individual_influence (KERN:39-56)."""

        def __init__(self):
            super().__init__()
            self.name = "cubic"
            self._is_interpolatory = False
            self._memory_efficient_required = False

        def setup(self, X, scales):
            assert self.scale == scales[self.name]

        def kernel(self, X, Y, which_dim):
            L = (X * which_dim[None, :]) @ Y.T
            return 1 + self.scale * (L + L ** 3 / 4)

    class Laplace(CHD.Modes.ModeKernel):
        """K = prod_d (1 + w_d exp(-|x_d - y_d|)) with its own individual_influence (the hook of KERN:39-56)."""

        def __init__(self):
            super().__init__()
            self.name = "laplace"
            self._is_interpolatory = True
            self._memory_efficient_required = True

        def setup(self, X, scales):
            assert self.scale == scales[self.name]

        def _prod(self, X, Y, w):
            P = np.ones((X.shape[0], Y.shape[0]))
            for d in np.nonzero(w)[0]:
                P *= 1 + w[d] * np.exp(-np.abs(X[:, d][:, None] - Y[:, d][None, :]))
            return P

        def kernel(self, X, Y, which_dim):
            return self.scale * self._prod(X, Y, which_dim)

        def individual_influence(self, X, Y, which_dim, which_dim_only):
            rest = which_dim * (1 - which_dim_only)
            return self.scale * (self._prod(X, Y, which_dim_only) - 1) * self._prod(X, Y, rest)

    return Cubic, Laplace


@pytest.mark.parametrize("clustered", [False, True])
def test_custom_mode_kernels_match_oracle(clustered):
    """Row f4: user ModeKernels (host Gram -> chd_regress, activations through individual_influence) against
    `oracle.chd.fit` running the same Python kernels through the reference's generic path (HELP:113-132)."""
    import computationalhypergraphdiscovery_b200 as CHD
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    Cubic, Laplace = _user_kernels()
    X, names = make_data(160, 7, seed=12)
    clusters = [names[0:2], names[2:3], names[3:5], names[5:6], names[6:7]] if clustered else None
    cub, lap = 0.5 * Cubic(), 1.5 * Laplace()
    gd = CHD.GraphDiscovery(X, names, kernels=[cub, lap], clusters=clusters, verbose=False)
    # without clusters every node prefers the Laplace kernel: force the cubic one there so both get chains
    gd.fit(kernel_chooser=None if clustered else _force_chooser(0))
    ocub, olap = 0.5 * Cubic(), 1.5 * Laplace()

    def lap_fn(A, B, w):
        return olap.kernel(A, B, w)

    lap_fn.individual_influence = olap.individual_influence
    res = ochd.fit(X, names, kernels=[("cubic", 0.5, ocub.kernel), ("laplace", 1.5, lap_fn)], clusters=clusters,
                   kernel_chooser=("min_noise",) if clustered else ("force", 0))
    assert set(res["chains"]) == set(gd.last_fit["chains"]) and len(res["chains"]) >= 1
    _compare_with_oracle(gd, res, names, {})
