"""End-to-end fit() wall time on the BASELINE.json config shapes that fit a single run (C1, C2, C3),
synthetic data of the named shapes (the reference's datasets are LFS pointers).  Prints one JSON line per
config; with --oracle also times the NumPy oracle on C1 and checks the graphs are identical."""
import argparse
import json
import os
import sys
import time

import networkx as nx
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import computationalhypergraphdiscovery_b200 as CHD  # noqa: E402
from computationalhypergraphdiscovery_b200 import _native  # noqa: E402
from computationalhypergraphdiscovery_b200.synthetic import make_data  # noqa: E402


def timed_fit(gd, **kw):
    torch.cuda.synchronize()
    l0 = _native.launch_count()
    t0 = time.perf_counter()
    gd.fit(**kw)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, _native.launch_count() - l0


def steps_of(gd):
    return int(sum(ch["noises"].shape[0] * (ch["noises"].shape[1] - 1) for ch in gd.last_fit["chains"].values()))


def c1(args):
    X, names = make_data(500, 11, seed=1)
    gd = CHD.GraphDiscovery(X, names, kernels=[CHD.Modes.LinearMode(), CHD.Modes.QuadraticMode()], verbose=False)
    timed_fit(gd)  # warm-up (context, allocations)
    gd = CHD.GraphDiscovery(X, names, kernels=[CHD.Modes.LinearMode(), CHD.Modes.QuadraticMode()], verbose=False)
    sec, launches = timed_fit(gd)
    out = {"config": "C1 Sachs-500 shape: n=500, N=11, Linear+Quadratic", "fit_seconds": sec,
           "node_prune_steps": steps_of(gd), "edges": gd.G.number_of_edges(), "gpu_launches": launches}
    if args.oracle:
        from oracle import chd as ochd
        t0 = time.perf_counter()
        res = ochd.fit(X, names, kernels=[("linear", 2.0, None), ("quadratic", 1.0, None)])
        out["oracle_cpu_seconds"] = time.perf_counter() - t0
        out["cpu_cores"] = os.cpu_count()
        same = all(sorted(a for a, _ in gd.G.in_edges(nm)) == sorted(res["nodes"].get(nm, {}).get("ancestors", []))
                   for nm in names)
        out["graph_identical_to_oracle"] = bool(same)
    return out


def c2(args):
    n, N = 7466, 11
    X, names = make_data(n, N, seed=2)
    rng = np.random.default_rng(3)
    pe = nx.DiGraph()
    pe.add_nodes_from(names)
    for a in names:
        for b in names:
            if a != b and rng.random() > 0.2:
                pe.add_edge(a, b)
    clusters = [names[0:3], names[3:6], names[6:8], names[8:11]]
    kernels = [0.1 * CHD.Modes.LinearMode(), 0.1 * CHD.Modes.QuadraticMode(), 0.1 * CHD.Modes.GaussianMode(l=1)]
    gd = CHD.GraphDiscovery(X, names, kernels=kernels, clusters=clusters, possible_edges=pe, verbose=False)
    sec, launches = timed_fit(gd)
    return {"config": "C2 Sachs-full shape: n=7466, N=11, 4 clusters, possible_edges, 0.1*(Linear,Quadratic,Gaussian)",
            "fit_seconds": sec, "node_prune_steps": steps_of(gd), "stage1_regressions": 33,
            "edges": gd.G.number_of_edges(), "gpu_launches": launches}


def c3(args):
    n, ns, b = 600, 1122, 32
    rng = np.random.default_rng(5)
    S = rng.random((n, ns))
    D = np.empty((n, b))
    for j in range(b):
        idx = rng.choice(ns, size=8, replace=False)
        D[:, j] = S[:, idx[0]] * S[:, idx[1]] - S[:, idx[2]] * S[:, idx[3]] + S[:, idx[4]] * S[:, idx[5]] \
            - 0.5 * S[:, idx[6]] + 0.01 * rng.standard_normal(n)
    X = np.concatenate([S, D], axis=1)
    names = [f"s{i}" for i in range(ns)] + [f"partial_{j}" for j in range(b)]
    pe = nx.DiGraph()
    edges = []
    for f in names[:ns]:
        for g in names[ns:]:
            edges.append((f, g))
            edges.append((g, f))
    pe.add_edges_from(edges)
    gd = CHD.GraphDiscovery(X, names, kernels=[CHD.Modes.QuadraticMode()], possible_edges=pe, verbose=False,
                            gamma_min=2e-6)
    sec, launches = timed_fit(gd, targets=names[ns:ns + args.c3_targets],
                              mode_chooser=CHD.decision.ThresholdModeChooser(threshold=0.025))
    return {"config": f"C3 Chemistry-Big shape: n=600, N={ns}+{b}, bipartite possible_edges, Quadratic, "
                      f"{args.c3_targets} targets, gamma_min=2e-6 (experiment.py:52-63)",
            "fit_seconds": sec, "node_prune_steps": steps_of(gd), "edges": gd.G.number_of_edges(),
            "gpu_launches": launches}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3")
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--c3-targets", type=int, default=32)
    args = ap.parse_args()
    for name in args.configs.split(","):
        out = {"c1": c1, "c2": c2, "c3": c3}[name](args)
        out["node_prune_steps_per_sec"] = out["node_prune_steps"] / out["fit_seconds"]
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
