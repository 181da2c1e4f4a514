import os, sys
import numpy as np
import networkx as nx
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import computationalhypergraphdiscovery_b200 as CHD
from computationalhypergraphdiscovery_b200.synthetic import make_data
from oracle import chd as ochd
np.set_printoptions(precision=10, linewidth=200)
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
gd = CHD.GraphDiscovery(X, names, kernels=kernels, clusters=clusters, possible_edges=pe, verbose=False, gamma_min=1e-8)
gd.fit(kernel_chooser=CHD.decision.ThresholdKernelChooser(0.9), mode_chooser=CHD.decision.ThresholdModeChooser(0.5))
adj = np.array(nx.adjacency_matrix(pe, nodelist=names).todense())
res = ochd.fit(X, names, kernels=[("linear", 0.5, None), ("quadratic", 0.5, None), ("gaussian", 0.7, 1.5)],
               clusters=clusters, adjacency=adj, gamma_min=1e-8, kernel_chooser=("threshold", 0.9),
               mode_chooser=("threshold", 0.5))
for k, ch in res["chains"].items():
    mine = gd.last_fit["chains"][k]
    print("chain kernel", k, "targets", ch["targets"])
    print(" gammas oracle\n", ch["gammas"], "\n gammas mine\n", mine["gammas"])
    print(" noises oracle\n", ch["noises"], "\n noises mine\n", mine["noises"])
    print(" Zlow oracle\n", ch["Z_lows"], "\n mine\n", mine["Z_lows"])
