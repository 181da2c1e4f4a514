"""Plotting pass-throughs (reference util.py: partition_layout UTIL:6-40, plot_noise_evolution
UTIL:158-245).  Out of the hot path; matplotlib is optional (absent in this image): without it the
functions return None instead of a figure so that `fit()` keeps working headless."""
import numpy as np
import networkx as nx

try:  # pragma: no cover - matplotlib is not installed in the build image
    import matplotlib.pyplot as plt
except Exception:  # noqa: BLE001
    plt = None


def have_matplotlib():
    return plt is not None


def _make_hypergraph(g, partition):
    communities = set(partition.values())
    hypergraph = nx.DiGraph()
    hypergraph.add_nodes_from(communities)
    hypergraph.add_nodes_from(g.nodes())
    for ni, nj, data in g.edges(data=True):
        ci, cj = partition[ni], partition[nj]
        if ci == cj:
            hypergraph.add_edge(ni, nj, intra_cluster=True, **data)
        else:
            hypergraph.add_edge(ni, cj, intra_cluster=False, **data)
    return hypergraph


def _position_communities(g, partition, **kwargs):
    between = {}
    for ni, nj in g.edges():
        ci, cj = partition[ni], partition[nj]
        if ci != cj:
            between[(ci, cj)] = between.get((ci, cj), 0) + 1
    hyper = nx.DiGraph()
    hyper.add_nodes_from(set(partition.values()))
    for (ci, cj), wgt in between.items():
        hyper.add_edge(ci, cj, weight=wgt)
    pos_communities = nx.spring_layout(hyper, seed=0, **kwargs)
    pos = {node: pos_communities[community] for node, community in partition.items()}
    return pos, pos_communities


def _position_nodes(g, partition, **kwargs):
    communities = {}
    for node, community in partition.items():
        communities.setdefault(community, []).append(node)
    pos = {}
    for _ci, nodes in communities.items():
        subgraph = g.subgraph(nodes)
        pos.update(nx.spring_layout(subgraph, seed=0, **kwargs))
    return pos


def partition_layout(g, partition, ratio=0.3):
    """Layout for a graph whose nodes are grouped in clusters; returns (pos, hypergraph)."""
    pos_nodes_communities, pos_communities = _position_communities(g, partition)
    assert set(pos_nodes_communities.keys()).isdisjoint(set(pos_communities.keys())), \
        "node names and community names must be different"
    pos_nodes = {k: ratio * v for k, v in _position_nodes(g, partition).items()}
    pos = {node: pos_nodes_communities[node] + pos_nodes[node] for node in g.nodes()}
    pos.update(pos_communities)
    return pos, _make_hypergraph(g, partition)


def plot_noise_evolution(ancestor_number, noises, Zs, node_name, ancestor_modes_number):
    """Noise-to-signal ratio along the prune chain + its increments; (None, None) without matplotlib."""
    if plt is None:
        return None, None
    fig, axes = plt.subplots(1, 2, figsize=(10, 4))
    noises = np.asarray(noises)
    axes[0].plot(ancestor_number, noises, marker="o", label="noise/(signal+noise)")
    axes[0].fill_between(ancestor_number, [z[0] for z in Zs], [z[1] for z in Zs], alpha=0.3, label="Z-test")
    axes[0].axvline(ancestor_modes_number, color="k", linestyle="--")
    axes[0].invert_xaxis()
    axes[0].set_xlabel("number of ancestors")
    axes[0].set_title(f"Noise evolution for {node_name}")
    axes[0].legend()
    inc = np.append(noises[1:] - noises[:-1], 1 - noises[-1])
    axes[1].plot(ancestor_number, inc, marker="o")
    axes[1].axvline(ancestor_modes_number, color="k", linestyle="--")
    axes[1].invert_xaxis()
    axes[1].set_xlabel("number of ancestors")
    axes[1].set_title("Noise increments")
    return fig, axes
