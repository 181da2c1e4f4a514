"""N>1 host path on CPU: world_size-2 gloo run of the target sharding + result all-gather."""
import os
import socket

import networkx as nx
import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from computationalhypergraphdiscovery_b200 import parallel


def test_shard_targets_partition():
    t = [f"v{i}" for i in range(11)]
    for world in (1, 2, 3, 4, 8):
        parts = [parallel.shard_targets(t, r, world) for r in range(world)]
        assert sum(parts, []) == t
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


N_VARS, N_SAMPLES, L_STEPS = 7, 3, 5
KERNELS = ["linear", "quadratic", "gaussian"]


def _fake_fit(names, mine):
    """Stand-in for process_results on this rank's targets: node attributes, edges and chains of the right shapes."""
    G = nx.DiGraph()
    G.add_nodes_from(names)
    chains = {}
    for nm in mine:
        i = int(nm[1:])
        if i == 3:
            continue                                   # a node without ancestors (kernel choice -1)
        am = np.zeros(N_VARS)
        am[(i + 1) % N_VARS] = 1
        am[(i + 2) % N_VARS] = 1
        k = i % 3
        G.nodes[nm].update({"active_modes": am, "kernel_index": k, "type": KERNELS[k], "gamma": 0.5 + i,
                            "noise": 0.1 * i, "coeff": np.arange(3.0) + i})
        for j in np.nonzero(am)[0]:
            G.add_edge(names[j], nm, signal=float(10 * i + j))
        ch = chains.setdefault(k, dict(targets=[], noises=[], Z_lows=[], Z_highs=[], gammas=[], chosen_modes=[],
                                       pruned=[]))
        ch["targets"].append(nm)
        for key, base in (("noises", 0.0), ("Z_lows", 1.0), ("Z_highs", 2.0), ("gammas", 3.0)):
            ch[key].append(base + i + 0.01 * np.arange(L_STEPS + 1))
        ch["chosen_modes"].append(i % (L_STEPS + 1))
        ch["pruned"].append((np.arange(L_STEPS) + i) % N_VARS)
    for ch in chains.values():
        for key in ("noises", "Z_lows", "Z_highs", "gammas", "pruned"):
            ch[key] = np.stack(ch[key])
        ch["chosen_modes"] = np.array(ch["chosen_modes"])
    return G, chains


def _graph_digest(G, names):
    return (sorted(G.edges(data="signal")),
            {n: (G.nodes[n].get("kernel_index"), G.nodes[n].get("type"), G.nodes[n].get("gamma"),
                 G.nodes[n].get("noise"), None if "coeff" not in G.nodes[n] else G.nodes[n]["coeff"].tolist(),
                 None if "active_modes" not in G.nodes[n] else G.nodes[n]["active_modes"].tolist()) for n in names})


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names = [f"v{i}" for i in range(N_VARS)]
    mine = parallel.shard_targets(names, *parallel.dist_info())
    counts = [len(parallel.shard_targets(names, r, world)) for r in range(world)]
    G, chains = _fake_fit(names, mine)
    rec = parallel.pack_node_results(G, mine, names, N_SAMPLES, L_STEPS, chains)
    assert rec.shape == (len(mine), parallel.record_words(N_SAMPLES, N_VARS, L_STEPS))
    rec = parallel.all_gather_records(rec, counts)                     # the ONE collective of fit()
    for nm in mine:
        G.remove_edges_from(list(G.in_edges(nm)))
    got = parallel.unpack_node_results(G, names, names, KERNELS, N_SAMPLES, L_STEPS, rec)
    rows = parallel.all_gather_ints([rank + 1, 10 * rank]).tolist()
    q.put((rank, _graph_digest(G, names), {k: {a: np.asarray(b).tolist() for a, b in v.items()} for k, v in got.items()},
           rows))
    dist.destroy_process_group()


def test_packed_all_gather_world2_equals_unsharded():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert outs[0][1:] == outs[1][1:]                                  # every rank holds the same graph and chains
    names = [f"v{i}" for i in range(N_VARS)]
    G1, chains1 = _fake_fit(names, names)                              # the unsharded result
    assert outs[0][1] == _graph_digest(G1, names)                      # bit-identical attributes and signals
    got = outs[0][2]
    assert sorted(got) == sorted(n for n in names if n != "v3")
    for k, ch in chains1.items():
        for i, nm in enumerate(ch["targets"]):
            assert got[nm]["noises"] == ch["noises"][i].tolist() and got[nm]["gammas"] == ch["gammas"][i].tolist()
            assert got[nm]["pruned"] == ch["pruned"][i].tolist() and got[nm]["chosen_mode"] == ch["chosen_modes"][i]
            assert got[nm]["kernel_index"] == k
    assert outs[0][3] == [[1, 0], [2, 10]]


def test_pack_unpack_roundtrip_wide_mask():
    """Bit-set packing across several 32-bit words (N = 70)."""
    N, n, L = 70, 4, 3
    names = [f"x{i}" for i in range(N)]
    G = nx.DiGraph()
    G.add_nodes_from(names)
    am = np.zeros(N)
    am[[0, 31, 32, 63, 64, 69]] = 1
    G.nodes["x5"].update({"active_modes": am, "kernel_index": 1, "type": "b", "gamma": 2.5, "noise": 0.25,
                          "coeff": np.array([1.0, -2.0, 3.5, 1e-300])})
    for j in np.nonzero(am)[0]:
        G.add_edge(names[j], "x5", signal=-float(j) - 0.5)
    rec = parallel.pack_node_results(G, ["x5", "x6"], names, n, L)
    H = nx.DiGraph()
    H.add_nodes_from(names)
    parallel.unpack_node_results(H, ["x5", "x6"], names, ["a", "b"], n, L, rec)
    assert sorted(H.edges(data="signal")) == sorted(G.edges(data="signal"))
    assert (H.nodes["x5"]["active_modes"] == am).all() and (H.nodes["x5"]["coeff"] == G.nodes["x5"]["coeff"]).all()
    assert "kernel_index" not in H.nodes["x6"]


def test_compose_graphs():
    a, b = nx.DiGraph(), nx.DiGraph()
    a.add_edge("x", "y", signal=1.0)
    b.add_edge("z", "y", signal=2.0)
    g = parallel.compose_graphs([a, b])
    assert sorted(g.edges) == [("x", "y"), ("z", "y")]
