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


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names = [f"v{i}" for i in range(7)]
    G = nx.DiGraph()
    G.add_nodes_from(names)
    mine = parallel.shard_targets(names, *parallel.dist_info())
    for nm in mine:  # stand-in for process_results on this rank's targets
        i = int(nm[1:])
        G.nodes[nm].update({"kernel_index": i % 3, "noise": 0.1 * i, "coeff": np.arange(3.0) + i})
        G.add_edge(names[(i + 1) % 7], nm, signal=float(i))
    merged = parallel.all_gather_results(parallel.export_node_results(G, mine))
    parallel.merge_node_results(G, merged)
    q.put((rank, sorted(G.edges(data="signal")), {n: G.nodes[n]["noise"] for n in names},
           float(sum(G.nodes[n]["coeff"].sum() for n in names))))
    dist.destroy_process_group()


def test_all_gather_results_world2():
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
    assert outs[0][1:] == outs[1][1:]
    edges, noises, csum = outs[0][1:]
    assert len(edges) == 7 and len(noises) == 7
    assert abs(csum - sum(3 * i + 3 for i in range(7))) < 1e-12


def test_compose_graphs():
    a, b = nx.DiGraph(), nx.DiGraph()
    a.add_edge("x", "y", signal=1.0)
    b.add_edge("z", "y", signal=2.0)
    g = parallel.compose_graphs([a, b])
    assert sorted(g.edges) == [("x", "y"), ("z", "y")]
