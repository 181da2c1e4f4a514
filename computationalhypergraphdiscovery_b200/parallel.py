"""Target sharding over GPUs: the path is independent per target node once X is known
(the reference's own scale-out is one process per GPU over target batches,
examples/Chemistry-Big/large_experiment_files/experiment_launcher.py:39-60), so ranks take disjoint
blocks of targets, X is replicated, and the only exchange is one all-gather of the per-node results
(nodes' attributes + in-edges) at the end of fit().  torch.distributed provides the plumbing
(NCCL on the GPUs, gloo in the CPU tests)."""
import networkx as nx


def dist_info():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:  # noqa: BLE001
        pass
    return 0, 1


def shard_targets(targets, rank, world):
    """Contiguous, balanced block of `targets` for `rank` (first `len % world` ranks get one more)."""
    targets = list(targets)
    base, rem = divmod(len(targets), world)
    start = rank * base + min(rank, rem)
    return targets[start:start + base + (1 if rank < rem else 0)]


def export_node_results(G, targets):
    """Per-target results written by process_results (MAIN:793-810): node attributes + in-edges."""
    out = {}
    for name in targets:
        out[name] = {"attrs": dict(G.nodes[name]),
                     "in_edges": [(a, dict(d)) for a, _b, d in G.in_edges(name, data=True)]}
    return out


def merge_node_results(G, results):
    for name, r in results.items():
        G.nodes[name].update(r["attrs"])
        for a, d in r["in_edges"]:
            G.add_edge(a, name, **d)
    return G


def all_gather_results(local):
    """All-gather of the per-node result dicts; returns the merged dict on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size()
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    merged = {}
    for part in gathered:
        merged.update(part)
    return merged


def all_gather_list(values):
    """All-gather of a small python list; returns the list of every rank's list."""
    import torch.distributed as dist
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, list(values))
    return gathered


def compose_graphs(graphs):
    """nx.compose merge of per-shard graphs (what the reference's Chemistry-Big notebook does with its
    per-batch pickles, `process big chemical reaction uniform data.ipynb:455-459`)."""
    out = nx.DiGraph()
    for g in graphs:
        out = nx.compose(out, g)
    return out
