"""Target sharding over GPUs: the path is independent per target node once X is known
(the reference's own scale-out is one process per GPU over target batches,
examples/Chemistry-Big/large_experiment_files/experiment_launcher.py:39-60), so ranks take disjoint
blocks of targets, X is replicated, and the only exchange is one all-gather of the per-node results
(packed FP64 records: node attributes, edge signals, chains) at the end of fit().  torch.distributed provides the plumbing
(NCCL on the GPUs, gloo in the CPU tests)."""
import networkx as nx


_FORCE_SINGLE = False


class single_process:
    """Context manager: fit() ignores the process group inside it (every rank computes ALL targets); used by the
    multi-GPU equality check of bench.py."""

    def __enter__(self):
        global _FORCE_SINGLE
        self._old, _FORCE_SINGLE = _FORCE_SINGLE, True

    def __exit__(self, *exc):
        global _FORCE_SINGLE
        _FORCE_SINGLE = self._old


def dist_info():
    if _FORCE_SINGLE:
        return 0, 1
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


def _comm_device():
    """Tensors of a collective live where the backend wants them: the current CUDA device for NCCL, host for gloo."""
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def record_words(n, N, L):
    """Length (FP64 words) of one node's packed result record."""
    return 6 + (N + 31) // 32 + N + n + 4 * (L + 1) + L


def pack_node_results(G, targets, names, n, L, chains=None):
    """Per-node results of this rank as one (len(targets), record_words) FP64 array: the node attributes
    written by process_results (MAIN:793-802: kernel index, gamma, noise, active_modes as a bit set, coeff), the
    per-variable edge signals (MAIN:803-810), and the node's chain (noise / Z_low / Z_high / gamma per chain index,
    pruned cluster per step, chosen index).  Layout of a record:
      [0] present (1: node has ancestors, 0: none)  [1] kernel index  [2] gamma  [3] noise  [4] chosen chain index
      [5] chain steps S   [6 .. 6+W) active_modes bits, 32 per word   then signals (N), coeff (n),
      chain noise | Z_low | Z_high | gamma (4 x (L+1), NaN padded), pruned cluster per step (L, -1 padded).
    Integers up to 2^32 are exact in FP64, so one FP64 all-gather moves everything bit-exactly."""
    import numpy as np
    names = list(names)
    N = len(names)
    W = (N + 31) // 32
    idx = {nm: i for i, nm in enumerate(names)}
    rec = np.zeros((len(targets), record_words(n, N, L)))
    o_sig, o_cf = 6 + W, 6 + W + N
    o_ch = o_cf + n
    o_pr = o_ch + 4 * (L + 1)
    rec[:, o_ch:o_pr] = np.nan
    rec[:, o_pr:] = -1
    where = {}
    for k, ch in (chains or {}).items():
        for i, nm in enumerate(ch["targets"]):
            where[nm] = (ch, i)
    for t, name in enumerate(targets):
        data = G.nodes[name]
        if "kernel_index" not in data:
            continue
        am = np.asarray(data["active_modes"]) != 0
        bits = np.zeros(W * 32, dtype=np.uint64)
        bits[:N] = am
        words = (bits.reshape(W, 32) << np.arange(32, dtype=np.uint64)[None, :]).sum(axis=1)
        rec[t, 0], rec[t, 1], rec[t, 2], rec[t, 3] = 1, data["kernel_index"], data["gamma"], data["noise"]
        rec[t, 6:6 + W] = words.astype(np.float64)
        for a, _b, sig in G.in_edges(name, data="signal"):
            rec[t, o_sig + idx[a]] = sig
        rec[t, o_cf:o_cf + n] = data["coeff"]
        if name in where:
            ch, i = where[name]
            S1 = ch["noises"].shape[1]
            rec[t, 4], rec[t, 5] = ch["chosen_modes"][i], S1 - 1
            for q, key in enumerate(("noises", "Z_lows", "Z_highs", "gammas")):
                rec[t, o_ch + q * (L + 1):o_ch + q * (L + 1) + S1] = ch[key][i]
            if "pruned" in ch:
                rec[t, o_pr:o_pr + S1 - 1] = ch["pruned"][i]
    return rec


def unpack_node_results(G, targets, names, kernels, n, L, rec):
    """Inverse of pack_node_results: writes nodes / edges of G for `targets` (same attributes, same order of edge
    insertion as process_results) and returns {name: chain dict}."""
    import numpy as np
    names = list(names)
    N = len(names)
    W = (N + 31) // 32
    o_sig, o_cf = 6 + W, 6 + W + N
    o_ch = o_cf + n
    o_pr = o_ch + 4 * (L + 1)
    chains = {}
    for t, name in enumerate(targets):
        r = rec[t]
        if r[0] != 1:
            continue
        words = r[6:6 + W].astype(np.uint64)
        bits = ((words[:, None] >> np.arange(32, dtype=np.uint64)[None, :]) & np.uint64(1)).reshape(-1)[:N]
        k = int(r[1])
        G.nodes[name].update({"active_modes": bits.astype(np.float64), "kernel_index": k, "type": str(kernels[k]),
                              "gamma": float(r[2]), "noise": float(r[3]), "coeff": r[o_cf:o_cf + n].copy()})
        for i in np.nonzero(bits)[0]:
            G.add_edge(names[i], name, signal=float(r[o_sig + i]))
        S = int(r[5])
        chains[name] = dict(kernel_index=k, chosen_mode=int(r[4]),
                            **{key: r[o_ch + q * (L + 1):o_ch + q * (L + 1) + S + 1].copy()
                               for q, key in enumerate(("noises", "Z_lows", "Z_highs", "gammas"))},
                            pruned=r[o_pr:o_pr + S].astype(np.int64))
    return chains


def all_gather_records(rec, counts):
    """ONE all-gather (NCCL over NVLink on the GPUs: `all_gather_into_tensor`; gloo in the CPU tests) of every rank's
    packed records.  `rec` is this rank's (count, words) array, `counts` the number of targets of every rank (known
    from shard_targets).  Returns the (sum(counts), words) array in rank order, on every rank."""
    import numpy as np
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    cap, words = max(counts), rec.shape[1]
    dev = _comm_device()
    send = torch.zeros((cap, words), dtype=torch.float64, device=dev)
    if rec.shape[0]:
        send[:rec.shape[0]] = torch.from_numpy(np.ascontiguousarray(rec)).to(dev)
    recv = torch.empty((world, cap, words), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1))
    recv = recv.cpu().numpy()
    return np.concatenate([recv[r, :counts[r]] for r in range(world)], axis=0)


def all_gather_ints(values):
    """All-gather of a short list of ints (one int64 tensor per rank); returns a (world, len) array."""
    import torch
    import torch.distributed as dist
    dev = _comm_device()
    send = torch.tensor(list(values), dtype=torch.int64, device=dev)
    recv = torch.empty((dist.get_world_size(), send.numel()), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(recv.view(-1), send)
    return recv.cpu().numpy()


def compose_graphs(graphs):
    """nx.compose merge of per-shard graphs (what the reference's Chemistry-Big notebook does with its
    per-batch pickles, `process big chemical reaction uniform data.ipynb:455-459`)."""
    out = nx.DiGraph()
    for g in graphs:
        out = nx.compose(out, g)
    return out
