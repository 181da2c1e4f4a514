"""torchrun --nproc-per-node N tools/fit_multigpu.py : sharded fit() must equal the single-GPU fit()."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import computationalhypergraphdiscovery_b200 as CHD  # noqa: E402
from computationalhypergraphdiscovery_b200 import parallel  # noqa: E402
from computationalhypergraphdiscovery_b200.synthetic import make_data  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    X, names = make_data(400, 10, seed=2)
    gd = CHD.GraphDiscovery(X, names, verbose=False)
    gd.fit()
    edges = sorted((a, b, round(float(d["signal"]), 12)) for a, b, d in gd.G.edges(data=True))
    if rank == 0:
        real = parallel.dist_info
        parallel.dist_info = lambda: (0, 1)
        ref = CHD.GraphDiscovery(X, names, verbose=False)
        ref.fit()
        parallel.dist_info = real
        ref_edges = sorted((a, b, round(float(d["signal"]), 12)) for a, b, d in ref.G.edges(data=True))
        assert [e[:2] for e in edges] == [e[:2] for e in ref_edges], "edge sets differ"
        for e, r in zip(edges, ref_edges):
            assert abs(e[2] - r[2]) <= 1e-9 * max(1.0, abs(r[2]))
        for nm in names:
            a, b = gd.G.nodes[nm], ref.G.nodes[nm]
            assert set(a) == set(b)
            if "noise" in a:
                assert a["noise"] == b["noise"] and a["gamma"] == b["gamma"] and (a["coeff"] == b["coeff"]).all()
        print(f"sharded fit on {world} GPUs == single-GPU fit: {len(edges)} edges identical (bitwise node results)")
    gathered = [None] * world
    dist.all_gather_object(gathered, edges)
    assert all(g == gathered[0] for g in gathered), "graphs differ across ranks"
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
