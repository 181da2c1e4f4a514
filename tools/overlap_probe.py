"""Does running the independent prune chains of a step on several CUDA streams overlap the latency-bound kernels
(panel QR, bulge chasing, W formation, band solve) of one chain with the GEMMs of another?  Times the C4 mixed step
(bench.py's workload) with the chains on 1 stream, on one stream per kernel group, and as half-batches.
    python tools/overlap_probe.py [--steps 3] [--targets 32] [--modes seq,k3,h6]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from computationalhypergraphdiscovery_b200 import _native  # noqa: E402
from computationalhypergraphdiscovery_b200 import random as chd_random  # noqa: E402
from computationalhypergraphdiscovery_b200.synthetic import make_data  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--N", type=int, default=512)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--targets", type=int, default=32)
ap.add_argument("--modes", default="seq,k3,h6")
ap.add_argument("--kernels", default="linear,quadratic,gaussian")
args = ap.parse_args()
n, N, Tb = args.n, args.N, args.targets
dev = torch.device("cuda", 0)
X, _ = make_data(n, N, seed=0)
X = np.ascontiguousarray((X - X.mean(0)) / X.std(0))
Xd = torch.from_numpy(X).to(dev)
cl = torch.arange(N, dtype=torch.int32, device=dev)
descs = {"linear": _native.make_desc(0, 2.0), "quadratic": _native.make_desc(1, 1.0, alpha=1.0),
         "gaussian": _native.make_desc(2, 1.0, alpha=1.0, quad_scale=1.0, l=1.0)}
keys_all = chd_random.split(chd_random.PRNGKey(0), N + 1)[1:]


def make_state(kernel, t0, T):
    plan = _native.Plan(n, N, N)
    tidx = (np.arange(T) + t0) % N
    w0 = np.ones((T, N), dtype=np.uint8)
    w0[np.arange(T), tidx] = 0
    w0d = torch.from_numpy(w0).to(dev)
    ga = torch.from_numpy(np.ascontiguousarray(X[:, tidx].T)).to(dev)
    keys = torch.from_numpy(keys_all[tidx].astype(np.uint32).view(np.int32).copy()).to(dev)
    gmin = torch.full((T,), 1e-9, dtype=torch.float64, device=dev)
    K = plan.gram(descs[kernel], Xd, w0d)
    r = plan.regress(K, ga, keys, gmin, mode=1, base=1e-9)
    torch.cuda.synchronize()
    del K
    pc = dict(P=torch.empty((T, n, plan.ldk), dtype=torch.float64, device=dev), valid=0) if kernel == "gaussian" else None
    return dict(plan=plan, desc=descs[kernel], w0=w0d, ga=ga, alive=torch.ones((T, N), dtype=torch.uint8, device=dev),
                yb=r["yb"], keys=r["keys"], lmin=r["lambda_min"],
                gmin=torch.full((T,), 1e-9, dtype=torch.float64, device=dev), step=0, p_cache=pc,
                stream=torch.cuda.Stream(device=dev))


def step(st):
    out = st["plan"].prune_chain(st["desc"], Xd, cl, st["w0"], st["alive"], st["ga"], st["yb"], st["keys"], st["lmin"],
                                 st["gmin"], st["step"], 1, keep_yb=True, p_cache=st["p_cache"])
    st["step"] += 1
    return out


def run(states, concurrent, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record()
    if concurrent:
        for st in states:
            st["stream"].wait_stream(main)
        for _ in range(steps):
            for st in states:
                with torch.cuda.stream(st["stream"]):
                    step(st)
        for st in states:
            main.wait_stream(st["stream"])
    else:
        for _ in range(steps):
            for st in states:
                step(st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


kernels = args.kernels.split(",")
res = {}
for mode in args.modes.split(","):
    if mode in ("seq", "k3"):
        states = [make_state(k, 0, Tb) for k in kernels]
    elif mode == "h6":
        states = [make_state(k, h * (Tb // 2), Tb // 2) for k in kernels for h in range(2)]
    elif mode == "q12":
        states = [make_state(k, h * (Tb // 4), Tb // 4) for k in kernels for h in range(4)]
    else:
        raise SystemExit(mode)
    run(states, mode != "seq", 2)          # warm-up
    ms = run(states, mode != "seq", args.steps)
    total = len(kernels) * Tb * args.steps
    res[mode] = dict(ms_per_step=ms / args.steps, node_prune_steps_per_sec=total / (ms * 1e-3), streams=len(states))
    print(mode, json.dumps(res[mode]), flush=True)
    del states
    torch.cuda.empty_cache()
print(json.dumps(res))
