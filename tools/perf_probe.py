"""Times the building blocks of one node-prune-step with CUDA events (development probe)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from computationalhypergraphdiscovery_b200 import _native  # noqa: E402
from computationalhypergraphdiscovery_b200.synthetic import make_data  # noqa: E402


def ev_time(fn, reps=3):
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--N", type=int, default=16)
    ap.add_argument("--T", type=int, default=0)
    ap.add_argument("--G", type=int, default=0)
    ap.add_argument("--kinds", default="0,1,2")
    a = ap.parse_args()
    X, _ = make_data(a.n, a.N, 0)
    X = np.ascontiguousarray((X - X.mean(0)) / X.std(0))
    plan = _native.Plan(a.n, a.N, a.N, ctas_per_matrix=a.G)
    T = a.T or plan.matrices_per_wave
    dev = plan.device
    print(f"n={a.n} N={a.N} T={T} G={plan.ctas_per_matrix} wave={plan.matrices_per_wave} "
          f"ws/T={plan.workspace_bytes(1)/1e6:.1f}MB")
    Xd = torch.from_numpy(X).to(dev)
    w0 = np.ones((T, a.N), dtype=np.uint8)
    for t in range(T):
        w0[t, t % a.N] = 0
    w0d = torch.from_numpy(w0).to(dev)
    ga = Xd.t().contiguous()[torch.arange(T, device=dev) % a.N].contiguous()
    keys = torch.arange(2 * T, dtype=torch.int32, device=dev).reshape(T, 2)
    cl = torch.arange(a.N, dtype=torch.int32, device=dev)
    descs = {0: _native.make_desc(0, 2.0), 1: _native.make_desc(1, 1.0, alpha=1.0),
             2: _native.make_desc(2, 1.0, alpha=1.0, quad_scale=1.0, l=1.0)}
    for kind in [int(k) for k in a.kinds.split(",")]:
        desc = descs[kind]
        K = torch.zeros((T, a.n, plan.ldk), dtype=torch.float64, device=dev)
        t_gram = ev_time(lambda: plan.gram(desc, Xd, w0d, out=K))
        Kc = K.clone()
        B = torch.randn((T, a.n, 100), dtype=torch.float64, device=dev)
        t_copy = ev_time(lambda: Kc.copy_(K))
        t_tri = ev_time(lambda: (Kc.copy_(K), plan.tridiag(Kc, ga, B)), reps=2)
        gmin = torch.full((T,), 1e-9, dtype=torch.float64, device=dev)
        r = {}

        def reg():
            Kc.copy_(K)
            r.update(plan.regress(Kc, ga, keys, gmin, mode=1))
        t_reg = ev_time(reg, reps=2)
        alive = torch.ones((T, a.N), dtype=torch.uint8, device=dev)
        t_act = ev_time(lambda: plan.activations(desc, Xd, cl, w0d, alive, r["yb"], ga, False))
        st = r["status"].cpu().numpy()
        print(f"kind={kind}: gram {t_gram:.2f} ms | tridiag {t_tri - t_copy:.2f} ms | regress total "
              f"{t_reg - t_copy:.2f} ms | activations {t_act:.2f} ms | per target-step "
              f"{(t_gram + t_reg - t_copy + t_act) / T:.2f} ms | status {np.bincount(st, minlength=8).tolist()} "
              f"gamma {r['gamma'][:3].cpu().numpy()}", flush=True)


if __name__ == "__main__":
    main()
