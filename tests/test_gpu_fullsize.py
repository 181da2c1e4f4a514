"""Parity at BASELINE.json's full sample size (n = 8192) through size-independent properties (the oracle's
eigendecomposition is too slow there): orthogonal-similarity invariants of the tridiagonal form, the
normal equations of the regression, and consistency of the reported noise ratio."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

N_FULL, N_VARS = 8192, 64


@pytest.fixture(scope="module")
def setup():
    from computationalhypergraphdiscovery_b200 import _native
    from computationalhypergraphdiscovery_b200.synthetic import make_data
    X, _ = make_data(N_FULL, N_VARS, seed=0)
    X = np.ascontiguousarray((X - X.mean(0)) / X.std(0))
    plan = _native.Plan(N_FULL, N_VARS, N_VARS)
    Xd = torch.from_numpy(X).to(plan.device)
    descs = [_native.make_desc(0, 2.0), _native.make_desc(1, 1.0, alpha=1.0),
             _native.make_desc(2, 1.0, alpha=1.0, quad_scale=1.0, l=1.0)]
    return _native, plan, X, Xd, descs


def test_tridiagonal_form_keeps_trace_and_frobenius_norm(setup):
    nat, plan, X, Xd, descs = setup
    n, dev = N_FULL, plan.device
    w = torch.ones((1, N_VARS), dtype=torch.uint8, device=dev)
    w[0, 0] = 0
    ga = Xd[:, 0].contiguous().reshape(1, n)
    for desc in descs:
        K = plan.gram(desc, Xd, w)
        tr = torch.diagonal(K[0, :, :n]).sum().item()
        fro2 = (K[0, :, :n] ** 2).sum().item()
        d, e, b0 = plan.tridiag(K, ga)
        torch.cuda.synchronize()
        d, e = d[0], e[0, :n - 1]
        assert abs(d.sum().item() - tr) <= 1e-11 * abs(tr)
        assert abs((d ** 2).sum().item() + 2 * (e ** 2).sum().item() - fro2) <= 1e-11 * fro2
        assert abs(b0.item() ** 2 - (ga ** 2).sum().item()) <= 1e-12 * n


def test_regression_satisfies_normal_equations_at_full_size(setup):
    """yb = -(K + gamma_used I)^-1 ga  and  noise = gamma_used |yb|^2 / (-ga.yb)  (interpolatory.py:76-80)."""
    nat, plan, X, Xd, descs = setup
    n, dev = N_FULL, plan.device
    T = 2
    w = torch.ones((T, N_VARS), dtype=torch.uint8, device=dev)
    ga = torch.empty((T, n), dtype=torch.float64, device=dev)
    for t in range(T):
        w[t, t] = 0
        ga[t] = Xd[:, t]
    keys = torch.tensor([[1, 2], [3, 4]], dtype=torch.int32, device=dev)
    for desc in descs:
        K = plan.gram(desc, Xd, w)
        Kc = K.clone()
        gmin = torch.full((T,), 1e-9, dtype=torch.float64, device=dev)
        r = plan.regress(K, ga, keys, gmin, mode=1, base=1e-9)
        torch.cuda.synchronize()
        for t in range(T):
            g = r["gamma"][t].item()
            gu = max(g, r["gamma_min"][t].item())
            yb = r["yb"][t]
            res = Kc[t, :, :n] @ yb + gu * yb + ga[t]
            lam_max = torch.linalg.matrix_norm(Kc[t, :, :n], ord=1).item()   # bound of the spectral norm
            rel = (res.norm() / ga[t].norm()).item()
            assert rel <= 1e-13 * (lam_max + gu) / gu * 10 + 1e-12, (desc.kind, t, rel, gu)
            noise = gu * (yb @ yb).item() / (-(ga[t] @ yb).item())
            assert abs(noise - r["noise"][t].item()) <= 1e-9 * abs(noise)
            assert 0.0 <= r["z_low"][t].item() <= r["z_high"][t].item() <= 1.0 + 1e-9
            assert r["lambda_min"][t].item() >= -1e-12 * lam_max * n


def test_prune_step_at_full_size_is_deterministic_and_consistent(setup):
    nat, plan, X, Xd, descs = setup
    n, dev = N_FULL, plan.device
    T = 2
    cl = torch.arange(N_VARS, dtype=torch.int32, device=dev)
    outs = []
    for rep in range(2):
        w0 = torch.ones((T, N_VARS), dtype=torch.uint8, device=dev)
        ga = torch.empty((T, n), dtype=torch.float64, device=dev)
        for t in range(T):
            w0[t, t] = 0
            ga[t] = Xd[:, t]
        keys = torch.tensor([[1, 2], [3, 4]], dtype=torch.int32, device=dev)
        K = plan.gram(descs[1], Xd, w0)
        gmin = torch.full((T,), 1e-9, dtype=torch.float64, device=dev)
        r = plan.regress(K, ga, keys, gmin, mode=1, base=1e-9)
        alive = torch.ones((T, N_VARS), dtype=torch.uint8, device=dev)
        out = plan.prune_chain(descs[1], Xd, cl, w0, alive, ga, r["yb"], r["keys"], r["lambda_min"],
                               torch.full((T,), 1e-9, dtype=torch.float64, device=dev), 0, 2)
        torch.cuda.synchronize()
        outs.append({k: out[k].cpu().numpy().copy() for k in ("pruned", "noise", "gamma", "act")})
        assert (alive.sum(dim=1).cpu().numpy() == N_VARS - 2).all()
        pr = out["pruned"].cpu().numpy()
        act = out["act"].cpu().numpy()
        for t in range(T):
            assert pr[t, 0] == int(np.argmin(act[t, 0])) and pr[t, 1] == int(np.argmin(act[t, 1]))
            assert act[t, 0, t] == 2.0 and act[t, 1, pr[t, 0]] == 2.0     # empty rows carry the 2.0 sentinel
        nz = out["noise"].cpu().numpy()
        assert ((nz > 0) & (nz <= 1.0 + 1e-9)).all()
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k   # bitwise reproducible
