"""GraphDiscovery: drop-in host side of the CHD hot path, backed by libchd_b200 (CUDA, sm_100a).

Mirrors the public API and semantics of the reference's GraphDiscovery
(src/ComputationalHypergraphDiscovery/_GraphDiscoveryMain.py, abbreviated MAIN below):
  ctor / validation / normalisation      MAIN:57-134
  pickling                               MAIN:136-158
  kernels property                       MAIN:160-168
  from_dataframe                         MAIN:170-184
  prepare_new_graph_with_clusters        MAIN:186-223
  prepare_functions                      MAIN:225-270  (here: kernel descriptors + device plan)
  fit                                    MAIN:272-421
  _prepare_modes_for_ancestor_finding    MAIN:423-446, helper_functions.py:9-37
  get_kernel_performance (stage 1)       MAIN:448-511
  choose_kernel                          MAIN:513-557
  prune_ancestors (stage 2)              MAIN:559-701
  process_results                        MAIN:703-817
  plot_graph                             MAIN:819-932
  predict                                MAIN:934-979
  show_functional_dependencies           MAIN:981-1002
The numeric work (Gram assembly, regression + gamma search + Z-test, activations, the prune
chain) is done by the C-ABI calls in `_native.Plan`; this file only orchestrates, decides and
writes the networkx graph.  No CPU fallback exists.
"""
import numpy as np
import networkx as nx
import torch

from . import _native
from . import random as chd_random
from . import parallel
from .Modes import ModeContainer, LinearMode, QuadraticMode, GaussianMode
from .decision import *  # noqa: F401,F403  (reference does the same, MAIN:14)
from .decision import MinNoiseKernelChooser, MaxIncrementModeChooser
from .util import partition_layout, plot_noise_evolution, have_matplotlib

_DEVICE_ATTRS = ["_plan", "_Xd", "_cluster_of_d"]


class GraphDiscovery:
    """Discovers the hypergraph structure of a dataset (see the reference's class docstring,
    MAIN:20-45).  `X` is (n_samples, n_features); results are stored in the networkx DiGraph `G`."""

    # kept for API compatibility (MAIN:48-55): the attributes dropped when pickling
    UNPICKLABLE_ATTRS = list(_DEVICE_ATTRS)

    def __init__(self, X, names, kernels=None, normalize=True, clusters=None, possible_edges=None, verbose=True,
                 gamma_min=1e-9, is_interpolatory: bool = None) -> None:
        X = np.asarray(X, dtype=np.float64)
        assert X.shape[1] == len(names), "X must have as many columns as there are names"
        assert len(X.shape) == 2, "X must be a 2D array"
        assert not np.isnan(np.sum(X)), "X must not contain NaN values"
        assert not np.isinf(np.sum(X)), "X must not contain infinite values"
        standard_devs = X.std(axis=0, keepdims=True)
        if np.any(standard_devs == 0):
            raise ValueError(
                "Some features have a standard deviation of 0. This can lead to numerical instability. "
                "Please remove these features.")
        if normalize:
            self.mean_x = X.mean(axis=0, keepdims=True)
            self.std_x = standard_devs
            self.X = (X - self.mean_x) / self.std_x
        else:
            self.mean_x = np.zeros((1, X.shape[1]))
            self.std_x = np.ones((1, X.shape[1]))
            self.X = X
        self.normalize = normalize
        self.verbose = verbose
        self.print_func = print if verbose else (lambda *a, **k: None)
        self.names = names
        self.name_to_index = {name: index for index, name in enumerate(names)}
        if possible_edges is not None:
            possible_edges.remove_edges_from(nx.selfloop_edges(possible_edges))
            self.print_func("Converting possible edges to dense adjacency matrix")
            self.possible_edges_adjacency = np.array(
                nx.adjacency_matrix(possible_edges, nodelist=list(self.names)).todense())
        self.possible_edges = possible_edges
        self.G = nx.DiGraph()
        self.G.add_nodes_from(names)

        self.has_clusters = clusters is not None
        self.modes = ModeContainer(self.names, clusters)
        if kernels is None:
            kernels = [LinearMode(), QuadraticMode(), GaussianMode(l=1)]
        else:
            assert len(kernels) == len(set([str(k) for k in kernels])), "Kernels must have different names"
        self._kernels = kernels
        self.gamma_min = gamma_min
        self.is_interpolatory = is_interpolatory   # no-op in the reference too (identical twin files)
        self.last_fit = None
        self.prepare_functions()

    # ------------------------------------------------------------------ pickling (MAIN:136-158)
    def __getstate__(self):
        state = self.__dict__.copy()
        for attr in _DEVICE_ATTRS + ["print_func", "last_fit"]:
            state.pop(attr, None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.print_func = print if self.__dict__.get("verbose", True) else (lambda *a, **k: None)
        self.last_fit = None
        self.prepare_functions()

    @property
    def kernels(self):
        return self._kernels

    @kernels.setter
    def kernels(self, *args, **kwargs):
        raise ValueError(
            "Kernels cannot be changed after initialization. Please create a new GraphDiscovery object with "
            "the desired kernels.")

    def from_dataframe(df, **kwargs):
        """Alternative constructor taking a pandas DataFrame (MAIN:170-184)."""
        X = df.values
        return GraphDiscovery(X=X, names=df.columns, **kwargs)

    def prepare_new_graph_with_clusters(self, clusters, kernels=None):
        """New GraphDiscovery with the given clusters, inheriting this graph minus the edges that
        point from a cluster to a later one (MAIN:186-223; quirks kept: data is normalised again,
        verbose=True)."""
        new_graph = GraphDiscovery(X=self.X, names=self.names, kernels=self.kernels if kernels is None else kernels,
                                   clusters=clusters, possible_edges=self.possible_edges, verbose=True,
                                   normalize=self.normalize)
        new_graph.print_func = self.print_func
        new_graph.G = self.G.copy()
        edges_to_remove = []
        flattened_clusters = [(i, item) for i, sublist in enumerate(clusters) for item in sublist]
        for i, (cluster_index, node) in enumerate(flattened_clusters):
            other_nodes = list(set([node for j, node in flattened_clusters if j > i]) - set(clusters[cluster_index]))
            for other_node in other_nodes:
                edges_to_remove.append((node, other_node))
        new_graph.G.remove_edges_from(edges_to_remove)
        return new_graph

    # ------------------------------------------------------------------ device side
    def prepare_functions(self):
        """Resolve kernel scales into device descriptors (MAIN:235-237).  The device plan and the
        resident copy of X are created lazily on first use so that construction, pickling and the
        host logic work without a GPU."""
        scales = {k.name: k.scale for k in self.kernels}
        for kernel in self.kernels:
            kernel.setup(self.X, scales=scales)
        self._plan = None
        self._Xd = None
        self._cluster_of_d = None

    def _device(self):
        if self._plan is None:
            n, N = self.X.shape
            self._plan = _native.Plan(n, N, len(self.modes.clusters))
            dev = self._plan.device
            self._Xd = torch.from_numpy(np.ascontiguousarray(self.X)).to(dev)
            self._cluster_of_d = torch.from_numpy(self.modes.cluster_of).to(dev)
        return self._plan

    def _batch_size(self, T, chain_steps=0, gaussian=False):
        """Targets per launch: one in-flight batch must fit in 70 % of the free device memory -- the library's
        workspace, the Gram output (stage 1) or the Gaussian product-map cache (chain), and the per-step chain
        outputs (yb of every step is kept, like the reference's ybs, MAIN:651-657)."""
        plan = self._device()
        free, _total = torch.cuda.mem_get_info(plan.device)
        held = (plan._ws.numel() if plan._ws is not None else 0) + getattr(self, "_held_bytes", 0)
        per = plan.workspace_bytes(1) + 8 * plan.n * plan.ldk * (2 if gaussian or not chain_steps else 1)
        per += 8 * (plan.n + len(self.modes.clusters) + 8) * max(int(chain_steps), 0)
        forced = getattr(self, "_force_batch_size", None)        # test hook: exercise the multi-batch path
        if forced:
            return int(max(1, min(T, forced)))
        return int(max(1, min(T, (0.7 * (free + held)) // per)))

    # ------------------------------------------------------------------ fit
    def fit(self, targets=None, key=None, kernel_chooser=None, mode_chooser=None, message=""):
        """Performs graph discovery for `targets` (default: every node).  `key` is a JAX-style
        threefry key (2 x uint32; default PRNGKey(0) as in the reference, MAIN:275)."""
        if kernel_chooser is None:
            kernel_chooser = MinNoiseKernelChooser()
        else:
            kernel_chooser.is_a_kernel_chooser()
        if mode_chooser is None:
            mode_chooser = MaxIncrementModeChooser()
        else:
            mode_chooser.is_a_mode_chooser()
        key = chd_random.as_key(key)
        if targets is None:
            targets = self.names
        all_targets = list(targets)
        subkeys_all = chd_random.split(key, len(all_targets) + 1)[1:]   # MAIN:311-313
        # multi-GPU: ranks take disjoint blocks of targets (parallel.py); every target keeps the subkey it
        # has in the unsharded run, so results do not depend on the number of GPUs
        rank, world = parallel.dist_info()
        targets = parallel.shard_targets(all_targets, rank, world)
        pos = {name: i for i, name in enumerate(all_targets)}
        subkeys = subkeys_all[[pos[t] for t in targets]] if targets else subkeys_all[:0]
        counts = [len(parallel.shard_targets(all_targets, r, world)) for r in range(world)]
        n_s, L = self.X.shape[0], max(len(self.modes.clusters) - 2, 0)
        if world > 1 and not targets:
            # more ranks than targets: nothing to compute here, but take part in the two collectives
            parallel.all_gather_ints([-1] * len(self.kernels))
            rec = parallel.all_gather_records(np.zeros((0, parallel.record_words(n_s, len(self.names), L))), counts)
            self.last_fit = dict(targets=[], gathered=parallel.unpack_node_results(
                self.G, all_targets, self.names, self.kernels, n_s, L, rec))
            return
        gas_idx, w0 = self._prepare_modes_for_ancestor_finding(targets)

        kernel_performance, stage1_dev = self.get_kernel_performance(w0, gas_idx, subkeys)
        ybs, gammas, subkeys, noisess, Z_lows, Z_highs, chosen_kernels = GraphDiscovery.choose_kernel(
            kernel_performance, kernel_chooser)

        self.last_fit = dict(targets=targets, stage1=kernel_performance, chosen_kernel=chosen_kernels, chains={})
        # the chain length is a property of the kernel group over ALL targets (quirk C-16, MAIN:659-660)
        im = self.modes.index_matrix
        rows0 = ((im[None, :, :] * w0[:, None, :]).sum(axis=2) > 0).sum(axis=1) if len(targets) else np.zeros(0, int)
        group_rows = [int(rows0[chosen_kernels == i].max()) if np.any(chosen_kernels == i) else -1
                      for i in range(len(self.kernels))]
        if world > 1:
            group_rows = [int(v) for v in parallel.all_gather_ints(group_rows).max(axis=0)]
        for i, kernel in enumerate(self.kernels):
            mask_kernel = chosen_kernels == i
            if not np.any(mask_kernel):
                continue
            (active_modess_kernel, noisess_kernel, Z_lows_kernel, Z_highs_kernel, activationss_kernel, gammas_kernel,
             ybs_kernel) = self.prune_ancestors(
                kernel, X=self.X, gas_idx=gas_idx[mask_kernel], w0=w0[mask_kernel], ybs_kernel=ybs[mask_kernel],
                gammas_kernel=gammas[mask_kernel], subkeys_kernel=subkeys[mask_kernel],
                noise_kernel=noisess[mask_kernel], Z_low_kernel=Z_lows[mask_kernel],
                Z_high_kernel=Z_highs[mask_kernel], lambda_min_kernel=stage1_dev["lambda_min"][mask_kernel, i],
                loop_number=len(self.modes.clusters) - 2, message=message, group_max_rows=group_rows[i])
            raw_noises = noisess_kernel
            noisess_kernel, Z_lows_kernel, Z_highs_kernel = self._anomaly_clamp(
                noisess_kernel, Z_lows_kernel, Z_highs_kernel)
            chosen_modes = np.array([
                int(mode_chooser(active_modess_kernel[t], noisess_kernel[t], (Z_lows_kernel[t], Z_highs_kernel[t])))
                for t in range(noisess_kernel.shape[0])])
            self.last_fit["chains"][i] = dict(
                targets=[t for t, m in zip(targets, mask_kernel) if m], active=active_modess_kernel,
                noises=noisess_kernel, raw_noises=raw_noises, Z_lows=Z_lows_kernel, Z_highs=Z_highs_kernel,
                activations=activationss_kernel, gammas=gammas_kernel, chosen_modes=chosen_modes,
                pruned=self._chain_pruned)
            self.process_results(targets, i, mask_kernel, chosen_modes, active_modess_kernel, noisess_kernel,
                                 Z_lows_kernel, Z_highs_kernel, activationss_kernel, kernel_performance,
                                 gammas_kernel, ybs_kernel)
        self.process_results(targets, -1, chosen_kernels == -1, *([None] * 6), kernel_performance, None, None)
        if world > 1:
            # ONE all-gather of the packed per-node results; every rank then rebuilds the nodes and
            # edges of ALL targets from the records, in target order, so G does not depend on the number of GPUs
            rec = parallel.pack_node_results(self.G, targets, self.names, n_s, L, self.last_fit["chains"])
            rec = parallel.all_gather_records(rec, counts)
            for name in targets:                     # drop the locally written copies: rebuilt below, in order
                self.G.remove_edges_from(list(self.G.in_edges(name)))
            self.last_fit["gathered"] = parallel.unpack_node_results(
                self.G, all_targets, self.names, self.kernels, n_s, L, rec)

    @staticmethod
    def _anomaly_clamp(noises, zl, zh):
        """MAIN:349-392."""
        def anom(a):
            return np.logical_or(a > 1.05, a < -0.05)

        an, al, ah = anom(noises), anom(zl), anom(zh)
        if np.any(an) or np.any(al) or np.any(ah):
            print("-" * 50 + "\n** Anomaly in noise or Z values, consider increasing gamma_min **\n" + "-" * 50)

            def clamp(a, mask):
                first = np.argmax(mask, axis=1)[:, None]
                anyrow = np.any(mask, axis=1)[:, None]
                return np.where(np.logical_and(np.arange(a.shape[1])[None, :] >= first, anyrow), 1.0, a)

            noises, zl, zh = clamp(noises, an), clamp(zl, al), clamp(zh, ah)
        return noises, zl, zh

    def _prepare_modes_for_ancestor_finding(self, names):
        """Target column indices and the initial variable masks (MAIN:423-446, HELP:15-29): the
        target's own column is removed, and with possible_edges every column j with adj[j, idx] == 0."""
        indexes = np.array([self.modes.name_to_index[name] for name in names])
        N = self.X.shape[1]
        w0 = np.ones((len(names), N), dtype=np.uint8)
        for t, idx in enumerate(indexes):
            if self.possible_edges is not None:
                w0[t, self.possible_edges_adjacency[:, idx] == 0] = 0
            w0[t, idx] = 0
        return indexes, w0

    def _active_modes(self, w, alive=None):
        """(T, C, N) 0/1 active-mode matrices from variable masks (and alive rows)."""
        im = self.modes.index_matrix
        am = im[None, :, :] * w[:, None, :]
        if alive is not None:
            am = am * alive[:, :, None]
        return am

    # ------------------------------------------------------------------ stage 1
    def get_kernel_performance(self, w0, gas_idx, subkeys):
        """Regression of every target with every kernel (MAIN:448-511).  Returns the reference's
        6-tuple of (T, nk, ...) host arrays and a dict of extra per-(target, kernel) device results."""
        plan = self._device()
        dev = plan.device
        T, nk, n = len(gas_idx), len(self.kernels), plan.n
        XT = self._Xd.t().contiguous()
        ga_all = XT[torch.from_numpy(gas_idx).to(dev)]                    # (T, n)
        ybs = torch.empty((T, nk, n), dtype=torch.float64, device=dev)
        outs = {k: torch.empty((T, nk), dtype=torch.float64, device=dev)
                for k in ("noise", "z_low", "z_high", "gamma", "lambda_min")}
        keys_out = torch.empty((T, nk, 2), dtype=torch.int64, device=dev)
        w0_d = torch.from_numpy(w0).to(dev)
        keys_d = torch.from_numpy(subkeys.astype(np.uint32).view(np.int32).copy()).to(dev)
        bs = self._batch_size(T)
        for ki, kernel in enumerate(self.kernels):
            for s in range(0, T, bs):
                sl = slice(s, min(T, s + bs))
                K = self._gram(kernel, w0_d[sl].contiguous(), w0[sl])
                gmin = torch.full((sl.stop - sl.start,), float(self.gamma_min), dtype=torch.float64, device=dev)
                r = plan.regress(K, ga_all[sl].contiguous(), keys_d[sl].contiguous(), gmin, mode=1,
                                 base=float(self.gamma_min))
                ybs[sl, ki] = r["yb"]
                for name in outs:
                    outs[name][sl, ki] = r[name]
                keys_out[sl, ki] = r["keys"].to(torch.int64) & 0xFFFFFFFF
                del K
        noises = outs["noise"].cpu().numpy()
        gam = outs["gamma"].cpu().numpy()
        if np.isnan(gam).any():                                           # MAIN:493-498
            print("Somme gammas were found to be NaNs. This error may not be corrected, so corresponding "
                  "signal-to-noise ratio are set to 1")
            noises = np.where(np.isnan(gam), 1, noises)
        if np.isnan(noises).any():                                        # MAIN:499-505
            raise ValueError(
                "The regression has returned NaNs, this is likely due to the kernel matrix not being positive "
                "definite\nConsider increasing gamma_min\n"
                f"smallest eigenvalues: {outs['lambda_min'].cpu().numpy()[np.isnan(noises)]}")
        perf = (ybs.cpu().numpy(), noises, outs["z_low"].cpu().numpy(), outs["z_high"].cpu().numpy(), gam,
                keys_out.cpu().numpy().astype(np.uint32))
        return perf, dict(lambda_min=outs["lambda_min"].cpu().numpy())

    # ------------------------------------------------------------------ user-defined kernels (generic path)
    def _gram(self, kernel, w_dev, w_host):
        """(T, n, ldk) Gram matrices: CUDA for the built-in modes, host-provided for user ModeKernels."""
        plan = self._device()
        if kernel.desc is not None:
            return plan.gram(kernel.desc, self._Xd, w_dev)
        T, n = w_host.shape[0], plan.n
        K = torch.zeros((T, n, plan.ldk), dtype=torch.float64, device=plan.device)
        for t in range(T):
            Kt = np.asarray(kernel(self.X, self.X, w_host[t].astype(np.float64)), dtype=np.float64)
            K[t, :, :n] = torch.from_numpy(np.ascontiguousarray(Kt)).to(plan.device)
        return K

    def _custom_activations(self, kernel, yb, ga, w_vars, alive):
        """HELP:113-132 for one target on the host: yb^T Infl_c yb / energy, 2.0 for empty rows."""
        im = self.modes.index_matrix
        energy = -float(np.dot(ga, yb))
        act = np.full(im.shape[0], 2.0)
        w = w_vars.astype(np.float64)
        for c in range(im.shape[0]):
            w_c = im[c] * w * float(alive[c])
            if not w_c.any():
                continue
            M = np.asarray(kernel.individual_influence(self.X, self.X, w, w_c))
            act[c] = float(yb @ (M @ yb)) / energy
        return act

    def _custom_chain(self, kernel, b, w0_host, first_step, nsteps):
        """Prune steps for a user ModeKernel: host-orchestrated (activations + Gram on the host,
        regression / gamma search / Z-test in chd_regress).  Same outputs as Plan.prune_chain."""
        plan = self._device()
        dev, f64 = plan.device, torch.float64
        tb, n, N, C = w0_host.shape[0], plan.n, plan.N, len(self.modes.clusters)
        cluster_of = self.modes.cluster_of
        out = dict(pruned=torch.empty((tb, nsteps), dtype=torch.int32, device=dev),
                   act=torch.empty((tb, nsteps, C), dtype=f64, device=dev),
                   yb=torch.empty((tb, nsteps, n), dtype=f64, device=dev))
        for k in ("noise", "z_low", "z_high", "gamma"):
            out[k] = torch.empty((tb, nsteps), dtype=f64, device=dev)
        ga_h = b["ga"].cpu().numpy()
        for s in range(nsteps):
            step = first_step + s
            p = N - step - 1
            if step % 50 == 0 or (p * (p + 1)) / 2 == N:                   # MAIN:622-631
                lm = b["lmin"]
                b["gmin"].copy_(torch.where(lm + 1e-9 < 0, -2 * lm, torch.full_like(lm, 1e-9)))
            yb_h, alive_h = b["yb"].cpu().numpy(), b["alive"].cpu().numpy()
            acts = np.empty((tb, C))
            pr = np.empty(tb, dtype=np.int32)
            for t in range(tb):
                w_vars = w0_host[t] * alive_h[t][cluster_of]
                acts[t] = self._custom_activations(kernel, yb_h[t], ga_h[t], w_vars, alive_h[t])
                nan = np.isnan(acts[t])
                pr[t] = int(np.argmax(nan)) if nan.any() else int(np.argmin(acts[t]))
                alive_h[t, pr[t]] = 0
            b["alive"].copy_(torch.from_numpy(alive_h).to(dev))
            w_new = (w0_host * alive_h[:, cluster_of]).astype(np.uint8)
            K = self._gram(kernel, None, w_new)
            r = plan.regress(K, b["ga"], b["keys"], b["gmin"], mode=0)
            b["yb"].copy_(r["yb"])
            b["keys"].copy_(r["keys"])
            b["lmin"].copy_(r["lambda_min"])
            out["pruned"][:, s] = torch.from_numpy(pr).to(dev)
            out["act"][:, s] = torch.from_numpy(acts).to(dev)
            out["yb"][:, s] = r["yb"]
            for k in ("noise", "z_low", "z_high", "gamma"):
                out[k][:, s] = r[k]
        return out

    def choose_kernel(kernel_performances, kernel_chooser):
        """Per-node kernel choice (MAIN:513-557)."""
        ybs, gammas, subkeys, noisess, Z_lows, Z_highs, chosen_kernels = [], [], [], [], [], [], []
        for i in range(kernel_performances[0].shape[0]):
            which_kernel, yb, noise, Z_low, Z_high, gamma, skey = kernel_chooser(
                [kernel_performances[j][i] for j in range(6)])
            chosen_kernels.append(int(which_kernel))
            ybs.append(yb)
            noisess.append(noise)
            Z_lows.append(Z_low)
            Z_highs.append(Z_high)
            gammas.append(gamma)
            subkeys.append(skey)
        return (np.array(ybs), np.array(gammas), np.array(subkeys), np.array(noisess), np.array(Z_lows),
                np.array(Z_highs), np.array(chosen_kernels))

    # ------------------------------------------------------------------ stage 2
    def prune_ancestors(self, kernel, X, gas_idx, w0, ybs_kernel, gammas_kernel, subkeys_kernel, noise_kernel,
                        Z_low_kernel, Z_high_kernel, lambda_min_kernel, loop_number, message,
                        group_max_rows=None):
        """Backward elimination for the targets that chose `kernel` (MAIN:559-701).  All targets of
        the group step in lock-step (quirk: the loop only stops when every mask is empty,
        MAIN:659-660); the steps run device-resident in chd_prune_chain.  Returns the reference's
        7 stacked arrays (host), except `ybs` which stays on the device: (T_k, L+1, n) tensor."""
        plan = self._device()
        dev = plan.device
        Tk, n, N = len(gas_idx), plan.n, plan.N
        C = len(self.modes.clusters)
        im = self.modes.index_matrix
        XT = self._Xd.t().contiguous()
        ga = XT[torch.from_numpy(gas_idx).to(dev)].contiguous()
        # number of non-empty rows per target decides when every mask is empty
        nonempty0 = ((im[None, :, :] * w0[:, None, :]).sum(axis=2) > 0)
        max_rows = int(nonempty0.sum(axis=1).max()) if Tk else 0
        if group_max_rows is not None:
            max_rows = max(max_rows, int(group_max_rows))
        # the reference checks `all masks empty` AFTER a step (MAIN:659-660): at least one step runs whenever
        # loop_number > 0, even if every initial mask of the group is already empty
        steps_planned = min(loop_number, max(1, max_rows)) if loop_number > 0 else 0
        self.print_func(f"Finding ancestors with kernel [{kernel}]: {steps_planned} steps x {Tk} targets {message}")

        gaussian = kernel.desc is not None and kernel.desc.kind == _native.KERNEL_GAUSSIAN
        bs = self._batch_size(Tk, chain_steps=steps_planned, gaussian=gaussian)
        # Chain outputs of the whole group; every batch writes its slice as soon as its launch is enqueued, so only
        # one batch's scratch (workspace, product-map cache, per-step outputs) is alive at a time.
        f64 = torch.float64
        cap = steps_planned                                   # grows only in the degenerate extra-step case below
        chain = dict(noise=torch.empty((Tk, cap), dtype=f64, device=dev), z_low=torch.empty((Tk, cap), dtype=f64, device=dev),
                     z_high=torch.empty((Tk, cap), dtype=f64, device=dev), gamma=torch.empty((Tk, cap), dtype=f64, device=dev),
                     pruned=torch.zeros((Tk, cap), dtype=torch.int32, device=dev),
                     act=torch.empty((Tk, max(cap, 1), C), dtype=f64, device=dev))
        yb_chain = torch.empty((Tk, cap + 1, n), dtype=f64, device=dev)
        yb_chain[:, 0] = torch.from_numpy(np.ascontiguousarray(ybs_kernel)).to(dev)
        self._held_bytes = sum(v.numel() * v.element_size() for v in chain.values()) + yb_chain.numel() * 8
        # Gaussian mode: product map prod(1 + E_d) carried from step to step (one division per prune); ONE buffer of
        # the batch size, reused by the batches (they run one after another) and invalidated in between
        p_buf = torch.empty((min(bs, Tk), n, plan.ldk), dtype=f64, device=dev) if gaussian and steps_planned else None

        def grow(extra):
            nonlocal yb_chain, cap
            for k, v in chain.items():
                if k == "act" and cap == 0:
                    continue
                pad = torch.empty(v.shape[:1] + (extra,) + v.shape[2:], dtype=v.dtype, device=dev)
                chain[k] = torch.cat([v, pad], dim=1)
            yb_chain = torch.cat([yb_chain, torch.empty((Tk, extra, n), dtype=f64, device=dev)], dim=1)
            cap += extra

        batches = []
        for s in range(0, Tk, bs):
            sl = slice(s, min(Tk, s + bs))
            tb = sl.stop - sl.start
            batches.append(dict(
                sl=sl, w0=torch.from_numpy(np.ascontiguousarray(w0[sl])).to(dev),
                alive=torch.ones((tb, C), dtype=torch.uint8, device=dev),
                yb=torch.from_numpy(np.ascontiguousarray(ybs_kernel[sl])).to(dev),
                keys=torch.from_numpy(subkeys_kernel[sl].astype(np.uint32).view(np.int32).copy()).to(dev),
                lmin=torch.from_numpy(np.ascontiguousarray(lambda_min_kernel[sl])).to(dev),
                gmin=torch.full((tb,), 1e-9, dtype=torch.float64, device=dev), ga=ga[sl].contiguous()))
        done = 0
        if loop_number == 0:
            # MAIN:662-691: no prune step, but the activations of the full set are still needed
            for b in batches:
                if kernel.desc is None:
                    ybh, gah, alh = b["yb"].cpu().numpy(), b["ga"].cpu().numpy(), b["alive"].cpu().numpy()
                    w0b = w0[b["sl"]]
                    act0 = torch.from_numpy(np.stack([self._custom_activations(kernel, ybh[t], gah[t], w0b[t], alh[t])
                                                      for t in range(w0b.shape[0])])).to(dev)
                else:
                    act0, _ = plan.activations(kernel.desc, self._Xd, self._cluster_of_d, b["w0"], b["alive"], b["yb"],
                                               b["ga"], False)
                chain["act"][b["sl"], 0] = act0
        else:
            def run(first, nsteps):
                for b in batches:
                    sl = b["sl"]
                    if kernel.desc is None:
                        o = self._custom_chain(kernel, b, w0[sl], first, nsteps)
                    else:
                        pc = dict(P=p_buf[:sl.stop - sl.start], valid=0) if p_buf is not None else None
                        o = plan.prune_chain(kernel.desc, self._Xd, self._cluster_of_d, b["w0"], b["alive"], b["ga"],
                                             b["yb"], b["keys"], b["lmin"], b["gmin"], first, nsteps, keep_yb=True,
                                             p_cache=pc)
                    for k in ("noise", "z_low", "z_high", "gamma", "pruned", "act"):
                        chain[k][sl, first:first + nsteps] = o[k]
                    yb_chain[sl, first + 1:first + 1 + nsteps] = o["yb"]
                    del o
            run(0, steps_planned)
            done = steps_planned
            # The reference keeps stepping until every mask of the group is empty or L steps are done
            # (MAIN:659-660).  An empty row only wins the argmin in degenerate cases (activation > 2 or
            # NaN), so this loop normally does not run.
            while done < loop_number and any(
                    self._nonempty(w0[b["sl"]], b["alive"].cpu().numpy()).any() for b in batches):
                grow(1)
                run(done, 1)
                done += 1
        torch.cuda.synchronize(dev)
        del p_buf
        self._held_bytes = 0

        S = done
        noises = np.empty((Tk, S + 1)); zl = np.empty((Tk, S + 1)); zh = np.empty((Tk, S + 1)); gm = np.empty((Tk, S + 1))
        noises[:, 0], zl[:, 0], zh[:, 0], gm[:, 0] = noise_kernel, Z_low_kernel, Z_high_kernel, gammas_kernel
        noises[:, 1:] = chain["noise"].cpu().numpy()
        zl[:, 1:] = chain["z_low"].cpu().numpy()
        zh[:, 1:] = chain["z_high"].cpu().numpy()
        gm[:, 1:] = chain["gamma"].cpu().numpy()
        pruned = chain["pruned"].cpu().numpy().astype(np.int64)
        activations = chain["act"].cpu().numpy()
        # chain of variable masks (T_k, S+1, N) from the pruned cluster indices
        active = np.empty((Tk, S + 1, N))
        alive_h = np.ones((Tk, C))
        cluster_of = self.modes.cluster_of
        active[:, 0] = w0 * alive_h[:, cluster_of]
        for st in range(S):
            alive_h[np.arange(Tk), pruned[:, st]] = 0
            active[:, st + 1] = w0 * alive_h[:, cluster_of]
        self._chain_pruned = pruned
        return active, noises, zl, zh, activations, gm, yb_chain

    def _nonempty(self, w0, alive):
        im = self.modes.index_matrix
        return ((im[None, :, :] * w0[:, None, :]).sum(axis=2) > 0) & (alive > 0)

    # ------------------------------------------------------------------ results
    def process_results(self, targets, chosen_kernel, mask_kernel, chosen_modes, ancestor_modess, noisess, Z_lows,
                        Z_highs, activationss, kernel_performances, gammas, ybs_kernel):
        """Prints the per-node report and writes nodes/edges of G (MAIN:703-817)."""
        index = 0
        for t, (name, mask) in enumerate(zip(targets, mask_kernel)):
            if not mask:
                continue
            self.print_func(f"\nResults for {name}")
            noises_kernel, Z_lows_kernels, Z_highs_kernels, gammas_kernel = (kernel_performances[j][t] for j in
                                                                            (1, 2, 3, 4))
            for i, kernel in enumerate(self.kernels):
                self.print_func(
                    f"Kernel [{kernel}] has n/(n+s)={noises_kernel[i]}, Z=({Z_lows_kernels[i]:.2f}, "
                    f"{Z_highs_kernels[i]:.2f}), gamma={max(self.gamma_min, gammas_kernel[i]):.2e}")
            if chosen_kernel == -1:
                self.print_func(f"{name} has no ancestors\n")
                index += 1
                continue
            chosen_mode = int(chosen_modes[index])
            ancestor_modes = ancestor_modess[index]
            noises = noisess[index]
            Z_low, Z_high = Z_lows[index], Z_highs[index]
            activations = activationss[index]
            cm = min(chosen_mode, noises.shape[0] - 1)
            gamma = gammas[index, cm]
            yb = ybs_kernel[index, cm].cpu().numpy()
            self.print_func(
                f"{name} has ancestors with the kernel [{self.kernels[chosen_kernel]}] | "
                f"(n/(s+n)={float(noises[cm]):.2f} after pruning)")
            if have_matplotlib() and self.verbose:
                import matplotlib.pyplot as plt
                ancestor_number = [np.sum(mode) for mode in ancestor_modes]
                fig, _axes = plot_noise_evolution(ancestor_number, noises, [(a, b) for a, b in zip(Z_low, Z_high)],
                                                  node_name=name, ancestor_modes_number=ancestor_number[cm])
                plt.show()
                plt.close(fig)
            # JAX clamps the out-of-range read activations[chosen_mode] (MAIN:790)
            act_c = activations[min(chosen_mode, activations.shape[0] - 1)]
            acivation_per_variable = np.sum(self.modes.index_matrix * act_c[:, None], axis=0)
            active_variables = ancestor_modes[cm]
            self.G.nodes[name].update({
                "active_modes": active_variables, "kernel_index": chosen_kernel,
                "type": str(self.kernels[chosen_kernel]), "gamma": float(gamma), "noise": float(noises[cm]),
                "coeff": yb})
            ancestor_names = []
            for i, (activation, active) in enumerate(zip(acivation_per_variable, active_variables)):
                if active == 1:
                    ancestor_names.append(self.names[i])
                    self.G.add_edge(self.names[i], name, signal=activation)
                elif active == 0:
                    continue
                else:
                    raise ValueError("Inconsistent activation")
            self.print_func(f"Ancestors of {name}: {ancestor_names}\n")
            index += 1

    def plot_graph(self, type_label=True, **kwargs):
        """Draws G (MAIN:819-932); out of the hot path: a plain networkx drawing (clusters through
        util.partition_layout), matplotlib required."""
        if not have_matplotlib():
            raise RuntimeError("plot_graph needs matplotlib, which is not installed")
        if self.has_clusters:
            partition = {item: i for i, sublist in enumerate(self.modes.clusters) for item in sublist}
            pos, _hyper = partition_layout(self.G, partition)
            pos = {k: v for k, v in pos.items() if k in self.G}
        else:
            pos = nx.kamada_kawai_layout(self.G)
        nx.draw_networkx(self.G, pos=pos, with_labels=kwargs.get("with_labels", True),
                         node_size=kwargs.get("node_size", 400), font_size=kwargs.get("font_size", 8),
                         alpha=kwargs.get("alpha", 0.6))
        if type_label:
            nx.draw_networkx_edge_labels(self.G, pos, edge_labels=nx.get_edge_attributes(self.G, "type"))

    # ------------------------------------------------------------------ predict
    def predict(self, names, X_pred):
        """Kernel predictions for `names` at new inputs (MAIN:934-979)."""
        X_pred = np.asarray(X_pred, dtype=np.float64)
        assert X_pred.shape[1] == self.X.shape[1]
        plan = self._device()
        dev = plan.device
        res = np.zeros((X_pred.shape[0], len(names)))
        X_pred_used = (X_pred - self.mean_x) / self.std_x
        k_dic = nx.get_node_attributes(self.G, "kernel_index")
        kernels = np.array([k_dic.get(name, -1) for name in names])
        active_mode_dic = nx.get_node_attributes(self.G, "active_modes")
        active_modess = np.stack([active_mode_dic.get(name, np.zeros(self.X.shape[1])) for name in names], axis=0)
        yb_dic = nx.get_node_attributes(self.G, "coeff")
        ybs = np.stack([yb_dic.get(name, np.zeros(self.X.shape[0])) for name in names], axis=0)
        Xq = torch.from_numpy(np.ascontiguousarray(X_pred_used)).to(dev)
        for i, kernel in enumerate(self.kernels):
            mask = kernels == i
            if not np.any(mask):
                continue
            if kernel.desc is None:   # user ModeKernel: host-provided cross-Gram
                cols = np.nonzero(mask)[0]
                for ci, am, yb_t in zip(cols, active_modess[mask], ybs[mask]):
                    K = np.asarray(kernel(X_pred_used, self.X, (am != 0).astype(np.float64)))
                    res[:, ci] = K @ (-yb_t)
                continue
            w = torch.from_numpy(np.ascontiguousarray(active_modess[mask] != 0).astype(np.uint8)).to(dev)
            yb = torch.from_numpy(np.ascontiguousarray(ybs[mask])).to(dev)
            pred = plan.predict(kernel.desc, self._Xd, Xq, w, yb)        # (T_k, nq)
            res[:, mask] = pred.cpu().numpy().T
        indexes = np.array([self.name_to_index[name] for name in names])
        return res * self.std_x[:, indexes] + self.mean_x[:, indexes]

    def show_functional_dependencies(self):
        """MAIN:981-1002."""
        for var in self.names:
            data = self.G.nodes.get(var)
            active_modes = data.get("active_modes")
            if active_modes is not None:
                print(f"{var.strip('$')}")
                kType, gamma, noise = data.get("type"), data.get("gamma"), data.get("noise")
                assert len(active_modes) == len(self.names)
                first = True
                for act, dep in zip(active_modes, self.names):
                    if act != 0.0:
                        lead = " = " if first else " + "
                        print(f"{lead}{act} * {dep.strip('$')} {{kernel={kType}, gamma={gamma:.3g}, "
                              f"noise={noise:.3g}}}")
                        first = False
