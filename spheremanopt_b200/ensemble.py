"""Ensemble driver: many *unmodified* optimiser instances share batched kernel launches (BASELINE config 5).

``Optimise_On_Multi_Sphere`` (SGD:692) calls ``f`` / ``Grad_f`` / ``Inner_Product`` strictly one after the other, so a single
optimisation can never fill a batched kernel.  Here K optimisations run in K Python threads; every call of an instance's
callables parks its thread at a rendezvous, and as soon as every live instance is parked all pending calls of one kind are
served by ONE batched backend call (one ``Sh23Fwd`` / ``Sh23Adj`` launch for the whole group).  Instances drift apart
(line searches take different numbers of trials, some instances converge early): a group is whatever subset currently
asks for the same kind of call.  Results are independent of the grouping - every instance is computed by its own threads
of the batched kernel - so an ensemble run reproduces the K individual runs.

    ens = SH23Ensemble(K, domain, dt, N_ITERS)                       # batched backends + per-instance snapshot rows
    out = ens.run(lambda i, f, grad, ip: Optimise_On_Multi_Sphere([X0[i]], [M0[i]], f, grad, ip, ...))

The rendezvous itself (``Rendezvous``) is independent of the backend: it only needs three batched callables.
Note: the unmodified optimiser appends to ``optimize_result.txt`` in the current directory (SGD:730); all instances of
an ensemble therefore share that file.
"""
import threading

import numpy as np
import torch


def shard_slice(total, world, rank):
    """instances [lo, hi) of an ensemble of `total` that rank `rank` of `world` owns: contiguous blocks, sizes differing by at
    most one (BASELINE config 5: 4096 instances over 8 GPUs = 512 each; no communication between the shards)"""
    total, world, rank = int(total), int(world), int(rank)
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank %d/%d" % (world, rank))
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class Rendezvous:
    """K workers, three batched services.  ``batched[kind](ids, args_list) -> list of results`` (ids ascending).

    Scales to hundreds of instances per GPU (config 5: 512 per GPU): every worker parks on an event of its OWN, and whoever
    completes a group wakes exactly the workers it served (no notify_all herd); worker threads get small stacks."""

    def __init__(self, n, batched_f, batched_grad, batched_ip):
        self.n = int(n)
        self.batched = {"f": batched_f, "grad": batched_grad, "ip": batched_ip}
        self.lock = threading.Lock()
        self.wake = {}           # id -> threading.Event of the parked worker
        self.pending = {}        # id -> (kind, args)
        self.results = {}        # id -> result or exception
        self.live = 0
        self.rounds = {"f": 0, "grad": 0, "ip": 0}     # batched backend calls made (for reporting)
        self.served = {"f": 0, "grad": 0, "ip": 0}     # individual calls served

    # -- called with self.lock held: serve one kind after the other while every live worker is parked
    def _serve_if_ready(self):
        while self.pending and len(self.pending) == self.live:
            # the kind most workers wait for goes first (keeps groups large)
            kinds = {}
            for i, (k, _) in self.pending.items():
                kinds.setdefault(k, []).append(i)
            kind = max(kinds, key=lambda k: len(kinds[k]))
            ids = sorted(kinds[kind])
            args = [self.pending[i][1] for i in ids]
            try:
                out = self.batched[kind](ids, args)
                if len(out) != len(ids):
                    raise RuntimeError("batched %s returned %d results for %d requests" % (kind, len(out), len(ids)))
            except Exception as e:           # hand the failure to every waiting caller of this group
                out = [e] * len(ids)
            self.rounds[kind] += 1
            self.served[kind] += len(ids)
            for i, r in zip(ids, out):
                del self.pending[i]
                self.results[i] = r
                self.wake[i].set()

    def call(self, i, kind, args):
        ev = self.wake[i]
        with self.lock:
            ev.clear()
            self.pending[i] = (kind, args)
            self._serve_if_ready()
        ev.wait()
        with self.lock:
            r = self.results.pop(i)
        if isinstance(r, Exception):
            raise r
        return r

    def callables(self, i):
        return (lambda X, *a, **k: self.call(i, "f", X),
                lambda X, *a, **k: self.call(i, "grad", X),
                lambda x, y, *a, **k: self.call(i, "ip", (x, y)))

    def run(self, target, ids=None):
        """runs target(i, f_i, grad_i, ip_i) for every instance in its own thread; returns the list of return values"""
        ids = list(range(self.n)) if ids is None else list(ids)
        out = {}

        def worker(i):
            try:
                out[i] = target(i, *self.callables(i))
            except Exception as e:
                out[i] = e
            finally:
                with self.lock:
                    self.live -= 1
                    self._serve_if_ready()      # the others may all be parked already

        with self.lock:
            self.live = len(ids)
            self.wake = {i: threading.Event() for i in ids}
        old = threading.stack_size()
        try:
            threading.stack_size(1 << 20)       # hundreds of workers: 1 MB stacks instead of the platform default (8 MB)
        except (ValueError, RuntimeError):
            pass
        try:
            threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in ids]
            for t in threads:
                t.start()
        finally:
            try:
                threading.stack_size(old)
            except (ValueError, RuntimeError):
                pass
        for t in threads:
            t.join()
        for i in ids:
            if isinstance(out[i], Exception):
                raise out[i]
        return [out[i] for i in ids]


class SH23Ensemble(Rendezvous):
    """Batched CUDA backends for K independent SH23 problems on one GPU (host vectors in, host vectors out).

    Every instance owns one row of a persistent snapshot store [K][smo_sh23_snapshot_bytes]; a group of instances is solved in a
    compact temporary store and its rows are scattered to / gathered from the persistent one, so ``Grad_f`` of an instance
    always replays the forward solve of the same instance whatever the grouping was."""

    def __init__(self, n, domain, dt, N_ITERS, Adjoint_type="Discrete"):
        from . import sh23
        self._sh = sh23
        self.domain, self.dt, self.nit, self.adj = domain, float(dt), int(N_ITERS), Adjoint_type
        self.row = domain.lib.smo_sh23_snapshot_bytes(domain.h, self.nit) // 8     # doubles per instance (opaque store)
        self.store = torch.zeros(n, self.row, dtype=torch.float64, device=domain.device)
        self.tag = [None] * n
        self._tmp = {}
        self._scratch = None
        super().__init__(n, self._f, self._grad, self._ip)

    def _tmp_store(self, k):
        # group sizes drift between 1 and n during a run: every size is a VIEW of the first k rows of one scratch block (a store
        # of its own per distinct size would add up to ~n^2/2 rows - tens of GB for 512 instances of config 1)
        if self._scratch is None:
            self._scratch = torch.zeros(self.n * self.row, dtype=torch.float64, device=self.domain.device)
        if k not in self._tmp:
            self._tmp[k] = self._sh.SnapshotStore(self.domain, self.nit, batch=k, buf=self._scratch[:k * self.row])
        return self._tmp[k]

    def _stack(self, vecs):
        a = np.ascontiguousarray(np.stack([np.asarray(v, dtype=np.float64).reshape(-1) for v in vecs]))
        return torch.from_numpy(a).to(self.domain.device).reshape(-1)

    def _solve(self, ids, Xs):
        st = self._tmp_store(len(ids))
        J = self._sh.forward_batch(self._stack([X[0] for X in Xs]), self.domain, self.dt, self.nit, st)
        idx = torch.tensor(ids, device=self.domain.device)
        self.store.index_copy_(0, idx, st.buf.view(len(ids), self.row))
        for i, X in zip(ids, Xs):
            self.tag[i] = self._sh.fingerprint(X[0])
        return J

    def _f(self, ids, Xs):
        J = self._solve(ids, Xs).cpu().numpy()
        return [(-1.) * float(j) for j in J]

    def _grad(self, ids, Xs):
        stale = [k for k, (i, X) in enumerate(zip(ids, Xs)) if self.tag[i] != self._sh.fingerprint(X[0])]
        if stale:        # Grad_f before f, or at another X: fill those rows first (never happens in the reference optimiser)
            self._solve([ids[k] for k in stale], [Xs[k] for k in stale])
        st = self._tmp_store(len(ids))
        idx = torch.tensor(ids, device=self.domain.device)
        st.buf.view(len(ids), self.row).copy_(self.store.index_select(0, idx))
        st.valid = True
        G = self._sh.adjoint_batch(self.domain, self.dt, self.nit, st, self.adj).view(len(ids), self.domain.M).cpu().numpy()
        return [[G[k].copy()] for k in range(len(ids))]

    def _ip(self, ids, pairs):
        # ONE launch for the whole group (smo_vec_dot_rows: one CTA per instance), one D2H of len(ids) doubles
        from . import _cabi
        from .devvec import _stream_ptr
        x = self._stack([p[0] for p in pairs])
        y = self._stack([p[1] for p in pairs])
        n = x.numel() // len(ids)
        out = torch.empty(len(ids), dtype=torch.float64, device=self.domain.device)
        with torch.cuda.device(self.domain.device):
            _cabi.check(self.domain.lib, self.domain.lib.smo_vec_dot_rows(x.data_ptr(), y.data_ptr(), len(ids), n, 1.0 / self.domain.M,
                                                                         out.data_ptr(), _stream_ptr()))
        return out.cpu().tolist()
