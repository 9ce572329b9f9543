"""Ensemble driver (spheremanopt_b200/ensemble.py): K optimiser instances in K threads share batched backend calls.
CPU: the rendezvous logic with the oracle as a (looping) batched backend - with the UNMODIFIED reference optimiser when
/root/reference is mounted, and with a small descent loop otherwise.  GPU: the batched CUDA backends."""
import numpy as np
import pytest

from oracle import sh23 as osh
from oracle import sphere as osp
from tests.common import relerr, sh23_input


def _descent(X0, M0, f, grad, ip, iters):
    """projected gradient descent with a halving line search (uses f / grad / ip the way the reference optimiser does;
    the number of f calls per iteration depends on the instance, so instances of an ensemble drift apart)"""
    X = X0.copy()
    hist = []
    fx = f([X])
    for _ in range(iters):
        g = osp.tangent_vector(X, grad([X])[0], ip)
        alpha = 2.0
        for _ in range(8):
            Xn = osp.Update_vector(X, -alpha, g, M0, ip)
            fn = f([Xn])
            if fn < fx:
                break
            alpha *= 0.5
        X, fx = Xn, fn
        hist.append(fx)
    return X, hist


def _oracle_backends(od, dt, nit, K):
    stores = [osh.GEN_BUFFER(od, nit) for _ in range(K)]
    calls = {"f": 0, "grad": 0, "ip": 0}

    def bf(ids, Xs):
        calls["f"] += 1
        return [osh.FWD_Solve_IVP_Lin(X, od, dt, nit, nit, stores[i]) for i, X in zip(ids, Xs)]

    def bg(ids, Xs):
        calls["grad"] += 1
        return [osh.ADJ_Solve_IVP_Lin(X, od, dt, nit, nit, stores[i]) for i, X in zip(ids, Xs)]

    def bi(ids, pairs):
        calls["ip"] += 1
        return [osh.Inner_Prod(x, y, od) for x, y in pairs]
    return bf, bg, bi, calls


def test_rendezvous_reproduces_individual_runs():
    from spheremanopt_b200.ensemble import Rendezvous
    od = osh.domain_sh23(64)
    K, dt, nit = 5, 0.1, 8
    M0 = [0.02 + 0.01 * i for i in range(K)]
    X0 = []
    for i in range(K):
        x = sh23_input(od, seed=10 + i)
        X0.append(x * np.sqrt(M0[i] / osh.Inner_Prod(x, x, od)))
    iters = [2, 3, 1, 4, 2]          # different lengths: instances finish at different times
    bf, bg, bi, calls = _oracle_backends(od, dt, nit, K)
    ens = Rendezvous(K, bf, bg, bi)
    out = ens.run(lambda i, f, grad, ip: _descent(X0[i], M0[i], f, grad, ip, iters[i]))
    assert ens.served["f"] > ens.rounds["f"] and ens.served["ip"] > ens.rounds["ip"]      # calls really were batched
    for i in range(K):
        st = osh.GEN_BUFFER(od, nit)
        Xi, hi = _descent(X0[i], M0[i], lambda X: osh.FWD_Solve_IVP_Lin(X, od, dt, nit, nit, st),
                          lambda X: osh.ADJ_Solve_IVP_Lin(X, od, dt, nit, nit, st), lambda x, y: osh.Inner_Prod(x, y, od), iters[i])
        assert out[i][1] == hi and np.array_equal(out[i][0], Xi)


def test_rendezvous_propagates_errors():
    from spheremanopt_b200.ensemble import Rendezvous

    def bad(ids, args):
        raise ValueError("backend failure")
    ens = Rendezvous(3, bad, bad, bad)
    with pytest.raises(ValueError):
        ens.run(lambda i, f, grad, ip: f([np.zeros(4)]))


def test_reference_optimiser_in_an_ensemble(refopt, tmp_path, monkeypatch):
    """three UNMODIFIED Optimise_On_Multi_Sphere instances (different M_0) through the rendezvous = three separate runs"""
    from spheremanopt_b200.ensemble import Rendezvous
    SGD, TG = refopt
    monkeypatch.chdir(tmp_path)
    od = osh.domain_sh23(64)
    K, dt, nit = 3, 0.1, 10
    M0 = [0.03, 0.05, 0.08]
    X0 = []
    for i in range(K):
        x = sh23_input(od, seed=3)
        X0.append(x * np.sqrt(M0[i] / osh.Inner_Prod(x, x, od)))

    def opt(i, f, grad, ip):
        return SGD.Optimise_On_Multi_Sphere([X0[i].copy()], [M0[i]], f, grad, ip, max_iters=3, alpha_k=np.pi, LS='LS_wolfe', CG=True,
                                            callback=None, verbose=False)
    bf, bg, bi, calls = _oracle_backends(od, dt, nit, K)
    ens = Rendezvous(K, bf, bg, bi)
    out = ens.run(opt)
    assert ens.served["f"] > ens.rounds["f"]
    for i in range(K):
        st = osh.GEN_BUFFER(od, nit)
        RES, FUN, Xo = opt(i, lambda X: osh.FWD_Solve_IVP_Lin(X, od, dt, nit, nit, st),
                           lambda X: osh.ADJ_Solve_IVP_Lin(X, od, dt, nit, nit, st), lambda x, y: osh.Inner_Prod(x, y, od))
        assert np.allclose(out[i][1], FUN, rtol=1e-13, atol=0) and relerr(out[i][2][0], Xo[0]) <= 1e-12


@pytest.mark.gpu
def test_sh23_ensemble_on_gpu():
    """batched CUDA backends: an ensemble of descents (instances drifting apart) equals the individual CUDA runs and the oracle"""
    from spheremanopt_b200 import sh23
    from spheremanopt_b200.ensemble import SH23Ensemble
    od = osh.domain_sh23(256)
    dom = sh23.Domain(256)
    K, dt, nit = 7, 0.1, 40
    M0 = np.linspace(0.05, 0.1, K)
    x = sh23_input(od, seed=4)
    X0 = [x * np.sqrt(m / osh.Inner_Prod(x, x, od)) for m in M0]
    iters = [2, 1, 3, 2, 1, 3, 2]
    ens = SH23Ensemble(K, dom, dt, nit)
    out = ens.run(lambda i, f, grad, ip: _descent(X0[i], M0[i], f, grad, ip, iters[i]))
    assert ens.served["f"] > ens.rounds["f"]
    st = sh23.GEN_BUFFER(dom, nit)
    for i in range(K):
        Xi, hi = _descent(X0[i], M0[i], lambda X: sh23.FWD_Solve_IVP_Lin(X, dom, dt, nit, nit, st),
                          lambda X: sh23.ADJ_Solve_IVP_Lin(X, dom, dt, nit, nit, st), lambda a, b: sh23.Inner_Prod(a, b, dom), iters[i])
        assert np.allclose(out[i][1], hi, rtol=1e-13, atol=0) and relerr(out[i][0], Xi) <= 1e-12
        D = osh.GEN_BUFFER(od, nit)
        Xo, ho = _descent(X0[i], M0[i], lambda X: osh.FWD_Solve_IVP_Lin(X, od, dt, nit, nit, D),
                          lambda X: osh.ADJ_Solve_IVP_Lin(X, od, dt, nit, nit, D), lambda a, b: osh.Inner_Prod(a, b, od), iters[i])
        assert np.allclose(out[i][1], ho, rtol=1e-9, atol=0)


def test_sh23_ensemble_group_stores_are_views_of_one_scratch_block():
    """group sizes drift during a run: the temporary store of every size is a view of the first k instance rows of ONE block
    (host logic only: a stub domain on the CPU, no library call besides the size query)"""
    import torch
    from spheremanopt_b200 import sh23
    from spheremanopt_b200.ensemble import SH23Ensemble

    class _Lib:
        def smo_sh23_snapshot_bytes(self, h, nit):
            return 8 * ((nit + 1) * 16 + 4)

    class _Dom:
        lib, h, device, M = _Lib(), None, torch.device("cpu"), 16

    K, nit = 6, 3
    ens = SH23Ensemble(K, _Dom(), 0.1, nit)
    row = ens.row
    assert row == (nit + 1) * 16 + 4 and ens.store.shape == (K, row)
    s2, s6, s2b = ens._tmp_store(2), ens._tmp_store(6), ens._tmp_store(2)
    assert s2 is s2b and s2.batch == 2 and s6.batch == 6
    assert s2.buf.numel() == 2 * row and s6.buf.numel() == 6 * row
    assert s2.buf.data_ptr() == s6.buf.data_ptr() == ens._scratch.data_ptr() and ens._scratch.numel() == K * row
    s6.buf.view(6, row)[1].fill_(7.0)
    assert float(s2.buf.view(2, row)[1].min()) == 7.0 and float(s2.buf.view(2, row)[0].abs().max()) == 0.0
    with pytest.raises(ValueError):
        sh23.SnapshotStore(_Dom(), nit, batch=2, buf=torch.zeros(2 * row + 1, dtype=torch.float64))
    own = sh23.SnapshotStore(_Dom(), nit, batch=3)
    assert own.buf.numel() == 3 * row and own.buf.data_ptr() != ens._scratch.data_ptr()


def test_rendezvous_scales_to_many_workers():
    """256 workers with drifting call sequences: per-worker wake-ups, no lost wake-up, every call served exactly once"""
    from spheremanopt_b200.ensemble import Rendezvous
    K = 256
    calls = {"f": 0, "grad": 0, "ip": 0}

    def bf(ids, args):
        calls["f"] += 1
        return [float(i) + a for i, a in zip(ids, args)]

    def bg(ids, args):
        calls["grad"] += 1
        return [[2.0 * a] for a in args]

    def bi(ids, args):
        calls["ip"] += 1
        return [x * y for x, y in args]
    R = Rendezvous(K, bf, bg, bi)

    def target(i, f, g, ip):
        acc = 0.0
        for it in range(3 + i % 5):              # different lengths: workers finish at different times
            acc += f(1.0 * it)
            if (i + it) % 3 == 0:
                acc += g(0.5)[0]
            for _ in range(i % 4):
                acc += ip(2.0, 0.25)
        return acc
    out = R.run(target)
    for i in range(K):
        want = 0.0
        for it in range(3 + i % 5):
            want += i + it
            if (i + it) % 3 == 0:
                want += 1.0
            want += 0.5 * (i % 4)
        assert out[i] == want
    assert R.served["f"] == sum(3 + i % 5 for i in range(K)) and calls["f"] == R.rounds["f"] < R.served["f"] / 20


def test_shard_slice_partitions_the_ensemble():
    """config 5's sharding: contiguous, disjoint, covering, balanced (4096 over 8 = 512 each; ragged totals differ by <= 1)"""
    from spheremanopt_b200.ensemble import shard_slice
    for total, world in ((4096, 8), (4096, 1), (10, 4), (3, 8), (4097, 8)):
        parts = [shard_slice(total, world, r) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == total
        assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1
    assert shard_slice(4096, 8, 3) == (1536, 2048)
    with pytest.raises(ValueError):
        shard_slice(8, 2, 2)
