"""Side outputs: the .npz stand-ins of the reference's CheckPoints / scalar_data HDF5 handlers and File_Manips."""
import importlib.util
import os
import tempfile

import numpy as np
import pytest

from oracle import kdyn as okd
from oracle import sh23 as osh
from tests.common import kdyn_field

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sideout():
    # loaded by path: importing the package would pull in torch.cuda-dependent modules, which is fine on CPU, but this test
    # only needs the formatting code
    spec = importlib.util.spec_from_file_location("smo_sideout", os.path.join(ROOT, "spheremanopt_b200", "sideout.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture()
def tmpcwd():
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    yield
    os.chdir(cwd)


def test_sh23_handlers_from_oracle_snapshots(tmpcwd):
    so = _sideout()
    dom, X0 = osh.Generate_IC(0.0725)
    D = osh.GEN_BUFFER(dom, 60)
    osh.FWD_Solve_IVP_Lin([X0], dom, 0.1, 60, 60, D)

    class Dm:
        N, L, interval = 256, dom.L, (0., dom.L)
    so.sh23_outputs(Dm, D['A_fwd'], 0.1, 60)
    s, c = np.load("scalar_data/scalar_data_s1.npz"), np.load("CheckPoints/CheckPoints_s1.npz")
    assert np.allclose(s["scales/sim_time"], [0., 2., 4., 6.]) and s["tasks/Kinetic energy"].shape == (4, 1)
    for j, n in enumerate((0, 20, 40, 60)):      # SH:483: inv_Vol*integ(u**2) = grid mean of u_n^2
        u = dom.to_grid_1d(D['A_fwd'][:, n])
        assert abs(s["tasks/Kinetic energy"][j, 0] - (u * u).mean()) <= 1e-13
    assert c["tasks/u"].shape == (2, 384) and c["scales/x/1.5"].shape == (384,) and c["tasks/u_hat"].shape == (2, 128)
    assert abs((c["tasks/u"][0] ** 2).mean() - 0.0725) <= 1e-12           # the same band-limited field on the 3/2 grid (SH:479)
    assert np.array_equal(c["tasks/u_hat"][1], D['A_fwd'][:, 60])
    so.file_manips(7)
    assert os.path.exists("scalar_data_iter_7.npz") and os.path.exists("CheckPoints_iter_7.npz")


def test_file_manips_without_outputs_is_loud(tmpcwd):
    with pytest.raises(FileNotFoundError):
        _sideout().file_manips(0)


@pytest.mark.gpu
def test_side_outputs_of_the_cuda_callables(tmpcwd):
    from spheremanopt_b200 import kdyn, sh23
    sh23.SIDE_OUTPUTS = kdyn.SIDE_OUTPUTS = True
    try:
        dom, X0 = sh23.Generate_IC(0.0725)
        st = sh23.GEN_BUFFER(dom, 40)
        sh23.FWD_Solve_IVP_Lin([X0], dom, 0.1, 40, 40, st)
        od, X0o = osh.Generate_IC(0.0725)
        D = osh.GEN_BUFFER(od, 40)
        osh.FWD_Solve_IVP_Lin([X0o], od, 0.1, 40, 40, D)
        s = np.load("scalar_data/scalar_data_s1.npz")
        for j, n in enumerate((0, 20, 40)):
            u = od.to_grid_1d(D['A_fwd'][:, n])
            assert abs(s["tasks/Kinetic energy"][j, 0] - (u * u).mean()) <= 1e-9 * (u * u).mean()
        sh23.File_Manips(0)
        # dynamo
        kd = kdyn.Domain(16)
        okd_ = okd.domain_kdyn(16)
        B0, U = kdyn_field(okd_, 1), kdyn_field(okd_, 2)
        ks = kdyn.GEN_BUFFER(16, kd, 40, checkpoint_every=0)
        kdyn.FWD_Solve_IVP_Lin([B0, U], kd, 1.0, 1e-3, 40, 40, ks)
        Dk = okd.GEN_BUFFER(16, okd_, 40)
        okd.FWD_Solve_IVP_Lin([B0, U], okd_, 1.0, 1e-3, 40, 40, Dk)
        s, c = np.load("scalar_data/scalar_data_s1.npz"), np.load("CheckPoints/CheckPoints_s1.npz")
        assert s["tasks/Magnetic energy"].shape == (3, 1, 1, 1) and c["tasks/A"].shape == (2, 24, 24, 24)
        for j, n in enumerate((0, 20, 40)):
            g = [okd_.to_grid_3d(Dk[k][..., n]) for k in ('A_fwd', 'B_fwd', 'C_fwd')]
            e = sum((gi * gi).mean() for gi in g)
            assert abs(s["tasks/Magnetic energy"][j, 0, 0, 0] - e) <= 1e-9 * e
        assert np.abs(c["tasks/A"][0] - B0.reshape(3, 24, 24, 24)[0]).max() <= 1e-12
        assert np.abs(c["tasks/C"][1] - okd_.to_grid_3d(Dk['C_fwd'][..., 40])).max() <= 1e-9 * np.abs(c["tasks/C"][1]).max()
        kdyn.File_Manips(1)
        assert os.path.exists("CheckPoints_iter_1.npz")
    finally:
        sh23.SIDE_OUTPUTS = kdyn.SIDE_OUTPUTS = False
