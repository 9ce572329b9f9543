"""Side outputs of the reference's forward solves: the two Dedalus file handlers

    CheckPoints  (SH:478-480, KD:606-610)  iter = N_SUB_ITERS : state on the 3/2 grid (+ coefficients / velocity) at iterations 0 and N_ITERS
    scalar_data  (SH:482-483, KD:612-613)  iter = 20          : "Kinetic energy" / "Magnetic energy" = grid mean of u^2 / |B|^2

as ``<name>/<name>_s1.npz`` whose keys are the HDF5 dataset paths the reference's plot scripts read (``tasks/u``,
``tasks/u_hat``, ``tasks/A`` .., ``tasks/u-velocity`` .., ``tasks/Kinetic energy``, ``scales/sim_time``, ``scales/x/1.5`` ..;
[D2-11]), and additionally as ``.h5`` with the same layout when h5py is importable (it is not in this image).  Off by
default - the reference's HDF5 traffic is what dominates its small-problem wall time; switch on with
``SMO_SIDE_OUTPUTS=1`` or ``sh23.SIDE_OUTPUTS = True`` / ``kdyn.SIDE_OUTPUTS = True`` so that ``callback=File_Manips``
(SH:731-746, KD:1006-1021) finds its files.  Everything here is output formatting on the host, computed from the
device-resident snapshot store after the solve; nothing on the f / Grad_f path depends on it.
"""
import os
import shutil

import numpy as np


def enabled(module_flag):
    return bool(module_flag) or os.environ.get("SMO_SIDE_OUTPUTS", "0") not in ("", "0")


def write_handler(name, datasets):
    """datasets: {hdf5 path: array}.  Writes <name>/<name>_s1.npz (always) and <name>/<name>_s1.h5 (when h5py exists)."""
    os.makedirs(name, exist_ok=True)
    stem = os.path.join(name, "%s_s1" % name)
    np.savez(stem + ".npz", **datasets)
    try:
        import h5py
        if not hasattr(h5py, "File"):
            raise ImportError
        with h5py.File(stem + ".h5", "w") as fh:
            for k, v in datasets.items():
                fh.create_dataset(k, data=v)
    except (ImportError, OSError):      # no h5py / no HDF5 library behind it: the .npz is the output
        pass
    return stem


def file_manips(k):
    """SH:731-746 / KD:1006-1021: keep the outputs of optimiser iteration k"""
    done = 0
    for name in ("scalar_data", "CheckPoints"):
        for ext in (".h5", ".npz"):
            src = os.path.join(name, "%s_s1%s" % (name, ext))
            if os.path.exists(src):
                shutil.copyfile(src, "%s_iter_%i%s" % (name, k, ext))
                done += 1
    if done == 0:
        raise FileNotFoundError("File_Manips: no scalar_data/ or CheckPoints/ outputs in %s - enable them with SMO_SIDE_OUTPUTS=1 "
                                "(or pass callback=None to the optimiser)" % os.getcwd())
    return None


def sh23_outputs(domain, coeffs, dt, n_iters):
    """coeffs: complex [Npts/2, n_iters+1] (the snapshot store in the reference's orientation)"""
    N, L = domain.N, domain.L
    c = np.asarray(coeffs)
    n20 = np.arange(0, n_iters + 1, 20)
    energy = (np.abs(c[0, n20]) ** 2 + 2.0 * (np.abs(c[1:, n20]) ** 2).sum(axis=0)).real       # Parseval = grid mean of u^2
    write_handler("scalar_data", {"tasks/Kinetic energy": energy.reshape(-1, 1), "scales/sim_time": n20 * float(dt),
                                  "scales/iteration": n20})
    M15 = 3 * N // 2
    wr = np.array([0, n_iters])
    half = np.zeros((2, M15 // 2 + 1), dtype=np.complex128)
    half[:, :N // 2] = c[:, wr].T
    half[:, 0] = half[:, 0].real
    u = np.fft.irfft(half, n=M15, axis=1) * M15                                               # coefficients are amplitudes [D2-2]
    write_handler("CheckPoints", {"tasks/u": u, "tasks/u_hat": c[:, wr].T.copy(), "scales/sim_time": wr * float(dt), "scales/iteration": wr,
                                  "scales/x/1.5": domain.interval[0] + L * np.arange(M15) / M15,
                                  "scales/x/1.0": domain.interval[0] + L * np.arange(N) / N})


def kdyn_outputs(domain, energy_iters, energies, B_first, B_last, U_proj, dt, n_iters):
    """B_first / B_last / U_proj: full host vectors (3*M^3) on the 3/2 grid; energies: <B^n,B^n> at energy_iters"""
    M, L, x0 = domain.M, domain.L, domain.interval[0]
    it = np.asarray(energy_iters)
    write_handler("scalar_data", {"tasks/Magnetic energy": np.asarray(energies, dtype=np.float64).reshape(-1, 1, 1, 1),
                                  "scales/sim_time": it * float(dt), "scales/iteration": it})
    wr = np.array([0, n_iters])
    comp = lambda v: np.asarray(v).reshape(3, M, M, M)
    Bf, Bl, Up = comp(B_first), comp(B_last), comp(U_proj)
    g = x0 + L * np.arange(M) / M
    d = {"scales/sim_time": wr * float(dt), "scales/iteration": wr, "scales/x/1.5": g, "scales/y/1.5": g, "scales/z/1.5": g}
    for i, nm in enumerate(("A", "B", "C")):
        d["tasks/" + nm] = np.stack([Bf[i], Bl[i]])
    for i, nm in enumerate(("u-velocity", "v-velocity", "w-velocity")):
        d["tasks/" + nm] = np.stack([Up[i], Up[i]])
    write_handler("CheckPoints", d)
