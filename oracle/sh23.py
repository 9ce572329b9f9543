"""numpy restatement of Swift_Hohenberg/FWD_Solve_SH23.py (reference alias ``SH``).  TEST INFRASTRUCTURE ONLY.

Same function names, argument order and return conventions as the reference so
that parity tests read like the reference's own driver (SH:750-784).  ``domain``
is ``oracle.fourier.Domain`` instead of a dedalus domain; ``X_FWD_DICT`` is the
same ``{'A_fwd': complex128[N/2, N_SUB_ITERS+1]}`` dictionary (SH:266-272).

Equation (SH:316-322): dt(u) + (1+dxx)^2 u - a u = 1.8 u^2 - u^3, a = -0.3, timestepper SBDF1
(SH:325): (1/dt + L_k) u^{n+1}_k = u^n_k/dt + N_k(u^n)   [D2-7], nonlinearity evaluated on the
dealiased (scale 2) grid and truncated [D2-3,4].  Parity unpinned - see oracle/__init__.py.
"""
import numpy as np
from .fourier import Domain, domain_sh23

PARAM_A = -0.3  # SH:309


def _Lk(domain):
    """(1+dxx)^2 - a in coefficient space: (1-k^2)^2 - a   (SH:316, SH:322)"""
    k = domain.kx()
    return (1.0 - k * k) ** 2 - PARAM_A


def filter_field(domain, c, frac=0.5):
    """SH:28-53 - zero coefficients whose index/size > frac (index-based, not wavenumber-based)."""
    idx = np.linspace(0, 1, domain.Nh, endpoint=False)
    c = c.copy()
    c[idx > frac] = 0j
    return c


def Integrate_Field(domain, Fgrid):
    """SH:66-87 - (1/L) integ(F) dx == mean of F over the dealiased grid [D2-5]."""
    return float(np.mean(Fgrid))


def Field_to_Vec(domain, Fgrid):
    """SH:89-128 - dealiased grid values, flattened."""
    return np.asarray(Fgrid, dtype=np.float64).reshape(-1).copy()


def Vec_to_Field(domain, Bx0):
    """SH:130-156 - reshape the vector to the dealiased grid."""
    return np.asarray(Bx0, dtype=np.float64).reshape(domain.grid_shape)


def Inner_Prod(x, y, domain, rand_arg=None):
    """SH:158-172 - mean over the 512-point grid of x_j*y_j (raw vectors, no projection)."""
    return Integrate_Field(domain, Vec_to_Field(domain, x) * Vec_to_Field(domain, y))


def GEN_BUFFER(domain, N_SUB_ITERS, Npts=256):
    """SH:238-272"""
    return {'A_fwd': np.zeros((domain.Nh, N_SUB_ITERS + 1), dtype=complex)}


def _sbdf1_step(domain, c, dt, A):
    """one SBDF1 step of SH23 from coefficient state c (SH:322-325, [D2-7])."""
    u = domain.to_grid_1d(c)
    Nhat = domain.to_coef_1d(1.8 * u * u - u * u * u)
    return (c / dt + Nhat) / A, u


def FWD_Solve_IVP_PREP(phi_grid, domain, dt=1e-02, N_ITERS=100, N_SUB_ITERS=100):
    """SH:334-407 - N_ITERS+1 SBDF1 steps (SH:378), returns the state on the dealiased grid."""
    A = 1.0 / dt + _Lk(domain)
    c = domain.to_coef_1d(phi_grid)
    for _ in range(N_ITERS + 1):
        c, _u = _sbdf1_step(domain, c, dt, A)
    return domain.to_grid_1d(c)


def Generate_IC(E_0=1.0, Npts=256, X=(0., 12. * np.pi)):
    """SH:174-236 - seed-42 noise, low-pass, normalise, 101 smoothing steps, renormalise."""
    domain = domain_sh23(Npts, X)
    rand = np.random.RandomState(seed=42)
    noise = rand.standard_normal(domain.grid_shape)
    c = filter_field(domain, domain.to_coef_1d(noise))
    phi = domain.to_grid_1d(c)
    SUM = Integrate_Field(domain, phi ** 2)
    phi = np.sqrt(E_0 / SUM) * phi
    phi = FWD_Solve_IVP_PREP(phi, domain)
    SUM = Integrate_Field(domain, phi ** 2)
    phi = np.sqrt(E_0 / SUM) * phi
    return domain, Field_to_Vec(domain, phi)


def FWD_Solve_IVP_Lin(X_k, domain, dt, N_ITERS, N_SUB_ITERS, X_FWD_DICT, filename=None, Adjoint_type="Discrete"):
    """SH:409-545.  Returns -J, J = dt * sum_{n=0}^{N_ITERS} <u^n,u^n> (SH:528-529)."""
    A = 1.0 / dt + _Lk(domain)
    c = domain.to_coef_1d(Vec_to_Field(domain, X_k[0]))   # truncation on first gather [D2-6]
    J_TRAP = 0.0
    snapshot_index = 0
    for iteration in range(N_ITERS + 1):                  # stop_iteration = N_ITERS+1 (SH:469)
        if (iteration >= (N_ITERS - N_SUB_ITERS)) and (snapshot_index <= N_SUB_ITERS):
            X_FWD_DICT['A_fwd'][:, snapshot_index] = c    # SH:499-503
            snapshot_index += 1
        c_new, u = _sbdf1_step(domain, c, dt, A)          # evaluator sees the pre-step state [D2-6]
        J_TRAP += dt * float(np.mean(u * u))              # SH:528-529
        c = c_new                                         # (the last step's result is unused)
    return (-1.) * J_TRAP


def Compatib_Cond(X_FWD_DICT, domain, dt):
    """SH:552-596 - LBVP (1/dt + L) q = -2 f, f = last snapshot.  Returns q in coefficient space."""
    f = X_FWD_DICT['A_fwd'][:, -1]
    return -2.0 * f / (1.0 / dt + _Lk(domain))


def ADJ_Solve_IVP_Lin(X_k, domain, dt, N_ITERS, N_SUB_ITERS, X_FWD_DICT, filename=None, Adjoint_type="Discrete"):
    """SH:598-729.  Returns [dJ/du0] as one float64[M] vector."""
    A = 1.0 / dt + _Lk(domain)
    if Adjoint_type == "Continuous":
        q = np.zeros(domain.Nh, dtype=complex)            # SH:646-647
        snapshot_index = -1                               # SH:656
    elif Adjoint_type == "Discrete":
        q = Compatib_Cond(X_FWD_DICT, domain, dt)         # SH:660-661
        snapshot_index = -2                               # SH:663
    else:
        raise ValueError(Adjoint_type)
    for _ in range(N_ITERS):                              # stop_iteration = N_ITERS (SH:670)
        uf = domain.to_grid_1d(X_FWD_DICT['A_fwd'][:, snapshot_index])   # SH:688
        snapshot_index -= 1
        qg = domain.to_grid_1d(q)
        rhs = domain.to_coef_1d((3.6 * uf - 3. * (uf ** 2)) * qg - 2. * uf)   # SH:640
        q = (q / dt + rhs) / A
    if Adjoint_type == "Discrete":
        # SH:702-715: dt*(a_0 q + b_0((1-a) q + 2 q_xx + q_xxxx)) == dt*(1/dt + L_k) q
        Ux0 = Field_to_Vec(domain, domain.to_grid_1d(dt * A * q))
    else:
        Ux0 = Field_to_Vec(domain, domain.to_grid_1d(q))
    return [Ux0]
