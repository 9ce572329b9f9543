"""numpy restatement of Kinematic_Dynamo/FWD_Solve_KDyn.py (reference alias ``KD``).  TEST INFRASTRUCTURE ONLY.

Same function names, argument order and return conventions as the reference
(driver block KD:1025-1067).  ``domain`` is ``oracle.fourier.Domain``; ``X_FWD_DICT`` is the same
``{'A_fwd','B_fwd','C_fwd'}`` dictionary of complex128[N/2, N-1, N-1, N_SUB_ITERS+1] (KD:347-355).

Forward problem (KD:431-440): for k != 0  div B = 0,  dt(B) - (1/Rm) Lap(B) - grad(Pi) = curl(U x B);
k = 0 mode of A,B,C,Pi is zero.  Timestepper CNAB1 (KD:443) applied to the WHOLE 4-variable pencil
system including the constraint row and Pi [D2-7]:  (M/dt + L/2) X^{n+1} = F^n + (M/dt - L/2) X^n.
With alpha = 1/dt + k^2/(2Rm), beta = 1/dt - k^2/(2Rm) the pencil solve has the closed form

    R   = F + beta*B + (i/2) k Pi
    Pi' = 2i (alpha k.B + k.R)/k^2
    B'  = (R + (i/2) k Pi')/alpha                      (=> k.B' = -k.B)

Parity unpinned - see oracle/__init__.py.
"""
import numpy as np
from .fourier import Domain, domain_kdyn


# ------------------------------------------------------------------------------------------------
# layout helpers
# ------------------------------------------------------------------------------------------------
def _K(domain):
    kx = domain.kx()[:, None, None]
    ky = domain.kc()[None, :, None]
    kz = domain.kc()[None, None, :]
    K2 = kx * kx + ky * ky + kz * kz
    shape = domain.coeff_shape
    KX = np.broadcast_to(kx, shape)
    KY = np.broadcast_to(ky, shape)
    KZ = np.broadcast_to(kz, shape)
    return KX, KY, KZ, K2


def _curl_c(K, a):
    """ik x a in coefficient space (KD:421-423, KD:841-843)"""
    KX, KY, KZ, _ = K
    return [1j * (KY * a[2] - KZ * a[1]),
            1j * (KZ * a[0] - KX * a[2]),
            1j * (KX * a[1] - KY * a[0])]


def _cross(a, b):
    """pointwise a x b on the grid (KD:417-419, KD:846-848)"""
    return [a[1] * b[2] - a[2] * b[1],
            a[2] * b[0] - a[0] * b[2],
            a[0] * b[1] - a[1] * b[0]]


def _to_grid3(domain, c):
    return [domain.to_grid_3d(ci) for ci in c]


def _to_coef3(domain, g):
    return [domain.to_coef_3d(gi) for gi in g]


def _cnab_step(K, Pi, X, F, alpha, beta):
    """closed-form solve of the CNAB1 4x4 pencil (module docstring; [D2-7]); k = 0 entries -> 0 (KD:437-440)."""
    KX, KY, KZ, K2 = K
    Kc = (KX, KY, KZ)
    R = [F[i] + beta * X[i] + 0.5j * Kc[i] * Pi for i in range(3)]
    kX = KX * X[0] + KY * X[1] + KZ * X[2]
    kR = KX * R[0] + KY * R[1] + KZ * R[2]
    K2s = np.where(K2 == 0.0, 1.0, K2)
    Pin = 2.0j * (alpha * kX + kR) / K2s
    Xn = [(R[i] + 0.5j * Kc[i] * Pin) / alpha for i in range(3)]
    zero = (K2 == 0.0)
    Pin = np.where(zero, 0.0, Pin)
    Xn = [np.where(zero, 0.0, x) for x in Xn]
    return Pin, Xn


def _project(K, v):
    """P_k v = v - k (k.v)/k^2, k = 0 -> 0"""
    KX, KY, KZ, K2 = K
    K2s = np.where(K2 == 0.0, 1.0, K2)
    kv = (KX * v[0] + KY * v[1] + KZ * v[2]) / K2s
    out = [v[0] - KX * kv, v[1] - KY * kv, v[2] - KZ * kv]
    return [np.where(K2 == 0.0, 0.0, o) for o in out]


# ------------------------------------------------------------------------------------------------
# general routines
# ------------------------------------------------------------------------------------------------
def filter_field(domain, c, frac=0.25):
    """KD:30-55 - zero coefficients whose index/size > frac on ANY axis of the (N/2,N-1,N-1) array.
    Index based: on the c2c axes this removes every negative wavenumber."""
    c = c.copy()
    filt = np.zeros(domain.coeff_shape, dtype=bool)
    for i, n in enumerate(domain.coeff_shape):
        idx = np.linspace(0, 1, n, endpoint=False)
        sh = [1, 1, 1]
        sh[i] = n
        filt = filt | (idx.reshape(sh) > frac)
    c[filt] = 0j
    return c


def Integrate_Field(domain, Fgrid):
    """KD:68-89 - (1/V) integ(F) dV == mean of F over the dealiased M^3 grid [D2-5]."""
    return float(np.mean(Fgrid))


def Field_to_Vec(domain, Fx, Fy, Fz):
    """KD:91-137 - concatenate the flattened dealiased-grid components."""
    return np.concatenate((np.asarray(Fx).reshape(-1), np.asarray(Fy).reshape(-1), np.asarray(Fz).reshape(-1)))


def Vec_to_Field(domain, Bx0):
    """KD:139-171 - np.split into 3 and reshape to (M,M,M)."""
    a1, a2, a3 = np.split(np.asarray(Bx0, dtype=np.float64), 3)
    g = domain.grid_shape
    return [a1.reshape(g), a2.reshape(g), a3.reshape(g)]


def Inner_Prod_3(x, y, domain, random_arg=None):
    """KD:173-181 - (1/M^3) sum over all 3M^3 entries of x_j y_j (raw vectors)."""
    X = Vec_to_Field(domain, x)
    Y = Vec_to_Field(domain, y)
    return Integrate_Field(domain, X[0] * Y[0] + X[1] * Y[1] + X[2] * Y[2])


def GEN_BUFFER(Npts, domain, N_SUB_ITERS):
    """KD:319-355"""
    shape = domain.coeff_shape + (N_SUB_ITERS + 1,)
    return {'A_fwd': np.zeros(shape, dtype=complex),
            'B_fwd': np.zeros(shape, dtype=complex),
            'C_fwd': np.zeros(shape, dtype=complex)}


# ------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------
class _Fwd:
    """FWD_Solve_Build_Lin (KD:362-450): U parameter fields + CNAB1 coefficients."""

    def __init__(self, domain, Rm, dt, Ux0):
        self.domain = domain
        self.K = _K(domain)
        K2 = self.K[3]
        self.alpha = 1.0 / dt + K2 / (2.0 * Rm)
        self.beta = 1.0 / dt - K2 / (2.0 * Rm)
        # parameter fields are moved to coefficient space by the evaluator => projected on the
        # retained modes at first use [D2-8]
        self.Ug = _to_grid3(domain, _to_coef3(domain, Vec_to_Field(domain, Ux0)))

    def rhs(self, Bc):
        Bg = _to_grid3(self.domain, Bc)
        E = _cross(self.Ug, Bg)                              # EMF = U x B   (KD:417-419)
        return _curl_c(self.K, _to_coef3(self.domain, E))    # curl(EMF)     (KD:421-423)

    def step(self, Pi, Bc):
        return _cnab_step(self.K, Pi, Bc, self.rhs(Bc), self.alpha, self.beta)


def FWD_Solve_IVP_Prep(Bx0, Ux0, domain, Rm, dt, N_ITERS):
    """KD:452-527 - N_ITERS+1 CNAB1 steps (KD:510); returns the three coefficient arrays."""
    S = _Fwd(domain, Rm, dt, Ux0)
    Bc = _to_coef3(domain, Vec_to_Field(domain, Bx0))
    Pi = np.zeros(domain.coeff_shape, dtype=complex)
    for _ in range(N_ITERS + 1):
        Pi, Bc = S.step(Pi, Bc)
    return Bc


def Generate_IC(Npts, X=(0., 2. * np.pi), M_0=1.0, U_Noise=False, Rm=1.0, dt=5e-04):
    """KD:183-317.  ``Rm`` and ``dt`` stand for the module-level globals that the reference's smoothing
    step uses instead of Rm_IC/dt_IC (KD:299-302 quirk); defaults are the literals at KD:1028-1029."""
    domain = domain_kdyn(Npts, X)
    K = _K(domain)

    def curl_of_noise():
        rand = np.random.RandomState(seed=42)
        noise = rand.standard_normal(domain.grid_shape)
        phi = filter_field(domain, domain.to_coef_3d(noise))
        px = domain.to_grid_3d(1j * K[0] * phi)
        py = domain.to_grid_3d(1j * K[1] * phi)
        pz = domain.to_grid_3d(1j * K[2] * phi)
        return [py - pz, pz - px, px - py]                   # KD:241-243

    B = curl_of_noise()
    if U_Noise is False:
        M = domain.M
        g = domain.interval[0] + domain.L * np.arange(M) / M
        x = g[:, None, None]; y = g[None, :, None]; z = g[None, None, :]
        one = np.ones((M, M, M))
        U = [0.5 * np.sin(y) * np.cos(z) / np.sqrt(3.) * one,    # KD:258-260
             0.5 * np.sin(z) * np.cos(x) / np.sqrt(3.) * one,
             0.5 * np.sin(x) * np.cos(y) / np.sqrt(3.) * one]
        # fields assigned at scale 1 in the reference; band-limited so the dealiased values are exact.
    else:
        U = curl_of_noise()                                   # same seed (KD:272)
    SUM = Integrate_Field(domain, U[0] ** 2 + U[1] ** 2 + U[2] ** 2)
    U = [np.sqrt(1. / SUM) * u for u in U]                    # KD:287-292

    N_ITERS = 100
    Bx0 = Field_to_Vec(domain, *B)
    Ux0 = Field_to_Vec(domain, *U)
    Bc = FWD_Solve_IVP_Prep(Bx0, Ux0, domain, Rm, dt, N_ITERS)   # KD:302 (globals Rm, dt)
    B = _to_grid3(domain, Bc)
    SUM = Integrate_Field(domain, B[0] ** 2 + B[1] ** 2 + B[2] ** 2)
    B = [np.sqrt(M_0 / SUM) * b for b in B]                   # KD:308-310
    return domain, Field_to_Vec(domain, *B), Field_to_Vec(domain, *U)


def FWD_Solve_IVP_Lin(X0, domain, Rm, dt, N_ITERS, N_SUB_ITERS, X_FWD_DICT, Cost_function="Final", Adjoint_type="Discrete"):
    """KD:529-689.  Returns -J (KD:689)."""
    Bx0, Ux0 = X0[0], X0[1]
    S = _Fwd(domain, Rm, dt, Ux0)
    Bc = _to_coef3(domain, Vec_to_Field(domain, Bx0))
    Pi = np.zeros(domain.coeff_shape, dtype=complex)
    J_TRAP = 0.
    snapshot_index = 0
    for iteration in range(N_ITERS + 1):                      # stop_iteration = N_ITERS+1 (KD:597)
        if (iteration >= (N_ITERS - N_SUB_ITERS)) and (snapshot_index <= N_SUB_ITERS):
            X_FWD_DICT['A_fwd'][:, :, :, snapshot_index] = Bc[0]   # KD:635-637
            X_FWD_DICT['B_fwd'][:, :, :, snapshot_index] = Bc[1]
            X_FWD_DICT['C_fwd'][:, :, :, snapshot_index] = Bc[2]
            snapshot_index += 1
        # flow property 'J(B)' is evaluated on the pre-step state [D2-6]
        need_J = (Cost_function == "Integrated") or (Cost_function == "Final" and iteration == N_ITERS)
        if need_J:
            Bg = _to_grid3(domain, Bc)
            JB = Integrate_Field(domain, Bg[0] * Bg[0] + Bg[1] * Bg[1] + Bg[2] * Bg[2])
            if Cost_function == "Integrated":
                J_TRAP += dt * JB                              # KD:668-669
            else:
                J_TRAP = JB                                    # KD:671-673
        if iteration < N_ITERS:                                # the (N_ITERS+1)-th step's result is unused
            Pi, Bc = S.step(Pi, Bc)
    return (-1.) * J_TRAP


# ------------------------------------------------------------------------------------------------
# adjoint
# ------------------------------------------------------------------------------------------------
def Compatib_Cond(X_FWD_DICT, domain, Rm, dt, Cost_function="Final"):
    """KD:696-764 - LBVP for the terminal adjoint state; returns coefficient arrays."""
    K = _K(domain)
    K2 = K[3]
    f = [X_FWD_DICT['A_fwd'][:, :, :, -1], X_FWD_DICT['B_fwd'][:, :, :, -1], X_FWD_DICT['C_fwd'][:, :, :, -1]]
    P = _project(K, [-2. * fi for fi in f])
    if Cost_function == "Final":
        den = 1.0 + dt * (.5 / Rm) * K2                       # KD:733-735
    elif Cost_function == "Integrated":
        den = 1.0 / dt + (.5 / Rm) * K2                       # KD:739-741
    else:
        raise ValueError(Cost_function)
    return [p / den for p in P]


def ADJ_Solve_IVP_Lin(X0, domain, Rm, dt, N_ITERS, N_SUB_ITERS, X_FWD_DICT, Cost_function="Final", Adjoint_type="Discrete"):
    """KD:766-1004.  Returns [dJ/dB0, dJ/dU] as two float64[3M^3] vectors."""
    Bx0, Ux0 = X0[0], X0[1]
    K = _K(domain)
    K2 = K[3]
    alpha = 1.0 / dt + K2 / (2.0 * Rm)
    beta = 1.0 / dt - K2 / (2.0 * Rm)
    Ug = _to_grid3(domain, _to_coef3(domain, Vec_to_Field(domain, Ux0)))      # [D2-8]

    if Adjoint_type == "Continuous":
        G = [-2. * X_FWD_DICT[key][:, :, :, -1] for key in ('A_fwd', 'B_fwd', 'C_fwd')]   # KD:906-908
        snapshot_index = -1
    elif Adjoint_type == "Discrete":
        G = Compatib_Cond(X_FWD_DICT, domain, Rm, dt, Cost_function)          # KD:914-918
        snapshot_index = -2
    else:
        raise ValueError(Adjoint_type)
    Pi = np.zeros(domain.coeff_shape, dtype=complex)
    P = np.zeros(domain.coeff_shape, dtype=complex)
    nu = [np.zeros(domain.coeff_shape, dtype=complex) for _ in range(3)]
    inv_dt = 1.0 / dt

    for _ in range(N_ITERS):                                   # stop_iteration = N_ITERS (KD:929)
        Bf_c = [X_FWD_DICT[key][:, :, :, snapshot_index] for key in ('A_fwd', 'B_fwd', 'C_fwd')]   # KD:955-957
        snapshot_index -= 1
        Bf = _to_grid3(domain, Bf_c)
        W = _to_grid3(domain, _curl_c(K, G))                   # curl G            (KD:841-843)
        FG = _to_coef3(domain, _cross(W, Ug))                  # (curl G) x U      (KD:857-859)
        if Cost_function == "Integrated":
            FG = [FG[i] - 2. * Bf_c[i] for i in range(3)]      # KD:862-864
        Fnu = _to_coef3(domain, [-w for w in _cross(W, Bf)])   # -(curl G) x B_f   (KD:875-877)
        # nu-system: dt(nu) + grad(P) = Fnu, div nu = 0, CNAB1 with no diffusion: alpha = beta = 1/dt.
        # L-row sign: +grad(P) here vs -grad(Pi) for G; the closed form is symmetric in that sign
        # once P is eliminated, so the same solver is used with P' = -Pi'.
        Pn, nu = _cnab_step(K, -P, nu, Fnu, inv_dt, inv_dt)
        P = -Pn
        Pi, G = _cnab_step(K, Pi, G, FG, alpha, beta)

    if Adjoint_type == "Discrete":
        # KD:979-989: dt*(G/dt - (.5/Rm) Lap G) == dt*alpha*G
        Bx0 = Field_to_Vec(domain, *_to_grid3(domain, [dt * alpha * g for g in G]))
    else:
        Bx0 = Field_to_Vec(domain, *_to_grid3(domain, G))
    Ux0 = Field_to_Vec(domain, *_to_grid3(domain, nu))          # KD:995
    return [Bx0, Ux0]
