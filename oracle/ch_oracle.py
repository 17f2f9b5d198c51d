"""CPU oracle for the chsimpy hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/scipy restatement of the reference's semi-implicit spectral Cahn-Hilliard
stepper.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this module; the product path
(`chsimpy_b200.Solver`, `libchs_b200.so`) never does and has no CPU fallback.

Parity status: PINNED.  The reference holds no golden vector for the solver path
(only the LCG known-answer test, reference tests/test.py:19-37), so this oracle is
pinned against the reference itself, run unmodified in the build container through
`oracle/ref_shim.py`: `oracle/validate_oracle.py` checks the two bit-for-bit, and
`tests/golden/make_golden.py` freezes reference outputs into `tests/golden/*.npz`,
which `tests/test_oracle.py` replays on every CPU test run.

Every function cites the reference lines it follows (paths relative to
/root/reference/).  The arithmetic keeps the reference's operation order so that the
results are bit-identical under the same numpy/scipy build.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np
import scipy.fftpack as _fftpack
from scipy.stats import qmc as _qmc
from threadpoolctl import threadpool_limits as _threadpool_limits

# TimeData column order, chsimpy/timedata.py:8-9
COLS = ("it", "E", "E2", "SA", "domtime", "Ra", "L2", "PS", "delt")


# --------------------------------------------------------------------------- scalars
def redlich_kister_A0(T):
    """chsimpy/utils.py:26-27"""
    return 186.0575 - 0.3654 * T


def redlich_kister_A1(T):
    """chsimpy/utils.py:30-31"""
    return 43.7207 - 0.1401 * T


def laplacian_eigs(N):
    """chsimpy/utils.py:34-36 -- lambda_i + lambda_j with the (N-1)-point spectrum (quirk Q1)."""
    lam = 2 * np.cos(np.pi * (np.arange(0, N - 1 + 1)) / (N - 1)) - 2
    return lam.reshape(N, 1) @ np.ones((1, N)) + np.ones((N, 1)) @ lam.reshape(1, N)


def spectral_multipliers(N, kappa_tilde, delt, delx2):
    """chsimpy/utils.py:39-49 -> (CHeig, Seig)."""
    lam1 = delt / delx2
    lam2 = kappa_tilde * lam1 / delx2
    leig = laplacian_eigs(N)
    CHeig = np.ones((N, N)) + lam2 * leig * leig
    Seig = lam1 * leig
    return CHeig, Seig


def lcg_field(n1, n2, seed):
    """chsimpy/mport.py:8-32 -- float64 BSD LCG, column-major fill, /(m-1)."""
    a = np.float64(1103515245)
    c = np.float64(12345)
    m = np.float64(2 ** 31)
    x = seed
    out = np.zeros((n1, n2))
    for i in range(n1 * n2):
        x = (a * x + c) % m
        out[int(i % n1), int(i / n1)] = x
    out /= (m - 1)
    return out


@dataclasses.dataclass
class Consts:
    """Derived scalars of chsimpy/solution.py:25-50 (kappa_tilde must be supplied:
    the sympy common-tangent solve of solution.py:39-46 is host code shared with the
    product and is validated separately)."""
    N: int
    L: float
    delx: float
    delx2: float
    RT: float
    BRT: float
    B: float
    Amr: float
    A0: float
    A1: float
    kappa_tilde: float
    M_tilde: float
    threshold: float
    delt0: float
    delt_max: float

    @staticmethod
    def from_params(N=512, L=2, temp=923.15, B=12.86, R=0.0083144626181532,
                    N_A=6.02214076e+23, delt=3e-8, delt_max=9e-8, M_tilde=1.71e-8,
                    threshold=0.875, kappa_tilde=None, A0=None, A1=None):
        Am = (25.13 * 1e6 / N_A) ** (2 / 3) * N_A          # solution.py:25
        delx = L / (N - 1)                                 # solution.py:28
        if A0 is None:
            A0 = redlich_kister_A0(temp)
        if A1 is None:
            A1 = redlich_kister_A1(temp)
        assert kappa_tilde is not None
        return Consts(N=N, L=L, delx=delx, delx2=delx ** 2, RT=R * temp, BRT=B * R * temp,
                      B=B, Amr=1 / Am, A0=A0, A1=A1, kappa_tilde=kappa_tilde,
                      M_tilde=M_tilde, threshold=threshold, delt0=delt, delt_max=delt_max)


# --------------------------------------------------------------------------- fields
def initial_field(N, c0, generator="uniform", seed=2023):
    """chsimpy/solver.py:56-82 -> (U_init, create_rand or None)."""
    if generator == "lcg":
        return c0 + (c0 * 0.01 * lcg_field(N, N, seed)), None       # un-centred, solver.py:66
    if generator == "sobol":
        q = _qmc.Sobol(d=N, seed=seed)
        draw = lambda n: q.random(n)                               # solver.py:70-71
    elif generator == "uniform":
        g = np.random.Generator(np.random.PCG64(seed))
        draw = lambda n: g.random((n, n))                          # solver.py:78-79
    else:
        raise ValueError("generator not available in the oracle: " + generator)
    return c0 + (c0 * 0.01 * (draw(N) - 0.5)), draw                # solver.py:81-82


def chemical_potential(U, k: Consts):
    """chsimpy/solver.py:166-175"""
    Uinv = 1 - U
    ratio = U / Uinv
    d = Uinv - U
    return np.real(k.RT * np.log(ratio) - k.BRT + (k.A0 + k.A1 * d) * d - 2 * k.A1 * U * Uinv)


def energies(U, k: Consts):
    """chsimpy/solver.py:213-221 (== :100-111 in prepare) -> (E, E2)."""
    gx, gy = np.gradient(U, k.delx, axis=[0, 1], edge_order=1)
    g2 = gx ** 2 + gy ** 2
    Uinv = 1 - U
    E2 = 0.5 * k.Amr * k.kappa_tilde * k.L ** 2 * np.mean(g2)
    E = k.Amr * k.L ** 2 * np.mean(
        np.real(k.RT * (U * (np.log(U) - k.B) + Uinv * np.log(Uinv))
                + (k.A0 + k.A1 * (Uinv - U)) * U * Uinv)) + E2
    return E, E2


def roughness_stats(U, N):
    """chsimpy/solver.py:223-226 -> (PS, Ra)."""
    Um = U - np.mean(U)
    PS = np.sum(np.abs(Um)) / (N ** 2)
    r = int(N / 2) + 1
    Ra = np.mean(np.abs(U[r, :] - np.mean(U[r, :])))
    return PS, Ra


class OracleSolver:
    """State machine equivalent to chsimpy.Solver (solver.py:45-252) + TimeData
    (timedata.py) for one simulation.  `rows` is the (n, 9) TimeData table."""

    def __init__(self, k: Consts, U_init, *, full_sim=False, adaptive_time=False,
                 jitter=None, time_max=None, create_rand=None):
        self.k = k
        self.U_init = U_init
        self.full_sim = full_sim
        self.adaptive_time = adaptive_time
        self.jitter = jitter
        self.time_max = time_max
        self.create_rand = create_rand
        # solver.py:50-54 -- persistent across prepare()/solve calls (quirk Q16)
        self.skip_check = False
        self.time_delta_sum = 0.0
        self.time_passed = 0.0
        self.delt = k.delt0
        self.prepared = False
        self.CHeig, self.Seig = spectral_multipliers(k.N, k.kappa_tilde, k.delt0, k.delx2)  # solution.py:52-55

    # solver.py:84-135
    def prepare(self):
        k = self.k
        U = self.U_init.copy()
        assert U.shape == (k.N, k.N)
        E, E2 = energies(U, k)
        PS, Ra = roughness_stats(U, k.N)
        self.rows = np.empty((0, 9))
        self._append(0, E, E2, 0, 0, Ra, 0, PS, self.delt)       # SA=0, domtime=0, L2=0 (quirk Q14)
        self.U = U
        self.tau0 = 0.0
        self.t0 = 0.0
        self.stop_reason = "None"
        self.computed_steps = 1
        self.prepared = True

    # timedata.py:8-10
    def _append(self, it, E, E2, SA, domtime, Ra, L2, PS, delt):
        self.rows = np.append(self.rows, [[it, E, E2, SA, domtime, Ra, L2, PS, delt]], axis=0)
        assert not np.any(np.isnan(self.rows[-1]))

    # timedata.py:51-63
    def _energy_falls(self, it):
        E2 = self.rows[:, 2]
        return E2[it - 1] > E2[it] > E2[0]

    def run(self, nsteps):
        """solver.py:137-252, under the single-thread BLAS cap the reference applies around
        Simulator.solve (simulator.py:14,36): np.linalg.norm -> BLAS ddot is thread-count
        dependent in its last bit."""
        with _threadpool_limits(limits=1, user_api="blas"):
            return self._run(nsteps)

    def _run(self, nsteps):
        assert self.prepared is True
        k = self.k
        N = k.N
        limit = None
        if self.time_max is not None and self.time_max > 0:
            limit = self.time_max * 60
        # NOTE solver.py:151-152 re-reads the *initial* multipliers from the Solution on
        # every call; an adaptive-dt update (solver.py:189-193) only lives in locals.
        CHeig, Seig = self.CHeig, self.Seig
        U = self.U
        hat_U = _fftpack.dctn(U, norm="ortho")                      # solver.py:159 (every call, Q2)
        first = 1 if self.computed_steps == 1 else 0                # solver.py:160-163 (Q3)
        for _ in range(first, nsteps):
            mu = chemical_potential(U, k)
            if self.adaptive_time and self.computed_steps > 500 and np.remainder(self.computed_steps, 2) == 0:
                alpha = 500 / (2) ** 3
                dyn = np.linalg.norm(k.delt_max / np.sqrt(1 + alpha * np.abs(mu) ** 2), ord=-1)
                new = max(k.delt0, dyn)
                if new / self.delt > 1.15:
                    self.delt = 0.75 * self.delt + 0.25 * new
                else:
                    self.delt = new
                CHeig, Seig = spectral_multipliers(N, k.kappa_tilde, self.delt, k.delx2)
            self.time_delta_sum += self.delt                        # solver.py:195 (Q15)
            self.time_passed = self.time_delta_sum / k.M_tilde
            if limit is not None and self.time_passed > limit:
                self.stop_reason = "time-limit"
                break
            hat_rhs = hat_U + Seig * _fftpack.dctn(mu, norm="ortho")  # solver.py:201
            hat_U = hat_rhs / CHeig                                 # solver.py:206
            U = _fftpack.idctn(hat_U, norm="ortho")                 # solver.py:208
            if self.jitter is not None and 0.0 < self.jitter < 0.1:
                U += self.jitter * (2 * self.create_rand(N) - 1)    # solver.py:210-211
            E, E2 = energies(U, k)
            PS, Ra = roughness_stats(U, N)
            L2 = np.linalg.norm(mu) / N ** 2                        # solver.py:225 (pre-update mu)
            SA = np.sum(U < k.threshold) / (N ** 2)                 # solver.py:228
            domtime = self.time_passed ** (1 / 3)
            self._append(self.computed_steps, E, E2, SA, domtime, Ra, L2, PS, self.delt)
            self.computed_steps += 1
            if not self.skip_check and self._energy_falls(self.computed_steps - 1):
                self.tau0 = self.computed_steps
                self.t0 = self.time_passed
                if not self.full_sim:
                    self.stop_reason = "energy"
                    break
                self.skip_check = True
        self.U = U
        return self


# --------------------------------------------------------------------------- host scalars
def kappa_tilde_from_common_tangent(R, T, B, A0, A1, at):
    """chsimpy/utils.py:143-171 + solution.py:46, via sympy (7-digit nsolve, quirk Q12).
    Imported lazily: sympy is slow to import."""
    import sympy as sym
    x1, x2, x = sym.symbols("x1 x2 x", real=True)

    def G(c):
        return R * T * (c * (sym.log(c) - B) + (1 - c) * sym.log(1 - c)) + (A0 + A1 * (1 - 2 * c)) * c * (1 - c)

    y1, y2 = G(x1), G(x2)
    d1, d2 = sym.diff(y1, x1, 1), sym.diff(y2, x2, 1)
    ca, cb = sym.nsolve((sym.Eq(d1, d2), sym.Eq(d1, (y2 - y1) / (x2 - x1))), (x1, x2), (0.7, 0.9999), prec=7)
    Ex = G(x)
    m = (Ex.subs(x, cb) - Ex.subs(x, ca)) / (cb - ca)
    dist = (Ex - m * (x - ca) - Ex.subs(x, ca)).subs(x, at)
    return float(np.float64(dist)) / (0.1602564 * 64) ** 2, (ca, cb)


def run_default(N=512, nsteps=None, seed=2023, generator="uniform", c0=0.875, fac_A0=1.0, fac_A1=1.0,
                temp=923.15, kappa_tilde=None, **kw):
    """Convenience: builds Consts + OracleSolver for the reference defaults
    (parameters.py:24-64) and runs it.  Used by tests and by bench.py's CPU baseline."""
    R, B = 0.0083144626181532, 12.86
    A0 = redlich_kister_A0(temp) * fac_A0
    A1 = redlich_kister_A1(temp) * fac_A1
    if kappa_tilde is None:
        kappa_tilde, _ = kappa_tilde_from_common_tangent(R, temp, B, A0, A1, c0)
    solver_kw = {n: kw.pop(n) for n in ("full_sim", "adaptive_time", "jitter", "time_max") if n in kw}
    k = Consts.from_params(N=N, temp=temp, A0=A0, A1=A1, kappa_tilde=kappa_tilde, threshold=c0, **kw)
    U0, draw = initial_field(N, c0, generator, seed)
    s = OracleSolver(k, U0, create_rand=draw, **solver_kw)
    s.prepare()
    if nsteps is None:
        nsteps = int(1e6)
    return s.run(nsteps)
