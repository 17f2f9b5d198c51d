"""Host-side scalars that feed the device stepper, plus small I/O helpers.

Mirrors the callable surface of reference chsimpy/utils.py that the hot path and the
ensemble driver use.  The thermodynamic helpers stay on sympy's `nsolve(prec=7)` on
purpose: kappa_tilde inherits that 7-digit rounding, so a
"better" root would change every spectral multiplier at the 1e-7 level."""
import functools
import importlib.util
import time
from datetime import datetime

import numpy as np

from .version import __version__  # noqa: F401


def A0(T):
    """Redlich-Kister coefficient (Kim & Sanders), reference utils.py:26-27  [kJ/mol]"""
    return 186.0575 - 0.3654 * T


def A1(T):
    """reference utils.py:30-31  [kJ/mol]"""
    return 43.7207 - 0.1401 * T


def laplace_spectrum_1d(N):
    """lambda_k = 2 cos(pi k/(N-1)) - 2, k=0..N-1: the factor table the device kernels read.
    Same expression (and therefore the same bits) as reference utils.py:35."""
    return 2 * np.cos(np.pi * (np.arange(0, N - 1 + 1)) / (N - 1)) - 2


def eigenvalues(N):
    """N x N matrix lambda_i + lambda_j (reference utils.py:34-36)."""
    lam = laplace_spectrum_1d(N)
    return lam[:, None] + lam[None, :]


def get_coefficients(N, kappa_tilde, delt, delx2):
    """(CHeig, Seig) as dense matrices, reference utils.py:39-49.  The device never reads
    these -- it regenerates them from laplace_spectrum_1d -- they exist for API parity."""
    lam1 = delt / delx2
    lam2 = kappa_tilde * lam1 / delx2
    leig = eigenvalues(N)
    return np.ones((N, N)) + lam2 * leig * leig, lam1 * leig


# ---------------------------------------------------------------------- thermodynamics
def _gibbs(c, R, T, B, A0_, A1_, log):
    return R * T * (c * (log(c) - B) + (1 - c) * log(1 - c)) + (A0_ + A1_ * (1 - 2 * c)) * c * (1 - c)


@functools.lru_cache(maxsize=4096)
def _gap_cached(R, T, B, A0_, A1_, xlower, xupper, prec):
    import sympy as sym
    xa, xb = sym.Symbol('x1', real=True), sym.Symbol('x2', real=True)
    ga = _gibbs(xa, R, T, B, A0_, A1_, sym.log)
    gb = _gibbs(xb, R, T, B, A0_, A1_, sym.log)
    dga, dgb = sym.diff(ga, xa, 1), sym.diff(gb, xb, 1)
    # common tangent: equal slopes, and the slope equals the chord
    system = (sym.Eq(dga, dgb), sym.Eq(dga, (gb - ga) / (xb - xa)))
    return sym.nsolve(system, (xa, xb), (xlower, xupper), prec=prec)


def get_miscibility_gap(R, T, B, A0, A1, xlower=0.7, xupper=0.9999, prec=7):
    """(c_A, c_B): common-tangent points of the Gibbs energy (reference utils.py:143-160)."""
    return _gap_cached(float(R), float(T), float(B), float(A0), float(A1), xlower, xupper, prec)


def get_distance_common_tangent(R, T, B, A0, A1, at):
    """Distance between G and its common tangent at composition `at` (reference
    utils.py:163-171); kappa_tilde = this / (0.1602564*64)^2 (solution.py:46)."""
    import sympy as sym
    x = sym.Symbol('x', real=True)
    G = _gibbs(x, R, T, B, A0, A1, sym.log)
    ca, cb = get_miscibility_gap(R=R, T=T, B=B, A0=A0, A1=A1)
    slope = (G.subs(x, cb) - G.subs(x, ca)) / (cb - ca)
    return np.float64((G - slope * (x - ca) - G.subs(x, ca)).subs(x, at))


def get_roots_of_EPP(R, T, A0, A1):
    """Spinodal compositions: roots of G'' in (0,1) (reference utils.py:174-180)."""
    import sympy as sym
    x = sym.Symbol('x', real=True, positive=True)
    gpp = (-2 * A0 * x ** 2 + 2 * A0 * x + 12 * A1 * x ** 3 - 18 * A1 * x ** 2 + 6 * A1 * x - R * T) / (x ** 2 - x)
    return list(sym.solveset(gpp, x, domain=sym.Interval(0, 1)))


# ---------------------------------------------------------------------- small helpers
def module_exists(name):
    return importlib.util.find_spec(name) is not None


def get_current_localtime():
    return time.strftime("%Y-%m-%d %H:%M:%S %Z", time.localtime())


def get_or_create_file_id(file_id):
    if file_id is None or str(file_id).lower() in ('auto', '', 'none'):
        return datetime.now().strftime('%d%m%Y-%H%M%S')
    return file_id


def get_number_physical_cores():
    import psutil
    return psutil.cpu_count(logical=False)


def sec_to_min_if(value, t=60):
    return (str(round(value / 60.0, 1)) + 'min') if value > t else (str(round(value, 1)) + 's')


def get_int_max_value():
    return np.iinfo(np.intp).max


def csv_export_matrix(V, fname):
    """reference utils.py:79-83"""
    if fname.endswith('bz2'):
        import pandas as pd
        pd.DataFrame(V).to_csv(fname, index=False, header=None, sep=',', compression='bz2')
    else:
        np.savetxt(fname, V, delimiter=',', fmt='%s')


def csv_import_matrix(fname):
    """reference utils.py:86-90"""
    if fname.endswith('bz2'):
        import pandas as pd
        return pd.read_csv(fname, sep=',', header=None, compression='bz2').values
    return np.loadtxt(fname, delimiter=',')


def yaml_import(fname):
    """reference utils.py:63-76: the object (Parameters or Solution) stored in a YAML file."""
    from . import yamlio
    return yamlio.load_object(fname)


def vars_to_list(obj):
    out = []
    for name in dir(obj):
        if name.startswith('_'):
            continue
        v = getattr(obj, name)
        if not callable(v):
            out.append(f"{name}, {v}")
    return out
