"""Result container: derived physical scalars + final field + TimeData.

Field names follow reference chsimpy/solution.py:17-67 (they are read by Simulator, the
views, the ensemble driver and user notebooks)."""
import numpy as np

from . import utils
from .timedata import TimeData

_TIME_SERIES = ('E', 'E2', 'SA', 'domtime', 'Ra', 'L2', 'PS', 'delt', 'it_range')


class Solution:
    def __init__(self, params=None):
        p = self.params = params
        self.U = None
        self.timedata = None
        self.Am = (25.13 * 1e6 / p.N_A) ** (2 / 3) * p.N_A        # molar area [um^2/mol], solution.py:25
        self.delx = p.L / (p.N - 1)                               # solution.py:28 (quirk Q1)
        self.delx2 = self.delx ** 2
        self.RT = p.R * p.temp
        self.BRT = p.B * p.R * p.temp
        self.Amr = 1 / self.Am
        self.A0 = p.func_A0(p.temp)
        self.A1 = p.func_A1(p.temp)
        self.time_fac = (1 / (p.M_tilde)) * p.delt
        self.M = p.M_tilde / self.Am
        if p.kappa_tilde is None:
            self.kappa_base = utils.get_distance_common_tangent(R=p.R, T=p.temp, B=p.B, A0=self.A0, A1=self.A1,
                                                                at=p.XXX)
            self.kappa_tilde = self.kappa_base / (0.1602564 * 64) ** 2         # solution.py:46
        else:
            self.kappa_tilde = p.kappa_tilde
        self.kappa = self.kappa_tilde * self.Amr
        self.restime = 0
        self.tau0 = 0
        self.t0 = 0
        self.computed_steps = 0
        self.stop_reason = 'None'
        self._eig = None

    # dense multiplier matrices only on demand (the device regenerates them from a 1-D table)
    def _multipliers(self):
        if self._eig is None:
            self._eig = utils.get_coefficients(N=self.params.N, kappa_tilde=self.kappa_tilde,
                                               delt=self.params.delt, delx2=self.delx2)
        return self._eig

    @property
    def CHeig(self):
        return self._multipliers()[0]

    @property
    def Seig(self):
        return self._multipliers()[1]

    def __getattr__(self, name):
        if name in _TIME_SERIES:
            td = self.__dict__.get('timedata')
            if td is not None:
                return getattr(td, name)
        raise AttributeError("No such attribute: " + name)

    def yaml_export_scalars(self, fname):
        from . import yamlio
        yamlio.dump_object(self, fname, tag="!Solution")

    def _scalars(self):
        skip = ('U', 'params', 'timedata', '_eig')
        return {k: v for k, v in self.__dict__.items() if k not in skip}

    def is_scalarwise_equal_with(self, other):
        return (isinstance(other, Solution) and self.params.is_scalarwise_equal_with(other.params)
                and self._scalars() == other._scalars())
