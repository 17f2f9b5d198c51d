"""Simulation parameters: the input contract of the stepper.

Attribute names, defaults and semantics follow reference chsimpy/parameters.py:24-64
(they are the public API that CLI, notebooks and the ensemble driver set directly).
YAML import/export uses PyYAML when ruamel.yaml is not installed."""
import copy

from . import utils
from .version import __reference_version__

# name -> default, in the reference's order (parameters.py:24-61)
_DEFAULTS = (
    ("seed", 2023),
    ("N", 512),                          # pixels per side
    ("L", 2),                            # domain length [um]
    ("XXX", 0.875),                      # mean initial mole fraction
    ("temp", 650 + 273.15),              # [K]
    ("B", 12.86),                        # Gibbs-energy tuning parameter (Charles 1967)
    ("R", 0.0083144626181532),           # gas constant [kJ/(K mol)]
    ("N_A", 6.02214076e+23),             # Avogadro
    ("delt", 3e-8),
    ("delt_max", 9e-8),
    ("M_tilde", 1.71e-8),                # mobility factor [um^2/(kJ s)]
    ("kappa_tilde", None),               # None -> from the common-tangent distance
    ("threshold", 0.875),                # component split for SA (== XXX by default)
    ("ntmax", int(1e6)),
    ("export_csv", None),
    ("png", False),
    ("png_anim", False),
    ("yaml", False),
    ("no_gui", False),
    ("file_id", "auto"),
    ("full_sim", False),
    ("compress_csv", False),
    ("time_max", None),                  # minutes of simulated time
    ("generator", "uniform"),            # uniform | sobol | lcg | simplex
    ("adaptive_time", False),
    ("jitter", None),
    ("update_every", 100),
    ("no_diagrams", False),
    ("Uinit_file", None),
)
_NON_SCALAR = ("func_A0", "func_A1")


class Parameters:
    version = __reference_version__

    def __init__(self):
        for name, default in _DEFAULTS:
            setattr(self, name, default)
        self.func_A0 = lambda temp: utils.A0(temp)
        self.func_A1 = lambda temp: utils.A1(temp)

    # -- comparison / copy ---------------------------------------------------------------
    def _scalars(self):
        return {k: v for k, v in self.__dict__.items() if k not in _NON_SCALAR and k != "version"}

    def is_scalarwise_equal_with(self, other):
        return isinstance(other, Parameters) and self._scalars() == other._scalars()

    def __eq__(self, other):
        return isinstance(other, Parameters) and self.__dict__ == other.__dict__

    def deepcopy(self):
        return copy.deepcopy(self)

    def __str__(self):
        return str(dict(sorted(self._scalars().items())))

    # -- YAML (scalars only) -------------------------------------------------------------
    def yaml_export_scalars(self, fname):
        from . import yamlio
        yamlio.dump_object(self, fname, tag="!Parameters")

    def yaml_import_scalars(self, fname):
        from . import yamlio
        for k, v in yamlio.load_mapping(fname).items():
            if hasattr(self, k) and not callable(v) and k not in _NON_SCALAR:
                setattr(self, k, v)
