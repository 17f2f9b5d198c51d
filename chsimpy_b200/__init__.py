"""chsimpy_b200 -- B200-native (sm_100a) drop-in for the hot path of uncertaintyhub/chsimpy.

Same public names as the reference package (chsimpy/__init__.py:1-12); the views
(PlotView/MapView) are host-side matplotlib code and are not part of this package."""
from .parameters import Parameters
from .timedata import TimeData
from .solution import Solution
from .solver import Solver, BatchStepper
from .simulator import Simulator
from .cli_parser import CLIParser
from .version import __version__

__all__ = ['CLIParser', 'Solver', 'Simulator', 'Parameters', 'Solution', 'TimeData', 'BatchStepper']
