"""Command line -> Parameters.  Flag names, defaults and range checks follow reference
chsimpy/cli_parser.py:24-161 so that shell scripts written for `chsimpy` keep working."""
import argparse

from . import parameters

_SIM = (
    (('-N',), dict(default=512, type=int, help='Number of pixels in one domain (NxN)')),
    (('-n', '--ntmax'), dict(default=int(1e6), type=int, help='Maximum number of simulation steps')),
    (('-t', '--time-max'), dict(type=float, help='Maximal time in minutes to simulate (ignores ntmax)')),
    (('-z', '--full-sim'), dict(action='store_true', help='Do not stop early when energy falls')),
    (('-a', '--adaptive-time'), dict(action='store_true', help='Adaptive time stepping (experimental)')),
    (('--cinit',), dict(type=float, default=0.875, help='Initial mean mole fraction of silica')),
    (('--threshold',), dict(type=float, default=0.875, help='Threshold mole fraction for c_A / c_B')),
    (('--temperature',), dict(type=float, default=923.15, help='Temperature in Kelvin')),
    (('--A0',), dict(type=float, help='A0 value (ignores temperature) [kJ / mol]')),
    (('--A1',), dict(type=float, help='A1 value (ignores temperature) [kJ / mol]')),
    (('-K', '--kappa-tilde'), dict(type=float, help='Value for kappa_tilde [kJ/mol]')),
    (('--dt',), dict(type=float, default=3e-8, help='Time delta of simulation')),
    (('-g', '--generator'), dict(choices=['uniform', 'simplex', 'sobol', 'lcg'], default='uniform',
                                 help='Generator for the initial random deviations')),
    (('-s', '--seed'), dict(default=2023, type=int, help='Start seed for random number generators')),
    (('-j', '--jitter'), dict(type=float, help='Per-step noise factor in [0, 0.1)')),
)
_INPUT = (
    (('-p', '--parameter-file'), dict(help='Input yaml file with parameter values (overrides CLI)')),
    (('--Uinit-file',), dict(help='Initial U matrix file (csv or csv.bz2)')),
)
_OUTPUT = (
    (('-f', '--file-id'), dict(default='auto', help='File id for outputs ("auto" = timestamp)')),
    (('--no-gui',), dict(action='store_true', help='Do not show a plot window')),
    (('--png',), dict(action='store_true', help='Export solution plot to PNG')),
    (('--png-anim',), dict(action='store_true', help='Export live plotting to a series of PNGs')),
    (('--yaml',), dict(action='store_true', help='Export parameters/solution scalars to yaml')),
    (('--export-csv',), dict(help='Solution members to export as csv (e.g. "U,E2")')),
    (('-C', '--compress-csv'), dict(action='store_true', help='Compress csv files with bz2')),
    (('--update-every',), dict(type=int, help='Every n steps the state is handed to the view (>=2)')),
    (('--no-diagrams',), dict(action='store_true', help='Only render the image map of U')),
)


class CLIParser:
    def __init__(self, progname='chsimpy'):
        self.parser = argparse.ArgumentParser(
            prog=progname, add_help=True, formatter_class=argparse.ArgumentDefaultsHelpFormatter,
            description='Phase separation in Na2O-SiO2 glasses (Cahn-Hilliard) on NVIDIA B200')
        self.parser.add_argument('--version', action='version',
                                 version=f"%(prog)s {parameters.Parameters.version}")
        for title, spec in (('Simulation', _SIM), ('Input', _INPUT), ('Output', _OUTPUT)):
            g = self.parser.add_argument_group(title)
            for flags, kw in spec:
                g.add_argument(*flags, **kw)
        self.args = None

    def get_parameters(self, argv=None):
        a = self.args = self.parser.parse_args(argv)
        p = parameters.Parameters()
        for dst, src in (('ntmax', 'ntmax'), ('N', 'N'), ('file_id', 'file_id'), ('seed', 'seed'),
                         ('full_sim', 'full_sim'), ('compress_csv', 'compress_csv'), ('export_csv', 'export_csv'),
                         ('png', 'png'), ('png_anim', 'png_anim'), ('yaml', 'yaml'), ('no_gui', 'no_gui'),
                         ('adaptive_time', 'adaptive_time'), ('time_max', 'time_max'), ('generator', 'generator'),
                         ('jitter', 'jitter'), ('update_every', 'update_every'), ('no_diagrams', 'no_diagrams'),
                         ('Uinit_file', 'Uinit_file')):
            setattr(p, dst, getattr(a, src))
        if a.kappa_tilde is not None:
            p.kappa_tilde = a.kappa_tilde
        p.XXX = self.get_if_range_ok(a.cinit, 0.85, 0.95, 'cinit')
        p.threshold = self.get_if_range_ok(a.threshold, 0.85, 0.95, 'threshold')
        p.delt = self.get_if_range_ok(a.dt, 1e-12, 1e-6, 'dt')
        if a.temperature is not None:
            p.temp = a.temperature
        err = self.parser.error
        if p.update_every is not None and p.update_every < 2:
            err('--update-every should be >=2')
        if p.png_anim and p.update_every is None:
            err("--png-anim requires --update-every.")
        if p.export_csv is not None and p.export_csv.lower() in ('', 'none'):
            err("--export-csv does not contain valid entries.")
        if p.compress_csv and p.export_csv is None:
            err("--compress-csv has no effect (no --export-csv given).")
        if a.parameter_file is not None:
            p.yaml_import_scalars(a.parameter_file)
        if a.A0 is not None:
            p.func_A0 = lambda T: a.A0
        if a.A1 is not None:
            p.func_A1 = lambda T: a.A1
        return p

    def print_info(self):
        print(f"{self.parser.prog} {parameters.Parameters.version} ('--help' for command parameters)")

    def get_if_range_ok(self, value, lower, upper, name=None):
        if lower <= value <= upper:
            return value
        self.parser.error(f"{name or 'value'} is out of the range [{lower},{upper}].")
