"""Command line of the package: `python -m chsimpy_b200 [flags]` takes the reference's flags
(chsimpy/cli_parser.py) and prints the same closing lines as its entry point
(chsimpy/__main__.py:8-25): the parameter dump, the stop summary and, if anything was exported,
the file id.  `run(argv)` returns the Solution so that tests can drive the CLI in-process."""
import sys

from . import cli_parser, simulator, utils


def _stop_summary(sol):
    took = utils.sec_to_min_if(sol.t0)
    return (f"computed_steps = {sol.computed_steps}, t0 = {sol.t0:g} s ({took}), "
            f"stop reason = {sol.stop_reason}")


def run(argv=None):
    cli = cli_parser.CLIParser('chsimpy_b200')
    cli.print_info()
    params = cli.get_parameters(argv)
    sim = simulator.Simulator(params)
    for i, chunk in enumerate(str(params).split(", '")):      # one parameter per line
        print(chunk if i == 0 else " '" + chunk)
    sol = sim.solve()
    sim.render()                                       # headless view stand-in: no-op unless a view is attached
    sim.export()
    print(_stop_summary(sol))
    if sim.export_requested():
        print("File ID = " + str(sim.solution_file_id))
    return sol


def main():
    run(sys.argv[1:])


if __name__ == '__main__':
    main()
