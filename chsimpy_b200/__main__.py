"""`python -m chsimpy_b200` -- same flow as reference chsimpy/__main__.py:8-25."""
from . import utils
from .cli_parser import CLIParser
from .simulator import Simulator


def main():
    parser = CLIParser('chsimpy_b200')
    parser.print_info()
    params = parser.get_parameters()
    simulator = Simulator(params)
    print(str(params).replace(", '", "\n '"))
    solution = simulator.solve()
    simulator.render()
    simulator.export()
    print(f"computed_steps = {solution.computed_steps}, t0 = {solution.t0:g} s "
          f"({utils.sec_to_min_if(solution.t0)}), stop reason = {solution.stop_reason}")
    if simulator.export_requested():
        print(f"File ID = {simulator.solution_file_id}")


if __name__ == '__main__':
    main()
