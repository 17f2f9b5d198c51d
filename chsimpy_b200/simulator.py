"""Simulator: orchestration layer with the call surface of reference chsimpy/simulator.py
(`Simulator(params, U_init)`, `.solve()`, `.render()`, `.export()`), driving the GPU Solver.

The matplotlib dashboards (reference plotview.py / mapview.py) are host-side GUI code and
out of scope here; when a view would be required a headless stand-in keeps the chunked
`update_every` solve path (reference simulator.py:56-87) working without drawing."""
import warnings

import numpy as np

from . import parameters, solver, utils


class _HeadlessView:
    """Accepts the calls Simulator makes on PlotView/MapView and draws nothing."""

    def __getattr__(self, name):
        return lambda *a, **k: None


class Simulator:
    def __init__(self, params=None, U_init=None, _backend=None):
        self.params = parameters.Parameters() if params is None else params
        if U_init is None and params.Uinit_file is not None:        # params=None fails here as in the reference (Q10)
            U_init = utils.csv_import_matrix(params.Uinit_file)
        self.solver = solver.Solver(params, U_init, _backend=_backend)
        self.steps_total = 0
        self.solution_file_id = None
        if self.gui_required():
            warnings.warn("chsimpy_b200 has no matplotlib views; running headless (pass no_gui=True to silence)")
            self.view = _HeadlessView()
        else:
            self.view = None
            self.params.update_every = None                         # reference simulator.py:33-34 (Q9)

    def solve(self):
        self.solution_file_id = utils.get_or_create_file_id(self.params.file_id)
        if self.steps_total == 0:
            self.solver.prepare()
        if self.params.update_every is None:
            return self.solver.solve_or_resume(self.params.ntmax)
        # ---- chunked solve, reference simulator.py:56-87
        p, slv = self.params, self.solver
        self.view.prepare(show=self.gui_requested())
        part = 0
        steps_end = p.ntmax
        if p.time_max is not None and p.time_max > 0:
            steps_end = utils.get_int_max_value()
        dsteps = min(steps_end, p.update_every)
        assert (dsteps > 0)
        while ((self.steps_total + dsteps) <= steps_end
               and (slv.solution.stop_reason == 'None' or p.full_sim is True)
               and (slv.solution.stop_reason != 'time-limit')):
            slv.solve_or_resume(dsteps)
            self.view.draw()
            self.steps_total += dsteps
            part += 1
            diff = steps_end - self.steps_total
            if 0 < diff < dsteps:
                dsteps = diff
            elif diff < 0:
                raise Exception("Something went wrong.")
        self.view.finish()
        if slv.solution.tau0 == 0:
            slv.solution.tau0 = slv.solution.computed_steps - 1
            slv.solution.t0 = slv.time_passed
        return slv.solution

    def export(self):
        """YAML scalars and CSV matrices of the solution (reference simulator.py:135-156)."""
        fname_sol = f"{self.solution_file_id}.solution"
        solution = self.solver.solution
        if self.params.yaml:
            solution.yaml_export_scalars(fname=fname_sol + '.yaml')
        if self.params.export_csv is not None:
            fext = 'csv.bz2' if self.params.compress_csv else 'csv'
            for member in self.params.export_csv.replace(' ', '').split(','):
                arr = getattr(solution, member, None)
                if isinstance(arr, np.ndarray):
                    utils.csv_export_matrix(arr, fname=f"{fname_sol}.{member}.{fext}")
        return fname_sol

    def render(self):
        return None          # no views in this package

    def export_requested(self):
        p = self.params
        return p.export_csv is not None or p.yaml or p.png or p.png_anim

    def gui_requested(self):
        return self.params.no_gui is False

    def gui_required(self):
        return self.params.png or self.params.png_anim or self.gui_requested()
