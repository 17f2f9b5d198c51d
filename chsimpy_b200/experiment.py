#!/usr/bin/env python
"""A0/A1 parameter-sweep ensemble on GPUs -- the `chsimpy-experiment` entry point of the
reference (chsimpy/experiment.py) with its `multiprocessing.Pool` over runs replaced by
batched, lock-step simulations on the device(s):

  * the table of A0/A1 factors is generated exactly as the reference does
    (uniform / sobol / grid / file, `--independent`; experiment.py:148-190);
  * every member shares seed and initial field and differs in (A0, A1, kappa_tilde)
    (experiment.py:87-101);  members of one rank step together in one `BatchStepper`;
  * members stop individually at their energy drop (device-side flags), the grid is
    compacted at every host poll;
  * one process per GPU (torchrun): run ids are sharded contiguously over ranks, there
    is NO collective on the data path; the 12-tuples of experiment.py:114-126 are
    gathered on rank 0 which writes `<id>-results.csv` / `<id>-results-agg.csv`.

The host-side sympy scalars (kappa_tilde, c_A/c_B, spinodal roots; ~0.9 s per member) are
computed by a host process pool, overlapped with nothing on the device only in the sense
that they must precede it (kappa_tilde is an input of the step).
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

from . import utils
from .cli_parser import CLIParser
from .parameters import Parameters
from .solution import Solution
from .timedata import TimeData

RESULT_COLUMNS = ['A0', 'A1', 'ca', 'cb', 'sa', 'sb', 'tau0', 't0', 'tsep', 'id', 'fac_A0', 'fac_A1']


class ExperimentParams:
    def __init__(self):
        self.runs = 2
        self.jitter_Arellow = 0.995
        self.jitter_Arelhigh = 1.005
        self.processes = -1
        self.independent = False
        self.A_source = 'uniform'
        self.A_seed = None


class ExperimentCLIParser:
    """Adds the experiment flags of reference experiment.py:37-59 to the common CLI."""

    def __init__(self):
        self.cliparser = CLIParser('chsimpy_b200 (experiment.py)')
        g = self.cliparser.parser.add_argument_group('Experiment')
        g.add_argument('-R', '--runs', default=3, type=int, help='Number of Monte-Carlo runs')
        g.add_argument('-P', '--processes', default=-1, type=int,
                       help='Host processes for the sympy scalars and the exports (-1 = physical cores)')
        g.add_argument('--independent', action='store_true', help='A0 and A1 do not vary at the same time')
        g.add_argument('--A-source', default='uniform', help="'uniform' | 'sobol' | 'grid' | <file with A0,A1 rows>")
        g.add_argument('--A-seed', default=85972, type=int, help='RNG seed for the A0/A1 factors')
        g.add_argument('--no-export', action='store_true', help='Skip the per-run yaml/csv files (results csv only)')

    def get_parameters(self, argv=None):
        params = self.cliparser.get_parameters(argv)
        a = self.cliparser.args
        ep = ExperimentParams()
        ep.runs, ep.independent, ep.A_source = a.runs, a.independent, a.A_source
        ep.processes, ep.A_seed = a.processes, a.A_seed
        ep.no_export = a.no_export
        params.no_gui = True
        params.yaml = True
        if a.export_csv is None:
            params.export_csv = 'U, E, E2, SA'
            params.compress_csv = True
        if ep.runs < 1:
            self.cliparser.parser.error('ERROR: --runs must be at least 1.')
        if params.png_anim:
            self.cliparser.parser.error('ERROR: --png-anim is not allowed.')
        return ep, params


def factor_table(ep):
    """(rand_values [items, 2] or None, A_list or None, number of items) -- reference
    experiment.py:148-190 and :204-209."""
    A_list = rand_values = None
    lo, hi = ep.jitter_Arellow, ep.jitter_Arelhigh
    if ep.A_source in ('uniform', 'sobol'):
        if ep.A_source == 'sobol':
            from scipy.stats import qmc
            q = qmc.Sobol(d=2, seed=ep.A_seed)
            pts = q.random_base2(int(np.ceil(np.log2(ep.runs))))
            cols = np.transpose(qmc.scale(pts, lo, hi)[:ep.runs])
        else:
            rng = np.random.Generator(np.random.PCG64(ep.A_seed))
            cols = np.transpose(rng.uniform(lo, hi, size=(ep.runs, 2)))
        if ep.independent:
            rand_values = np.ones((2 * ep.runs, 2))
            rand_values[:ep.runs, 0] = cols[0]
            rand_values[ep.runs:, 1] = cols[1]
        else:
            rand_values = np.ones((ep.runs, 2))
            rand_values[:, 0] = cols[0]
            rand_values[:, 1] = cols[1]
    elif ep.A_source == 'grid':
        nx = int(np.floor(np.sqrt(ep.runs)))
        ep.runs = nx * nx
        xvec = np.linspace(lo, hi, nx)
        if ep.independent:
            rand_values = np.ones((2 * nx, 2))
            rand_values[:nx, 0] = xvec
            rand_values[nx:, 1] = xvec
        else:
            rand_values = np.array([[v, w] for v in xvec for w in xvec])
    else:
        A_list = utils.csv_import_matrix(ep.A_source)
    n_items = rand_values.shape[0] if A_list is None else A_list.shape[0]
    if ep.independent and ep.A_source in ('sobol', 'uniform'):
        n_items = min(2 * ep.runs, n_items)
    else:
        n_items = min(ep.runs, n_items)
    return rand_values, A_list, n_items


def shard(n_items, rank, world):
    """Contiguous, balanced range of run ids of `rank` (no communication needed)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def member_params(init_params, run_id, rand_values, A_list):
    """Parameters of one member, reference experiment.py:87-101."""
    p = init_params.deepcopy()
    p.seed = init_params.seed
    p.file_id = f"{init_params.file_id}-run{run_id}"
    if A_list is None:
        f0, f1 = float(rand_values[run_id, 0]), float(rand_values[run_id, 1])
        p.func_A0 = lambda temp, f0=f0: utils.A0(temp) * f0
        p.func_A1 = lambda temp, f1=f1: utils.A1(temp) * f1
    else:
        a0, a1 = float(A_list[run_id][0]), float(A_list[run_id][1])
        p.func_A0 = lambda temp, a0=a0: a0
        p.func_A1 = lambda temp, a1=a1: a1
        f0 = f1 = None
    return p, f0, f1


def _host_scalars(job):
    """Per-member sympy work (reference solution.py:39-46, experiment.py:110-112), in a pool."""
    R, T, B, a0, a1, at, kappa_given = job
    if kappa_given is None:
        kappa = float(utils.get_distance_common_tangent(R=R, T=T, B=B, A0=a0, A1=a1, at=at)) / (0.1602564 * 64) ** 2
    else:
        kappa = kappa_given
    gap = utils.get_miscibility_gap(R, T, B, a0, a1)
    sa, sb = utils.get_roots_of_EPP(R, T, a0, a1)
    return kappa, float(gap[0]), float(gap[1]), float(sa), float(sb)


def initial_field(params):
    """The field every member starts from (same seed for all, quirk Q11)."""
    if params.Uinit_file is not None:
        return utils.csv_import_matrix(params.Uinit_file)
    N, c0 = params.N, params.XXX
    if params.generator == 'lcg':
        from . import mport
        return c0 + c0 * 0.01 * mport.matlab_lcg_sample(N, N, params.seed)
    if params.generator == 'sobol':
        from scipy.stats import qmc
        return c0 + c0 * 0.01 * (qmc.Sobol(d=N, seed=params.seed).random(N) - 0.5)
    if params.generator == 'uniform':
        return c0 + c0 * 0.01 * (np.random.Generator(np.random.PCG64(params.seed)).random((N, N)) - 0.5)
    raise ValueError("generator not supported in ensembles: " + str(params.generator))


def _noise_source(params, st):
    """draw(n) -> the next n (N, N) uniform draws of the members' generator, positioned after the U_init draw.
    PCG64: regenerated bit-exactly on the device (BatchStepper.pcg64_noise); Sobol: host draws."""
    N = params.N
    if params.generator == 'uniform':
        rng = np.random.Generator(np.random.PCG64(params.seed))
        rng.bit_generator.advance(N * N)                # U_init = the first (N, N) draw (solver.py:78-82)

        def draw(n):
            if hasattr(st, "pcg64_noise"):
                dev = st.pcg64_noise(rng.bit_generator.state, n)
                rng.bit_generator.advance(n * N * N)
                return dev
            return rng.random((n, N, N))
        return draw
    from scipy.stats import qmc
    sob = qmc.Sobol(d=N, seed=params.seed)
    sob.fast_forward(N)                                 # U_init = the first N points (solver.py:70-71)
    return lambda n: np.stack([sob.random(N) for _ in range(n)])


def solve_ensemble(init_params, rand_values=None, A_list=None, run_ids=None, U_init=None, host_procs=None,
                   batch_max=2048, poll_every=128, keep_fields=True, backend=None, device=None, timings=None,
                   pipeline_batch=256, scalars=None):
    """Runs the members `run_ids` in lock-step on one GPU.  Returns a list of dicts with the
    reference's 12-tuple (`tuple`), the TimeData (`timedata`), the final field (`U`, optional)
    and the per-member Solution scalars."""
    from .solver import BatchStepper, make_params_struct
    if run_ids is None:
        run_ids = range(rand_values.shape[0] if A_list is None else A_list.shape[0])
    run_ids = list(run_ids)
    t0 = time.perf_counter()
    members = [member_params(init_params, rid, rand_values, A_list) for rid in run_ids]
    jobs = []
    for p, _, _ in members:
        jobs.append((p.R, p.temp, p.B, float(p.func_A0(p.temp)), float(p.func_A1(p.temp)), p.XXX, p.kappa_tilde))
    nproc = host_procs or min(len(jobs), utils.get_number_physical_cores() or 1)
    pool = None
    if scalars is not None:                            # computed by the caller (work-queue driver: one shared pool)
        assert len(scalars) == len(jobs)
        scal_iter = iter(scalars)
    elif nproc > 1 and len(jobs) > 1:
        # forkserver: the workers start from a clean process (no CUDA context, no torch threads).
        # imap streams the results in order: the device runs batch k while the pool is still
        # working on the scalars of batch k+1 (the sympy work is the larger part of an ensemble)
        pool = mp.get_context("forkserver").Pool(nproc)
        scal_iter = pool.imap(_host_scalars, jobs, chunksize=max(1, min(8, len(jobs) // (4 * nproc))))
        batch_max = min(batch_max, max(int(pipeline_batch), 1))
    else:
        scal_iter = iter([_host_scalars(j) for j in jobs])
    scal = []
    t_host = 0.0
    generated = U_init is None and init_params.Uinit_file is None
    if U_init is None:
        if init_params.generator == 'lcg' and init_params.Uinit_file is None:
            from .solver import _CudaBackend, lcg_sample      # the serial float64 LCG as a device kernel
            be_ = backend if backend is not None else _CudaBackend(device)
            U_init = init_params.XXX + init_params.XXX * 0.01 * lcg_sample(be_, init_params.N, init_params.N, init_params.seed)
        else:
            U_init = initial_field(init_params)
    assert U_init.shape == (init_params.N, init_params.N)
    # jitter (reference solver.py:210-211): every member owns a generator with the SAME seed (quirk Q11), whose
    # stream continues after the U_init draw -- so one noise stream serves all members of a batch
    jitter_on = init_params.jitter is not None and 0.0 < init_params.jitter < 0.1
    if jitter_on and (not generated or init_params.generator not in ('uniform', 'sobol')):
        raise TypeError("'NoneType' object is not callable")      # create_rand is None (quirk Q7)
    out = []
    t_dev = 0.0
    be_probe = backend if backend is not None else None
    lib_ = be_probe.lib if be_probe is not None else None
    if lib_ is None:
        from . import _lib as _libmod
        lib_ = _libmod.load()
    if not lib_.chs_supports_n(int(init_params.N)):
        # Sizes the batched kernels do not cover (the reference accepts any N, cli_parser.py:27): the members run one
        # after the other through the single-simulation engines of Solver (slab path / tensor-core GEMM path) -- the
        # same device code a single `chsimpy -N <n>` run uses; only the lock-step batching is missing.
        from .solver import Solver
        for i, (p, f0, f1) in enumerate(members):
            tw = time.perf_counter()
            scal.append(next(scal_iter))
            t_host += time.perf_counter() - tw
            p.kappa_tilde = scal[i][0]
            td = time.perf_counter()
            slv = Solver(p, None if generated else U_init, _backend=backend)
            slv.prepare()
            sol = slv.solve_or_resume(init_params.ntmax)       # AssertionError on a NaN row, as the reference would
            if not keep_fields:
                sol.U = None
            t_dev += time.perf_counter() - td
            kappa, ca, cb, sa, sb = scal[i]
            tup = (sol.A0, sol.A1, ca, cb, sa, sb, sol.tau0, sol.t0, int(np.argmax(sol.E2)), run_ids[i], f0, f1)
            out.append({"tuple": tup, "solution": sol, "params": p, "run_id": run_ids[i]})
            del slv
        members = []
    for c0 in range(0, len(members), batch_max):
        chunk = list(range(c0, min(c0 + batch_max, len(members))))
        tw = time.perf_counter()
        scal.extend(next(scal_iter) for _ in chunk)    # waits for this batch's scalars only
        t_host += time.perf_counter() - tw
        sols, structs = [], []
        for i in chunk:
            p = members[i][0]
            p.kappa_tilde = scal[i][0]                 # computed in the pool; Solution() will not redo it
            s = Solution(p)
            sols.append(s)
            structs.append(make_params_struct(p, s))
        td = time.perf_counter()
        st = BatchStepper(init_params.N, structs, rows_cap=max(poll_every, 16), backend=backend, device=device)
        st.set_U(U_init)
        row0 = st.prepare()
        iters = max(init_params.ntmax, 0) - 1          # first solve_or_resume call: ntmax-1 iterations (Q3)
        rows, _ = st.run(iters, draw_noise=_noise_source(init_params, st) if jitter_on else None,
                         poll_every=min(poll_every, 64) if jitter_on else poll_every)
        fields = st.get_U() if keep_fields else None
        states = [st.get_state(j) for j in range(len(chunk))]
        t_dev += time.perf_counter() - td
        for j, i in enumerate(chunk):
            p, f0, f1 = members[i]
            data = TimeData(capacity=len(rows[j]) + 1)
            data.extend(row0[j][None, :])
            sol = sols[j]
            sol.timedata = data
            s_ = states[j]
            sol.computed_steps = int(s_.computed_steps)
            sol.tau0 = int(s_.tau0) if s_.tau0 != 0 else 0.0
            sol.t0 = float(s_.t0)
            sol.stop_reason = {0: 'None', 1: 'energy', 2: 'time-limit'}.get(s_.stop_reason, 'None')
            if fields is not None:
                sol.U = fields[j]
            data.extend(rows[j])                       # AssertionError on NaN, as the reference would
            kappa, ca, cb, sa, sb = scal[i]
            tup = (sol.A0, sol.A1, ca, cb, sa, sb, sol.tau0, sol.t0, int(np.argmax(sol.E2)), run_ids[i], f0, f1)
            out.append({"tuple": tup, "solution": sol, "params": p, "run_id": run_ids[i]})
        del st
    if pool is not None:
        pool.close()
        pool.join()
    if timings is not None:
        timings.update(host_scalars_s=t_host, device_s=t_dev, host_procs=nproc, total_s=time.perf_counter() - t0)
    return out


def scalar_jobs(init_params, rand_values, A_list, run_ids):
    """The _host_scalars job of every member in run_ids (what solve_ensemble would build itself)."""
    jobs = []
    for rid in run_ids:
        p, _, _ = member_params(init_params, rid, rand_values, A_list)
        jobs.append((p.R, p.temp, p.B, float(p.func_A0(p.temp)), float(p.func_A1(p.temp)), p.XXX, p.kappa_tilde))
    return jobs


def solve_from_queue(init_params, rand_values, A_list, claim, host_procs, keep_fields=True, backend=None, device=None,
                     timings=None):
    """Dynamic load balance over ranks: `claim()` hands out the next chunk of run ids (a shared
    counter, see work_queue) or None.  The stop step varies by +-15 % with (A0, A1), so a static split leaves
    ranks idle; here a rank that finishes early simply claims more.  The sympy scalars of chunk k+1 are computed
    by the host pool while the device steps chunk k."""
    pool = mp.get_context("forkserver").Pool(max(1, host_procs)) if host_procs and host_procs > 1 else None

    def submit(ids):
        jobs = scalar_jobs(init_params, rand_values, A_list, ids)
        if pool is None:
            return [_host_scalars(j) for j in jobs]
        return pool.map_async(_host_scalars, jobs, chunksize=max(1, len(jobs) // (4 * host_procs)))
    out, chunks = [], 0
    t0 = time.perf_counter()
    t_dev = t_wait = 0.0
    nxt = claim()
    nxt_s = submit(nxt) if nxt is not None else None
    U_init = None
    while nxt is not None:
        cur, cur_s = nxt, nxt_s
        nxt = claim()
        nxt_s = submit(nxt) if nxt is not None else None
        tw = time.perf_counter()
        scal = cur_s if isinstance(cur_s, list) else cur_s.get()
        t_wait += time.perf_counter() - tw
        td = time.perf_counter()
        res = solve_ensemble(init_params, rand_values, A_list, run_ids=cur, U_init=U_init, host_procs=1,
                             keep_fields=keep_fields, backend=backend, device=device, scalars=scal)
        t_dev += time.perf_counter() - td
        out.extend(res)
        chunks += 1
    if pool is not None:
        pool.close()
        pool.join()
    if timings is not None:
        timings.update(host_scalars_s=t_wait, device_s=t_dev, host_procs=host_procs, chunks=chunks,
                       total_s=time.perf_counter() - t0)
    return out


def work_queue(n_items, chunk, store=None, key="chs_ensemble_next_chunk"):
    """claim() for solve_from_queue: an atomic counter in the process group's store (world > 1) or a local one."""
    state = {"next": 0}

    def claim():
        c = (store.add(key, 1) - 1) if store is not None else state["next"]
        state["next"] += 1
        lo = c * chunk
        return list(range(lo, min(lo + chunk, n_items))) if lo < n_items else None
    return claim


def _export_member(job):
    """Per-run files of reference simulator.py:135-156 (yaml scalars + csv matrices)."""
    fname_sol, yaml_on, export_csv, compress, sol = job
    if yaml_on:
        sol.yaml_export_scalars(fname=fname_sol + '.yaml')
    if export_csv is not None:
        fext = 'csv.bz2' if compress else 'csv'
        for member in export_csv.replace(' ', '').split(','):
            arr = getattr(sol, member, None)
            if isinstance(arr, np.ndarray):
                utils.csv_export_matrix(arr, fname=f"{fname_sol}.{member}.{fext}")
    return fname_sol


def main(argv=None):
    import pandas as pd
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    parser = ExperimentCLIParser()
    if rank == 0:
        parser.cliparser.print_info()
    ep, init_params = parser.get_parameters(argv)
    if init_params.file_id is None or init_params.file_id == 'auto':
        init_params.file_id = utils.get_or_create_file_id(init_params.file_id)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")               # control plane only: gathers result tuples
        fid = [init_params.file_id]
        dist.broadcast_object_list(fid, src=0)
        init_params.file_id = fid[0]
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    rand_values, A_list, n_items = factor_table(ep)
    if rank == 0:
        print(str(init_params).replace(", '", "\n '"))
        utils_meta = [f"runs, {ep.runs}", f"A_source, {ep.A_source}", f"A_seed, {ep.A_seed}",
                      f"independent, {ep.independent}", f"world_size, {world}",
                      f"localtime, {utils.get_current_localtime()}", f"argv, '{' '.join(sys.argv)}'"]
        with open(f"{init_params.file_id}-metadata.csv", 'w') as f:
            f.write("\n".join(utils_meta))
    tm = {}
    t0 = time.perf_counter()
    # the ranks of a box share its host cores: each gets its share for the sympy pool
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    share = max(1, (utils.get_number_physical_cores() or 1) // max(1, local_world))
    if world > 1:
        # dynamic load balance: chunks of members are claimed from a counter in the process group's store
        import torch.distributed as dist
        store = dist.distributed_c10d._get_default_store()
        chunk = max(16, min(256, -(-n_items // (4 * world))))
        res = solve_from_queue(init_params, rand_values, A_list, work_queue(n_items, chunk, store),
                               host_procs=share if ep.processes == -1 else max(1, ep.processes),
                               keep_fields=not ep.no_export, timings=tm)
    else:
        res = solve_ensemble(init_params, rand_values, A_list, run_ids=shard(n_items, rank, world),
                             host_procs=None if ep.processes == -1 else max(1, ep.processes), timings=tm,
                             keep_fields=not ep.no_export)
    t_solve = time.perf_counter() - t0
    if not ep.no_export:
        jobs = [(f"{r['params'].file_id}.solution", init_params.yaml, init_params.export_csv,
                 init_params.compress_csv, r["solution"]) for r in res]
        # threads, not processes: Solution objects hold lambdas (not picklable) and a CUDA process
        # should not fork; bz2 compression releases the GIL
        from concurrent.futures import ThreadPoolExecutor
        nthr = max(1, min(len(jobs), utils.get_number_physical_cores() or 1))
        with ThreadPoolExecutor(nthr) as ex_:
            list(ex_.map(_export_member, jobs))
    tuples = [r["tuple"] for r in res]
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(tuples, gathered, dst=0)
        if rank == 0:
            tuples = sorted((t for part in gathered for t in part), key=lambda t: t[9])     # by run id
        dist.barrier()
    if rank == 0:
        df = pd.DataFrame(tuples, columns=RESULT_COLUMNS)
        df[['tau0', 'id']] = df[['tau0', 'id']].astype(int)
        df.to_csv(f"{init_params.file_id}-results.csv")
        agg = df.loc[:, df.columns != 'id'].describe()
        agg.loc['cv'] = agg.loc['std'] / agg.loc['mean']
        print(agg.T)
        agg.T.to_csv(f"{init_params.file_id}-results-agg.csv")
        print(f"members: {len(tuples)} on {world} GPU(s); rank-0 solve {t_solve:.2f} s "
              f"(waiting for the host sympy scalars {tm.get('host_scalars_s', 0):.2f} s on {tm.get('host_procs')} procs, "
              f"device {tm.get('device_s', 0):.2f} s, overlapped)")
        print('Output files:')
        for suffix in ('metadata', 'results-agg', 'results'):
            print(f"  {init_params.file_id}-{suffix}.csv")
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
