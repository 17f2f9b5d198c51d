"""Ensemble driver: factor tables as the reference builds them (experiment.py:148-190),
sharding of run ids over ranks, and the N>1 path on world_size-2 gloo (CPU) with the
emulated kernels standing in for the device."""
import os
import sys

import numpy as np
import pytest

import chsimpy_b200 as ch
from chsimpy_b200 import experiment as ex

HERE = os.path.dirname(os.path.abspath(__file__))


def test_factor_table_uniform_matches_reference_recipe():
    ep = ex.ExperimentParams()
    ep.runs, ep.A_seed = 1024, 85972
    rv, A_list, n = ex.factor_table(ep)
    want = np.random.Generator(np.random.PCG64(85972)).uniform(0.995, 1.005, size=(1024, 2))
    assert A_list is None and n == 1024 and np.array_equal(rv, want)
    ep.independent = True
    rv, _, n = ex.factor_table(ep)
    assert n == 2048 and np.all(rv[:1024, 1] == 1) and np.all(rv[1024:, 0] == 1)
    assert np.array_equal(rv[:1024, 0], want[:, 0]) and np.array_equal(rv[1024:, 1], want[:, 1])


def test_factor_table_grid_and_sobol():
    ep = ex.ExperimentParams()
    ep.runs, ep.A_source = 10, 'grid'
    rv, _, n = ex.factor_table(ep)
    assert ep.runs == 9 and n == 9 and rv.shape == (9, 2)
    assert np.allclose(rv[0], [0.995, 0.995]) and np.allclose(rv[-1], [1.005, 1.005]) and np.allclose(rv[1], [0.995, 1.0])
    ep = ex.ExperimentParams()
    ep.runs, ep.A_source, ep.A_seed = 5, 'sobol', 85972
    rv, _, n = ex.factor_table(ep)
    assert n == 5 and rv.shape == (5, 2) and (rv >= 0.995).all() and (rv <= 1.005).all()


def test_factor_tables_equal_the_reference_output():
    """Every A-source against the tables the unmodified reference built (tests/golden/make_factor_tables.py)."""
    z = np.load(os.path.join(HERE, "golden", "factor_tables.npz"))
    cases = {"uniform_r7": (7, "uniform", 85972, False), "uniform_r7_indep": (7, "uniform", 85972, True),
             "sobol_r5": (5, "sobol", 85972, False), "sobol_r6_indep": (6, "sobol", 11, True),
             "grid_r10": (10, "grid", None, False), "grid_r17_indep": (17, "grid", None, True)}
    for name, (runs, src, seed, indep) in cases.items():
        ep = ex.ExperimentParams()
        ep.runs, ep.A_source, ep.A_seed, ep.independent = runs, src, seed, indep
        rv, A_list, n = ex.factor_table(ep)
        assert A_list is None and n == int(z[name + "_n"]), name
        assert rv.shape == z[name].shape and np.array_equal(rv, z[name]), name      # bit-identical


def test_ensemble_jitter_uses_the_members_noise_stream():
    """--jitter in an ensemble: every member continues the PCG64(seed) stream after the U_init draw
    (reference solver.py:78-82,210-211; same seed for all members, Q11) -- so a one-member ensemble with
    factors (1, 1) must equal the plain Solver run, and the noise must actually be applied."""
    sys.path.insert(0, HERE)
    from emu_lib import EmuBackend
    be = EmuBackend()
    p = _params()
    p.jitter, p.ntmax = 0.004, 9
    rv = np.ones((2, 2))
    rv[1] = [1.003, 0.998]
    res = ex.solve_ensemble(p, rv, None, host_procs=1, backend=be)
    s = ch.Solver(p.deepcopy(), _backend=be)
    s.prepare()
    sol = s.solve_or_resume(p.ntmax)
    got = res[0]["solution"]
    assert got.computed_steps == sol.computed_steps == 9
    assert np.array_equal(got.timedata.data(), sol.timedata.data())
    assert np.array_equal(got.U, sol.U) and not np.array_equal(got.U, s.U_init)
    assert not np.array_equal(res[1]["solution"].U, got.U)
    p2 = _params()
    p2.jitter, p2.generator = 0.004, "lcg"
    with pytest.raises(TypeError):                      # create_rand is None for -g lcg (quirk Q7)
        ex.solve_ensemble(p2, rv, None, host_procs=1, backend=be)


def test_shard_covers_everything_once():
    for n in (1, 7, 8, 1024, 1025):
        for w in (1, 2, 3, 8):
            parts = [list(ex.shard(n, r, w)) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _params():
    p = ch.Parameters()
    p.N, p.no_gui, p.kappa_tilde, p.ntmax, p.full_sim = 32, True, 3e-4, 12, True
    p.file_id = "t"
    return p


def _rank_main(rank, world, port, q):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from emu_lib import EmuBackend
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ep = ex.ExperimentParams()
    ep.runs, ep.A_seed = 5, 85972
    rv, A_list, n = ex.factor_table(ep)
    mine = ex.shard(n, rank, world)
    res = ex.solve_ensemble(_params(), rv, A_list, run_ids=mine, host_procs=1, backend=EmuBackend())
    tuples = [r["tuple"] for r in res]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(tuples, gathered, dst=0)
    if rank == 0:
        q.put([t for part in gathered for t in part])
    dist.barrier()
    dist.destroy_process_group()


def _rank_queue(rank, world, port, q):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from emu_lib import EmuBackend
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ep = ex.ExperimentParams()
    ep.runs, ep.A_seed = 7, 85972
    rv, A_list, n = ex.factor_table(ep)
    store = dist.distributed_c10d._get_default_store()
    res = ex.solve_from_queue(_params(), rv, A_list, ex.work_queue(n, 2, store), host_procs=1, backend=EmuBackend())
    tuples = [r["tuple"] for r in res]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(tuples, gathered, dst=0)
    if rank == 0:
        q.put([[t[9] for t in part] for part in gathered] + [sorted((t for part in gathered for t in part), key=lambda t: t[9])])
    dist.barrier()
    dist.destroy_process_group()


def test_work_queue_two_ranks_covers_every_member_once():
    """Dynamic load balance: chunks of run ids claimed from a shared counter; every member is solved
    exactly once, whichever rank claims it, with the same bits as the single-process run."""
    import torch.multiprocessing as tmp
    from emu_lib import EmuBackend
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 27500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_queue, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ids0, ids1, tuples = got
    assert sorted(ids0 + ids1) == list(range(7)) and ids0 and ids1          # both ranks got work, nothing twice
    ep = ex.ExperimentParams()
    ep.runs, ep.A_seed = 7, 85972
    rv, A_list, n = ex.factor_table(ep)
    ref = [r["tuple"] for r in ex.solve_ensemble(_params(), rv, A_list, host_procs=1, backend=EmuBackend())]
    assert tuples == ref


def test_two_rank_gloo_equals_single_process():
    import torch.multiprocessing as tmp
    from emu_lib import EmuBackend
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ep = ex.ExperimentParams()
    ep.runs, ep.A_seed = 5, 85972
    rv, A_list, n = ex.factor_table(ep)
    ref = [r["tuple"] for r in ex.solve_ensemble(_params(), rv, A_list, host_procs=1, backend=EmuBackend())]
    assert len(got) == 5 and [t[9] for t in got] == [0, 1, 2, 3, 4]
    for a, b in zip(got, ref):
        assert a == b                      # deterministic kernels: bit-identical regardless of the sharding
    # members differ only through A0/A1 (kappa is pinned here): tsep/tau0 columns are well-formed
    assert all(isinstance(t[8], int) for t in got)


def test_ensemble_at_a_size_without_batched_kernels():
    """The reference runs its ensemble at any N (cli_parser.py:27); N = 120 has no lock-step kernels here (neither a
    power of two nor <= 104), so the members go one after the other through Solver's single-simulation engine
    (BigEngine): every member must equal a stand-alone Solver run with that member's parameters."""
    from emu_lib import EmuBackend
    be = EmuBackend()
    assert be.lib.chs_supports_n(120) == 0
    p0 = _params()
    p0.N, p0.ntmax = 120, 5
    ep = ex.ExperimentParams()
    ep.runs, ep.A_seed = 2, 85972
    rv, A_list, n = ex.factor_table(ep)
    res = ex.solve_ensemble(p0, rv, A_list, host_procs=1, backend=be)
    assert [r["run_id"] for r in res] == [0, 1]
    for r in res:
        pm, f0, f1 = ex.member_params(p0, r["run_id"], rv, A_list)
        pm.kappa_tilde = r["params"].kappa_tilde
        s = ch.Solver(pm, _backend=be)
        s.prepare()
        ref = s.solve_or_resume(p0.ntmax)
        assert r["solution"].computed_steps == ref.computed_steps == 5
        assert np.array_equal(r["solution"].timedata.data(), ref.timedata.data())
        assert np.array_equal(r["solution"].U, ref.U)
        assert r["tuple"][0] == ref.A0 and r["tuple"][1] == ref.A1 and r["tuple"][9] == r["run_id"]
    assert res[0]["tuple"][0] != res[1]["tuple"][0]           # the members really differ
