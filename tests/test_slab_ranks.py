"""Row-slab decomposition over ranks (chsimpy_b200/slab.py) on world_size-2 gloo, with the
host-compiled kernels standing in for the device: transposes through all-to-all, global row
indices, y-edge ownership, the all-reduced diagnostic sums and the replicated control kernel.
The GPU run of the same path (NCCL / peer-memory transposes) is tools/slab_check.py."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _rank_main(rank, world, port, q):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import chsimpy_b200 as ch
    from chsimpy_b200.slab import SlabEngine
    from emu_lib import EmuBackend
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(GOLD, "n64_k200.npz"))
    m = json.loads(str(z["meta"]))
    p = ch.Parameters()
    p.no_gui = True
    for k, v in m["params"].items():
        setattr(p, k, v)
    s = ch.Solver(p, _backend=EmuBackend(), _world=(rank, world))
    assert isinstance(s._stepper, SlabEngine) and s._stepper.R == 64 // world
    s.prepare()
    s.solve_or_resume(7)
    sol = s.solve_or_resume(5)                       # re-entry: hat_U recomputed from the gathered field
    if rank == 0:
        q.put((sol.timedata.data(), sol.U, sol.computed_steps, sol.tau0))
    dist.barrier()
    dist.destroy_process_group()


def _rank_jitter(rank, world, port, q):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import chsimpy_b200 as ch
    from emu_lib import EmuBackend
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = ch.Parameters()
    p.N, p.no_gui, p.full_sim, p.kappa_tilde, p.seed, p.ntmax, p.jitter = 64, True, True, 2.7e-4, 5, 9, 0.004
    s = ch.Solver(p, _backend=EmuBackend(), _world=(rank, world))
    s.prepare()
    s.solve_or_resume(6)
    sol = s.solve_or_resume(3)                       # re-entry from the jittered field (quirk Q2)
    if rank == 0:
        q.put((sol.timedata.data(), sol.U, sol.computed_steps))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_two_ranks_jitter_matches_oracle():
    """--jitter on two row slabs: each rank generates ITS rows of the PCG64 noise stream, the mean of the whole
    draw is all-reduced, the stencil gradient energy takes a 1-row halo from the neighbour (all-gather)."""
    import torch.multiprocessing as tmp
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    import ch_oracle as orc
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_jitter, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows, U, steps = q.get(timeout=900)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    o = orc.run_default(N=64, nsteps=6, seed=5, kappa_tilde=2.7e-4, full_sim=True, jitter=0.004)
    o.run(3)
    assert steps == o.computed_steps == 9 and rows.shape == o.rows.shape
    rel = np.abs(rows - o.rows) / np.maximum(np.abs(o.rows), 1e-300)
    rel[o.rows == 0] = np.abs(rows[o.rows == 0])
    assert rel.max() < 1e-9, rel.max(axis=0)
    assert np.abs(U - o.U).max() < 1e-11


def test_slab_two_ranks_gloo_matches_reference_fixture():
    import torch.multiprocessing as tmp
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows, U, steps, tau0 = q.get(timeout=900)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    z = np.load(os.path.join(GOLD, "n64_k200.npz"))
    ref = z["rows"][:12]
    assert steps == 12 and rows.shape == ref.shape
    rel = np.abs(rows - ref) / np.maximum(np.abs(ref), 1e-300)
    rel[ref == 0] = np.abs(rows[ref == 0])
    assert rel.max() < 1e-11, rel.max(axis=0)
    assert U.shape == (64, 64) and abs(U.mean() - 0.875) < 1e-3
