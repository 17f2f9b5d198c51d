"""Multi-rank check + timing of the slab path (run under torchrun, one rank per GPU):
   parity of N=2048 (10 steps) against the golden fixture, then steps/s at a large N."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import chsimpy_b200 as ch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W = (rank, world) if world > 1 else None
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

for case in ("n2048_k10", "n8192_k4") + (("n2048_jitter_adaptive",) if "--jitter" in sys.argv else ()):       # frozen outputs of the reference (tests/golden/make_golden.py)
    z = np.load(os.path.join(gold, case + ".npz")); m = json.loads(str(z["meta"]))
    p = ch.Parameters(); p.no_gui = True
    for k, v in m["params"].items(): setattr(p, k, v)
    s = ch.Solver(p, _world=W); s.prepare(); sol = s.solve_or_resume(p.ntmax)
    rows, ref = sol.timedata.data(), z["rows"]
    rel = np.abs(rows - ref) / np.maximum(np.abs(ref), 1e-300); rel[ref == 0] = np.abs(rows[ref == 0])
    st = p.N // 64
    du = np.abs(sol.U[::st, ::st] - z["U_sample"]).max()
    if rank == 0:
        print(f"[slab parity] world={world} N={p.N} rows {rows.shape} max rel row err {rel.max():.2e} max|dU| {du:.2e}", flush=True)
    assert rel.max() < 1e-9 and du < 1e-11
    del s

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(argv[0]) if len(argv) > 0 else 8192
steps = int(argv[1]) if len(argv) > 1 else 20
p = ch.Parameters(); p.no_gui = True; p.N = N; p.full_sim = True; p.kappa_tilde = 2.989112919661156e-4
t = time.perf_counter(); s = ch.Solver(p, _world=W); s.prepare(); torch.cuda.synchronize()
if rank == 0: print(f"[slab] N={N} setup+prepare {time.perf_counter()-t:.1f}s", flush=True)
eng = s._stepper
eng.run(3)                                 # warm-up
torch.cuda.synchronize()
if world > 1: dist.barrier()
t = time.perf_counter()
rows, done = eng.run(steps)                # begin + steps iterations + polls; the field stays on the device
torch.cuda.synchronize()
dt = time.perf_counter() - t
sol = s.solution
sol.timedata.extend(rows[0])
sol.computed_steps = int(eng.get_state(0).computed_steps)
if world > 1:
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX); dt = float(tt.item())
if rank == 0:
    by = 32.0 * N * N
    print(json.dumps({"slab": True, "N": N, "n_gpus": world, "steps": steps, "ms_per_step": round(dt / steps * 1e3, 3),
                      "steps_per_s": round(steps / dt, 2), "algorithmic_GBps_32N2": round(by * steps / dt / 1e9, 1),
                      "E_last": float(sol.E[-1]), "computed_steps": sol.computed_steps}), flush=True)
if world > 1: dist.destroy_process_group()
