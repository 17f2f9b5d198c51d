"""Small end-to-end runs of every kernel family, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import chsimpy_b200 as ch

def run(**kw):
    p = ch.Parameters(); p.no_gui = True; p.kappa_tilde = 3e-4
    force_slab = kw.pop("_force_slab", False)
    for k, v in kw.items(): setattr(p, k, v)
    s = ch.Solver(p, _force_slab=force_slab); s.prepare(); sol = s.solve_or_resume(p.ntmax)
    torch.cuda.synchronize()
    print(kw, "->", sol.computed_steps, float(sol.E[-1]), flush=True)

run(N=64, ntmax=12, full_sim=True)                                   # batched FFT path
run(N=64, ntmax=12, full_sim=True, jitter=0.005)                     # jitter: device PCG64, k_diag
run(N=64, ntmax=520, full_sim=True, adaptive_time=True, delt_max=4e-10)   # adaptive dt column sums (> 500 steps)
run(N=100, ntmax=6, full_sim=True)                                   # GEMM path (DMMA)
run(N=64, ntmax=6, full_sim=True, _force_slab=True)                  # slab path, point-major tiles
run(N=2048, ntmax=3, full_sim=True)                                  # slab path, line-major tiles
print("done")
