"""Config 4 (jitter + adaptive dt) single run, for ncu launch lists."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chsimpy_b200 as ch
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 700
p = ch.Parameters()
p.no_gui, p.full_sim, p.jitter, p.adaptive_time, p.delt_max, p.ntmax = True, True, 0.01, True, 2e-10, steps
p.kappa_tilde = 2.989112919661156e-4
s = ch.Solver(p); s.prepare(); torch.cuda.synchronize()
t = time.perf_counter(); sol = s.solve_or_resume(p.ntmax); torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"{steps-1} steps in {dt*1e3:.1f} ms = {(steps-1)/dt:.0f} steps/s, {dt/(steps-1)*1e6:.1f} us/step")
