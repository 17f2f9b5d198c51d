"""Small slab-path run for ncu: N (default 8192), a few steps on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chsimpy_b200 as ch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = ch.Parameters(); p.no_gui = True; p.N = N; p.full_sim = True; p.kappa_tilde = 2.989112919661156e-4
s = ch.Solver(p); s.prepare()
s._stepper.run(steps)
torch.cuda.synchronize()
print("ok", N, steps)
