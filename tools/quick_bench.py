"""Quick device-side throughput probe (not the contract bench): sim-steps/s vs batch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import chsimpy_b200 as ch
from chsimpy_b200.solver import BatchStepper, make_params_struct

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
batches = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 8, 64, 256, 1024]
p = ch.Parameters(); p.N = N; p.no_gui = True; p.full_sim = True; p.kappa_tilde = 2.989112919661156e-4
s0 = ch.Solver(p)
ps = make_params_struct(p, s0.solution)
for B in batches:
    st = BatchStepper(N, [ps] * B, rows_cap=512)
    st.set_U(s0.U_init); st.prepare(); st.begin()
    K = 100 if B >= 64 else 400
    st.steps(20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st.steps(K); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rate = B * K / (ms * 1e-3)
    print(f"N={N} B={B:5d} K={K} {ms/K*1e3:9.1f} us/step  {rate:12.0f} sim-steps/s  "
          f"roofline(32N^2) {rate*32*N*N/1e9:8.1f} GB/s = {rate*32*N*N/6554.9e9:.3f} of measured HBM peak")
    st.poll(); st.end()
    del st
