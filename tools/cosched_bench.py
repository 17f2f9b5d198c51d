"""Experiment: two half-batches on two streams, column kernel of one co-resident with the row
kernel of the other (persistent build, CHS_CTAS_PER_SM caps each kernel's share of an SM)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import chsimpy_b200 as ch
from chsimpy_b200.solver import BatchStepper, make_params_struct

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
K = int(sys.argv[2]) if len(sys.argv) > 2 else 60
offset = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
N = 512
p = ch.Parameters(); p.N = N; p.no_gui = True; p.full_sim = True; p.kappa_tilde = 2.989112919661156e-4
U0 = 0.875 + 0.875 * 0.01 * (np.random.Generator(np.random.PCG64(2023)).random((N, N)) - 0.5)
ps = make_params_struct(p, ch.Solution(p))
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
sts = []
for s in streams:
    with torch.cuda.stream(s):
        st = BatchStepper(N, [ps] * (B // 2), rows_cap=K + 40)
        st.set_U(U0); st.prepare(); st.begin(); st.steps(10)
        sts.append(st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(torch.cuda.default_stream())
for s in streams: s.wait_event(e0)
if offset:                      # delay stream 1 by roughly one kernel
    with torch.cuda.stream(streams[1]):
        x = sts[1].be.empty((B // 2, N, N)); sts[1].lib.chs_dctn(sts[1]._h, sts[1].be.ptr(sts[1].T), sts[1].be.ptr(x)) if False else None
for it in range(K):
    for i, (s, st) in enumerate(zip(streams, sts)):
        with torch.cuda.stream(s):
            if offset and it == 0 and i == 1:
                torch.cuda._sleep(int(0.9e6))      # ~0.5 ms spin on stream 1
            st.steps(1)
for s in streams:
    ev = torch.cuda.Event(); ev.record(s); torch.cuda.default_stream().wait_event(ev)
e1.record(torch.cuda.default_stream())
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"cosched B={B} K={K} offset={offset} CTAS_PER_SM={os.environ.get('CHS_CTAS_PER_SM')} {ms/K*1e3:8.1f} us/step {B*K/(ms*1e-3):10.0f} sim-steps/s")
