"""Where the wall time of an ensemble-to-stop run goes (1024 members, poll every 128 steps): device time of the
step launches (CUDA events) vs host time of poll / take_rows per chunk."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import chsimpy_b200 as ch
import bench
from chsimpy_b200.solver import BatchStepper, make_params_struct
members = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
fac, kap = bench.member_scalars(members)
structs = []
for f, k in zip(fac, kap):
    p = ch.Parameters(); p.no_gui, p.kappa_tilde = True, float(k)
    p.func_A0 = (lambda f0: (lambda T: ch.utils.A0(T) * f0))(float(f[0]))
    p.func_A1 = (lambda f1: (lambda T: ch.utils.A1(T) * f1))(float(f[1]))
    structs.append(make_params_struct(p, ch.Solution(p)))
U0 = 0.875 + 0.875 * 0.01 * (np.random.Generator(np.random.PCG64(2023)).random((512, 512)) - 0.5)
st = BatchStepper(512, structs, rows_cap=128)
st.set_U(U0); st.prepare(); torch.cuda.synchronize()
t0 = time.perf_counter()
st.begin()
tg = tp = tr = 0.0
running, chunks, work = st.batch, 0, 0
while running > 0:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st.steps(128); e1.record(); torch.cuda.synchronize()
    tg += e0.elapsed_time(e1) * 1e-3
    a = time.perf_counter(); prev = running; running, _, _, rw = st.poll(); b = time.perf_counter()
    work += int(rw.sum())
    got = st.take_rows(); c = time.perf_counter()
    tp += b - a; tr += c - b; chunks += 1
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{members} members: wall {dt:.3f} s, device (events around steps) {tg:.3f} s, poll {tp:.3f} s, take_rows {tr:.3f} s, {chunks} chunks, "
      f"{work} member-steps = {work/tg/1e3:.1f} k/s on the device, {work/dt/1e3:.1f} k/s wall")
