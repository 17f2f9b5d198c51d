"""Phase timestamps (%globaltimer) of tile 0 of the two step kernels of ONE simulation: where the fixed latency of a
step goes.  Needs a library built with -DCHS_TRACE=1 (tools/build_variant.sh trace -DCHS_TRACE=1; CHS_B200_LIB=...)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import chsimpy_b200 as ch
from chsimpy_b200.solver import BatchStepper, make_params_struct

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
p = ch.Parameters(); p.N = N; p.no_gui = True; p.full_sim = True; p.kappa_tilde = 2.989112919661156e-4
s0 = ch.Solver(p)
st = BatchStepper(N, [make_params_struct(p, s0.solution)], rows_cap=512)
st.set_U(s0.U_init); st.prepare(); st.begin()
buf = torch.zeros(32, dtype=torch.int64, device="cuda")
st.lib.chs_debug_trace.restype = C.c_int
st.lib.chs_debug_trace.argtypes = [C.c_void_p, C.c_void_p]
st.lib.chs_debug_trace(st._h, buf.data_ptr())
names = {0: "col: after PDL wait", 1: "col: prologue issued", 2: "col: T tile landed", 3: "col: forward stages",
         4: "col: fused update pass", 5: "col: inverse stages", 6: "col: tile stored",
         8: "row: after PDL wait", 9: "row: prologue issued", 10: "row: T tile landed", 11: "row: pre + first inverse stage",
         12: "row: inverse stage", 13: "row: last inverse + physics + first forward", 14: "row: sums + ticket",
         15: "row: forward stage", 16: "row: last forward + post", 17: "row: tile stored", 18: "row: end of tile 0",
         20: "row: last CTA enters step_control", 19: "row: step_control done (last CTA)"}
acc = {}
R = 50
for r in range(R):
    st.steps(3); torch.cuda.synchronize()
    t = buf.cpu().numpy().astype(np.int64)
    base = t[0]
    for k in names:
        acc.setdefault(k, []).append(int(t[k] - base))
print(f"N={N}: median ns since 'col: after PDL wait' of the same step (last of 3), {R} samples")
prev = 0
for k in sorted(names, key=lambda k: np.median(acc[k])):
    m = float(np.median(acc[k]))
    print(f"  {names[k]:48s} {m:9.0f} ns   (+{m - prev:7.0f})")
    prev = m
