"""Selection of the DCT variant per N by measurement: sim-steps/s of the FFT path vs the
DCT-as-GEMM path (CHS_FORCE_GEMM=1 routes FFT-capable N through the GEMM kernel)."""
import os, sys, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np, torch
    import chsimpy_b200 as ch
    from chsimpy_b200.solver import BatchStepper, make_params_struct
    N, B, K = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    p = ch.Parameters(); p.N = N; p.no_gui = True; p.full_sim = True; p.kappa_tilde = 2.989112919661156e-4
    s0 = ch.Solver(p); ps = make_params_struct(p, s0.solution)
    st = BatchStepper(N, [ps] * B, rows_cap=K + 40)
    st.set_U(s0.U_init); st.prepare(); st.begin(); st.steps(20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st.steps(K); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"N": N, "B": B, "gemm": bool(st.lib.chs_uses_gemm(N, B)), "us_per_step": ms / K * 1e3,
                      "sim_steps_per_s": B * K / (ms * 1e-3)}))
    sys.exit(0)
for N in (32, 64):
    for B in (1, 148, 1184):
        for force in ("0", "1"):
            env = dict(os.environ, CHS_FORCE_GEMM=force)
            out = subprocess.run([sys.executable, __file__, "child", str(N), str(B), "200"], env=env, capture_output=True, text=True)
            print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
for N, B in ((100, 1), (100, 148), (100, 1184), (48, 1184)):
    out = subprocess.run([sys.executable, __file__, "child", str(N), str(B), "200"], capture_output=True, text=True)
    print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
