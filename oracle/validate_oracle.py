#!/usr/bin/env python
"""Pins oracle/ch_oracle.py against the UNMODIFIED reference, run live through
oracle/ref_shim.py (build container only: /root/reference does not exist on the GPU box).

    python oracle/validate_oracle.py            # a few quick cases, bit-for-bit comparison

tests/golden/make_golden.py performs the same comparison for every frozen fixture and stores
the verdict (`oracle_bitexact`) that tests/test_oracle.py asserts on."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ch_oracle as orc   # noqa: E402
import ref_shim           # noqa: E402


def compare(N, steps, **kw):
    ch = ref_shim.import_reference()
    p = ch.Parameters()
    p.no_gui, p.N, p.ntmax = True, N, steps
    for k, v in kw.items():
        setattr(p, k, v)
    sim = ch.Simulator(p)
    sol = sim.solve()
    k = orc.Consts.from_params(N=N, temp=p.temp, delt=p.delt, delt_max=p.delt_max, threshold=p.threshold,
                               kappa_tilde=sol.kappa_tilde, A0=sol.A0, A1=sol.A1)
    U0, draw = orc.initial_field(N, p.XXX, p.generator, p.seed)
    o = orc.OracleSolver(k, U0, full_sim=p.full_sim, adaptive_time=p.adaptive_time, jitter=p.jitter,
                         time_max=p.time_max, create_rand=draw)
    o.prepare()
    o.run(max(p.ntmax, 0))
    ok = (np.array_equal(o.rows, sol.timedata.data()) and np.array_equal(o.U, sol.U)
          and o.stop_reason == sol.stop_reason and o.computed_steps == sol.computed_steps)
    kt, _ = orc.kappa_tilde_from_common_tangent(p.R, p.temp, p.B, sol.A0, sol.A1, p.XXX)
    ok = ok and kt == sol.kappa_tilde
    print(f"N={N:4d} steps={steps:4d} {kw}: {'bit-identical' if ok else 'MISMATCH'}")
    return ok


if __name__ == "__main__":
    good = all([compare(64, 150, full_sim=True), compare(128, 80, full_sim=True, jitter=0.005),
                compare(64, 10 ** 6, XXX=0.89, threshold=0.89), compare(256, 60, full_sim=True, generator="sobol"),
                compare(100, 40, full_sim=True, generator="lcg")])
    sys.exit(0 if good else 1)
