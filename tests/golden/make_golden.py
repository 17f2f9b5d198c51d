#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (through
oracle/ref_shim.py) in the build container.  /root/reference does not travel to the
GPU box, so these frozen outputs are what the CPU and GPU parity tests replay.

    python tests/golden/make_golden.py [case ...]        # default: all cases, 8 procs

Each fixture stores: the case's parameters, the complete TimeData table, the stop
tuple (tau0, t0, stop_reason, computed_steps, argmax E2), the derived scalars
(A0, A1, kappa_tilde, ...), U snapshots, and whether oracle/ch_oracle.py reproduced
the reference bit-for-bit on that case (`oracle_bitexact`).
"""
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

# name -> dict(params overrides, extras)
CASES = {
    # BASELINE config 2: default run to the energy stop
    "n512_stop": dict(p=dict(), keep_U=True),
    # BASELINE config 1: -n 2000 --full-sim
    "n512_full2000": dict(p=dict(ntmax=2000, full_sim=True), keep_U=True),
    # ensemble corners (experiment.py:91-101)
    "n512_corner_lo_lo": dict(p=dict(), fac=(0.995, 0.995)),
    "n512_corner_lo_hi": dict(p=dict(), fac=(0.995, 1.005)),
    "n512_corner_hi_lo": dict(p=dict(), fac=(1.005, 0.995)),
    "n512_corner_hi_hi": dict(p=dict(), fac=(1.005, 1.005)),
    # config 4 pieces
    "n512_jitter700": dict(p=dict(ntmax=700, full_sim=True, jitter=0.01), keep_U=True),
    "n512_adaptive1200": dict(p=dict(ntmax=1200, full_sim=True, adaptive_time=True, delt_max=2e-10), keep_U=True),
    "n512_adaptive_default_nan": dict(p=dict(ntmax=1000, full_sim=True, adaptive_time=True), expect_nan=True),
    "n512_jitter_adaptive1200": dict(p=dict(ntmax=1200, full_sim=True, adaptive_time=True, delt_max=2e-10, jitter=0.01)),
    "n512_jitter_stop": dict(p=dict(jitter=0.01)),                       # quirk Q5: stops at 3
    # chunked re-entry (simulator.py:56-81 semantics; quirks Q2/Q3)
    "n512_chunked_3x100": dict(p=dict(full_sim=True), chunks=[100, 100, 100], keep_U=False),
    "n256_chunked_adaptive": dict(p=dict(N=256, full_sim=True, adaptive_time=True, delt_max=4e-10), chunks=[400, 150, 151]),
    "n128_chunked_jitter": dict(p=dict(N=128, full_sim=True, jitter=0.005), chunks=[20, 20, 20], keep_U=True),
    # time limit (solver.py:195-199, Q15)
    "n512_timelimit": dict(p=dict(time_max=1.0, full_sim=True), keep_U=False),
    # other sizes, k-step snapshots
    "n32_k60": dict(p=dict(N=32, ntmax=60, full_sim=True), keep_U=True, kappa=3e-4),
    "n64_k200": dict(p=dict(N=64, ntmax=200, full_sim=True), keep_U=True),
    "n128_k200": dict(p=dict(N=128, ntmax=200, full_sim=True), keep_U=True),
    "n256_k200": dict(p=dict(N=256, ntmax=200, full_sim=True), keep_U=True),
    "n1024_k50": dict(p=dict(N=1024, ntmax=50, full_sim=True), keep_U=False),
    "n2048_k10": dict(p=dict(N=2048, ntmax=10, full_sim=True), keep_U=False),
    # BASELINE config 5 sizes (the reference needs about 6.5 s and 8 GiB per step at N=8192)
    "n4096_k6": dict(p=dict(N=4096, ntmax=6, full_sim=True), keep_U=False),
    "n8192_k4": dict(p=dict(N=8192, ntmax=4, full_sim=True), keep_U=False),
    # jitter + adaptive dt on the slab path (N > 1024); delt_max scaled with 512/N (the column SUM makes delt_dyn ~ N, Q6)
    "n2048_jitter_adaptive": dict(p=dict(N=2048, ntmax=530, full_sim=True, adaptive_time=True, delt_max=5e-11, jitter=0.005),
                                  keep_U=False),
    "n100_k100": dict(p=dict(N=100, ntmax=100, full_sim=True), keep_U=True),     # benchmark.py -N 100 smoke size
    # other generators / user-supplied field
    "n64_lcg_k100": dict(p=dict(N=64, ntmax=100, full_sim=True, generator="lcg"), keep_U=True, keep_Uinit=True),
    "n128_sobol_k100": dict(p=dict(N=128, ntmax=100, full_sim=True, generator="sobol"), keep_U=True, keep_Uinit=True),
    "n64_cinit089_stop": dict(p=dict(N=64, XXX=0.89, threshold=0.89), keep_U=True),
    "n256_T900_k300": dict(p=dict(N=256, temp=900.0, ntmax=300, full_sim=True), keep_U=True),
}


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def run_case(name):
    import ref_shim
    ch = ref_shim.import_reference()
    import ch_oracle as orc
    spec = CASES[name]
    t_start = time.time()
    params = ch.Parameters()
    params.no_gui = True
    for k, v in spec["p"].items():
        setattr(params, k, v)
    if "kappa" in spec:
        params.kappa_tilde = spec["kappa"]
    fac = spec.get("fac")
    if fac is not None:
        f0, f1 = fac
        params.func_A0 = lambda T: ch.utils.A0(T) * f0
        params.func_A1 = lambda T: ch.utils.A1(T) * f1
    out = {}
    nan_row = -1
    sim = ch.Simulator(params)
    solver = sim.solver
    U_init = solver.U_init.copy()
    chunks = spec.get("chunks")
    snaps = {}
    try:
        if chunks is None:
            sol = sim.solve()
        else:
            # the reference only drives chunked solves from inside Simulator.solve, i.e.
            # under its single-thread BLAS cap (simulator.py:14,36); keep that here
            from threadpoolctl import threadpool_limits
            with threadpool_limits(limits=1, user_api="blas"):
                solver.prepare()
                for i, c in enumerate(chunks):
                    sol = solver.solve_or_resume(c)
                    snaps[f"U_chunk{i}"] = sol.U.copy()
    except AssertionError:
        if not spec.get("expect_nan"):
            raise
        sol = solver.solution
        nan_row = sol.timedata.data().shape[0] - 1
    rows = sol.timedata.data().copy()
    # ---- oracle cross-check (bit-for-bit) -------------------------------------------------
    k = orc.Consts.from_params(N=params.N, L=params.L, temp=params.temp, B=params.B, R=params.R,
                               N_A=params.N_A, delt=params.delt, delt_max=params.delt_max,
                               M_tilde=params.M_tilde, threshold=params.threshold,
                               kappa_tilde=sol.kappa_tilde, A0=sol.A0, A1=sol.A1)
    U0o, draw = orc.initial_field(params.N, params.XXX, params.generator, params.seed)
    os_ = orc.OracleSolver(k, U0o, full_sim=params.full_sim, adaptive_time=params.adaptive_time,
                           jitter=params.jitter, time_max=params.time_max, create_rand=draw)
    os_.prepare()
    o_nan = -1
    try:
        if chunks is None:
            os_.run(max(params.ntmax, 0))
        else:
            for c in chunks:
                os_.run(c)
    except AssertionError:
        o_nan = os_.rows.shape[0] - 1
    bitexact = (np.array_equal(U0o, U_init) and os_.rows.shape == rows.shape
                and np.array_equal(os_.rows, rows, equal_nan=True) and o_nan == nan_row)
    if nan_row < 0:
        bitexact = bitexact and np.array_equal(os_.U, sol.U) and os_.tau0 == sol.tau0 and os_.t0 == sol.t0 \
            and os_.stop_reason == sol.stop_reason and os_.computed_steps == sol.computed_steps
    if fac is None and "kappa" not in spec:
        kt, _ = orc.kappa_tilde_from_common_tangent(params.R, params.temp, params.B, sol.A0, sol.A1, params.XXX)
        bitexact = bitexact and (kt == sol.kappa_tilde)
    # ---- fixture --------------------------------------------------------------------------
    meta = dict(name=name, params={k: v for k, v in spec["p"].items()}, fac=fac, chunks=chunks,
                kappa_override=spec.get("kappa"),
                N=params.N, seed=params.seed, generator=params.generator, XXX=params.XXX,
                threshold=params.threshold, temp=params.temp, delt=params.delt, delt_max=params.delt_max,
                ntmax=params.ntmax, full_sim=params.full_sim, adaptive_time=params.adaptive_time,
                jitter=params.jitter, time_max=params.time_max,
                A0=float(sol.A0), A1=float(sol.A1), RT=float(sol.RT), BRT=float(sol.BRT),
                Amr=float(sol.Amr), delx=float(sol.delx), kappa_tilde=float(sol.kappa_tilde),
                kappa_base=float(getattr(sol, "kappa_base", float("nan"))),
                tau0=float(sol.tau0), t0=float(sol.t0), stop_reason=sol.stop_reason,
                computed_steps=int(sol.computed_steps), nan_row=int(nan_row),
                argmax_E2=int(np.nanargmax(rows[:, 2])),
                time_passed=float(solver.time_passed), delt_final=float(solver.delt),
                skip_check=bool(solver.skip_check),
                U_init_sha=_digest(U_init), U_init_sum=float(U_init.sum()),
                U_sha=_digest(sol.U), U_min=float(np.min(sol.U)), U_max=float(np.max(sol.U)),
                U_mean=float(np.mean(sol.U)),
                oracle_bitexact=bool(bitexact),
                numpy=np.__version__, scipy=__import__("scipy").__version__,
                sympy=__import__("sympy").__version__, seconds=round(time.time() - t_start, 2))
    out["meta"] = np.array(json.dumps(meta))
    out["rows"] = rows
    # a strided sample + row/col sums of the final field always travel; the full field only if asked
    out["U_rowsum"] = sol.U.sum(axis=1)
    out["U_colsum"] = sol.U.sum(axis=0)
    st = max(1, params.N // 64)
    out["U_sample"] = sol.U[::st, ::st].copy()
    if spec.get("keep_U"):
        out["U"] = sol.U
    if spec.get("keep_Uinit"):
        out["U_init"] = U_init
    for kname, v in snaps.items():
        if spec.get("keep_U"):
            out[kname] = v
        else:
            out[kname + "_sample"] = v[::st, ::st].copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return name, meta


def main():
    names = sys.argv[1:] or list(CASES)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    with mp.Pool(min(8, len(names))) as pool:
        for name, meta in pool.imap_unordered(run_case, names):
            print(f"{name:28s} steps={meta['computed_steps']:5d} stop={meta['stop_reason']:10s} "
                  f"tau0={meta['tau0']:.0f} nan_row={meta['nan_row']} oracle_bitexact={meta['oracle_bitexact']} "
                  f"({meta['seconds']} s)", flush=True)


if __name__ == "__main__":
    main()
