"""GPU parity tests proper: the CUDA path (through the C ABI / public Solver API) against
the golden fixtures frozen from the unmodified reference, and against the CPU oracle.

Tolerances: U max-abs <= 1e-11; E/E2 and the other TimeData columns
rel <= 1e-9 per row; tau0 / computed_steps / stop_reason exact; t0 rel <= 1e-12;
SA within 2/N^2 (threshold ties)."""
import json
import os

import numpy as np
import pytest
import scipy.fftpack as fp

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
U_TOL = 1e-11
ROW_RTOL = 1e-9


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return z, json.loads(str(z["meta"]))


def make_params(meta):
    import chsimpy_b200 as ch
    p = ch.Parameters()
    p.no_gui = True
    for k, v in meta["params"].items():
        setattr(p, k, v)
    if meta["kappa_override"] is not None:
        p.kappa_tilde = meta["kappa_override"]
    if meta["fac"] is not None:
        f0, f1 = meta["fac"]
        p.func_A0 = lambda T: ch.utils.A0(T) * f0
        p.func_A1 = lambda T: ch.utils.A1(T) * f1
    return p


def check_rows(rows, ref, N):
    assert rows.shape == ref.shape, (rows.shape, ref.shape)
    for c, name in enumerate(("it", "E", "E2", "SA", "domtime", "Ra", "L2", "PS", "delt")):
        a, b = rows[:, c], ref[:, c]
        if name == "SA":
            assert np.abs(a - b).max() <= 2.0 / N ** 2 + 1e-15, name
        else:
            err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
            err[b == 0] = np.abs(a[b == 0])
            assert err.max() <= ROW_RTOL, (name, float(err.max()), int(err.argmax()))


def run_case(name):
    import chsimpy_b200 as ch
    z, m = load(name)
    p = make_params(m)
    s = ch.Solver(p)
    assert abs(s.U_init.sum() - m["U_init_sum"]) < 1e-9
    s.prepare()
    nan_raised = False
    try:
        if m["chunks"]:
            for i, c in enumerate(m["chunks"]):
                sol = s.solve_or_resume(c)
                key = f"U_chunk{i}"
                if key in z:
                    assert np.abs(sol.U - z[key]).max() <= U_TOL
                elif key + "_sample" in z:
                    st = max(1, p.N // 64)
                    assert np.abs(sol.U[::st, ::st] - z[key + "_sample"]).max() <= U_TOL
        else:
            sol = s.solve_or_resume(p.ntmax)
    except AssertionError:
        if m["nan_row"] < 0:
            raise
        nan_raised = True
        sol = s.solution
    rows, ref = sol.timedata.data(), z["rows"]
    if m["nan_row"] >= 0:
        assert nan_raised
        assert rows.shape[0] == ref.shape[0] and rows.shape[0] - 1 == m["nan_row"]
        check_rows(rows[:-1], ref[:-1], p.N)
        assert np.isnan(rows[-1]).any()
        return s, sol, z, m
    check_rows(rows, ref, p.N)
    assert sol.stop_reason == m["stop_reason"]
    assert sol.computed_steps == m["computed_steps"]
    assert sol.tau0 == m["tau0"]
    assert abs(sol.t0 - m["t0"]) <= 1e-12 * max(1.0, abs(m["t0"]))
    assert int(np.argmax(sol.E2)) == m["argmax_E2"]
    assert abs(s.time_passed - m["time_passed"]) <= 1e-12 * max(1.0, m["time_passed"])
    assert abs(s.delt - m["delt_final"]) <= 1e-12 * m["delt_final"]
    assert bool(s.skip_check) == m["skip_check"]
    st = max(1, p.N // 64)
    assert np.abs(sol.U[::st, ::st] - z["U_sample"]).max() <= U_TOL
    assert np.abs(sol.U.sum(axis=1) - z["U_rowsum"]).max() <= U_TOL * p.N
    if "U" in z:
        d = np.abs(sol.U - z["U"])
        rel_l2 = np.linalg.norm(sol.U - z["U"]) / np.linalg.norm(z["U"])
        assert d.max() <= U_TOL and rel_l2 <= 1e-12, (float(d.max()), float(rel_l2))
    return s, sol, z, m


@pytest.mark.parametrize("N", [32, 64, 128, 256, 512, 1024])
def test_dctn_matches_scipy(N):
    from chsimpy_b200 import _lib
    from chsimpy_b200.solver import BatchStepper
    ps = _lib.Params(RT=1, BRT=1, B=1, A0=1, A1=1, Amr=1, kappa_tilde=1, L=2, delx=2 / (N - 1), delt=1e-8,
                     delt_max=1e-8, M_tilde=1, threshold=0.5, time_limit_s=0, jitter=0, full_sim=1, adaptive_time=0)
    b = 3
    st = BatchStepper(N, [ps] * b)
    x = np.random.default_rng(N).random((b, N, N)) - 0.3
    y = st.dctn(x)
    ref = np.stack([fp.dctn(x[i], norm="ortho") for i in range(b)])
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) <= 1e-13
    back = st.dctn(ref, inverse=True)
    assert np.abs(back - x).max() <= 1e-13


SMALL = ["n32_k60", "n64_k200", "n128_k200", "n256_k200", "n64_lcg_k100", "n128_sobol_k100",
         "n256_T900_k300", "n64_cinit089_stop", "n128_chunked_jitter", "n256_chunked_adaptive", "n1024_k50"]
N512 = ["n512_stop", "n512_full2000", "n512_jitter700", "n512_adaptive1200", "n512_adaptive_default_nan",
        "n512_jitter_adaptive1200", "n512_jitter_stop", "n512_chunked_3x100", "n512_timelimit",
        "n512_corner_lo_lo", "n512_corner_lo_hi", "n512_corner_hi_lo", "n512_corner_hi_hi"]


@pytest.mark.parametrize("name", SMALL)
def test_golden_small(name):
    run_case(name)


@pytest.mark.parametrize("name", N512)
def test_golden_n512(name):
    run_case(name)


def test_default_stop_matches_survey_values():
    """The headline numbers of SURVEY.md 6 (config 2): stop at 1674, t0, argmax(E2)."""
    s, sol, z, m = run_case("n512_stop")
    assert sol.stop_reason == "energy" and sol.computed_steps == 1674 and sol.tau0 == 1674
    assert abs(sol.t0 - 2935.0877192982052) < 1e-9
    assert int(np.argmax(sol.E2)) == 1672


def test_against_oracle_seeded():
    """CUDA path vs oracle on fresh seeded inputs (not in the fixtures)."""
    import ch_oracle as orc
    import chsimpy_b200 as ch
    for N, seed, steps in ((64, 7, 120), (128, 11, 80), (256, 5, 40)):
        p = ch.Parameters()
        p.N, p.seed, p.ntmax, p.full_sim, p.no_gui, p.kappa_tilde = N, seed, steps, True, True, 2.5e-4
        s = ch.Solver(p)
        s.prepare()
        sol = s.solve_or_resume(steps)
        o = orc.run_default(N=N, nsteps=steps, seed=seed, kappa_tilde=2.5e-4, full_sim=True)
        check_rows(sol.timedata.data(), o.rows, N)
        assert np.abs(sol.U - o.U).max() <= U_TOL


def test_user_field_and_wrong_shape():
    import chsimpy_b200 as ch
    p = ch.Parameters()
    p.N, p.no_gui, p.kappa_tilde, p.full_sim = 64, True, 3e-4, True
    U0 = 0.875 + 0.004 * (np.random.default_rng(3).random((64, 64)) - 0.5)
    s = ch.Solver(p, U_init=U0)
    assert s.create_rand is None
    s.prepare()
    sol = s.solve_or_resume(10)
    assert sol.computed_steps == 10
    with pytest.raises(SystemExit):
        ch.Solver(p, U_init=np.zeros((8, 8)))


def test_properties_full_size():
    """Size-independent properties at N=512, batch 8: mean conservation (quirk Q4),
    dctn/idctn round trip, lock-step batch == single run, determinism."""
    import chsimpy_b200 as ch
    from chsimpy_b200.solver import BatchStepper, make_params_struct
    p = ch.Parameters()
    p.no_gui, p.full_sim, p.ntmax = True, True, 60
    single = ch.Solver(p)
    single.prepare()
    sol = single.solve_or_resume(60)
    assert abs(sol.U.mean() - single.U_init.mean()) < 1e-14
    ps = make_params_struct(p, single.solution)
    st = BatchStepper(512, [ps] * 8)
    st.set_U(single.U_init)
    st.prepare()
    rows, done = st.run(59)
    for i in range(8):
        assert np.array_equal(rows[i], sol.timedata.data()[1:]), i      # bit-identical across the batch
    U = st.get_U()
    for i in range(8):
        assert np.array_equal(U[i], sol.U)
    y = st.dctn(U)
    back = st.dctn(y, inverse=True)
    assert np.abs(back - U).max() < 1e-14
    # Parseval: orthonormal transform keeps the Frobenius norm
    assert abs(np.linalg.norm(y[0]) - np.linalg.norm(U[0])) < 1e-10


def test_ensemble_corners_and_centre():
    """experiment.py path: 5 members in lock-step (4 corners + centre), each with its own sympy
    kappa_tilde, stop step, t0, tsep and c_A/c_B -- against the reference's single runs."""
    import chsimpy_b200 as ch
    from chsimpy_b200 import experiment as ex
    rv = np.array([[0.995, 0.995], [0.995, 1.005], [1.005, 0.995], [1.005, 1.005], [1.0, 1.0]])
    names = ["n512_corner_lo_lo", "n512_corner_lo_hi", "n512_corner_hi_lo", "n512_corner_hi_hi", "n512_stop"]
    p = ch.Parameters()
    p.no_gui = True
    p.file_id = "ens"
    res = ex.solve_ensemble(p, rv, None)
    for r, name in zip(res, names):
        z, m = load(name)
        sol = r["solution"]
        assert sol.kappa_tilde == m["kappa_tilde"]
        assert (sol.tau0, sol.computed_steps, sol.stop_reason) == (m["tau0"], m["computed_steps"], m["stop_reason"])
        assert abs(sol.t0 - m["t0"]) <= 1e-12 * m["t0"]
        assert r["tuple"][8] == m["argmax_E2"] and r["tuple"][9] == r["run_id"]
        check_rows(sol.timedata.data(), z["rows"], 512)
        st = 8
        assert np.abs(sol.U[::st, ::st] - z["U_sample"]).max() <= U_TOL
    ca, cb = res[4]["tuple"][2], res[4]["tuple"][3]
    assert abs(ca - 0.8121353) < 5e-7 and abs(cb - 0.9723917) < 5e-7


@pytest.mark.parametrize("name,force", [("n2048_k10", False), ("n256_k200", True), ("n1024_k50", True),
                                        ("n4096_k6", False), ("n8192_k4", False)])
def test_slab_path_single_gpu(name, force):
    """Row-slab path (chs_slab.cuh) on one GPU: automatically for N > 1024, forced for smaller N."""
    import chsimpy_b200 as ch
    from chsimpy_b200.slab import SlabEngine
    z, m = load(name)
    p = make_params(m)
    s = ch.Solver(p, _force_slab=force)
    assert isinstance(s._stepper, SlabEngine)
    s.prepare()
    sol = s.solve_or_resume(p.ntmax)
    check_rows(sol.timedata.data(), z["rows"], p.N)
    st = max(1, p.N // 64)
    assert np.abs(sol.U[::st, ::st] - z["U_sample"]).max() <= U_TOL
    assert np.abs(sol.U.sum(axis=1) - z["U_rowsum"]).max() <= U_TOL * p.N
    assert sol.computed_steps == m["computed_steps"]


def test_slab_jitter_adaptive_n2048_golden():
    """--jitter + --adaptive-time on the slab path (N > 1024) against the unmodified reference's run
    (tests/golden/n2048_jitter_adaptive.npz): rows, the delt sequence, and the field."""
    import chsimpy_b200 as ch
    z, m = load("n2048_jitter_adaptive")
    p = make_params(m)
    s = ch.Solver(p)
    s.prepare()
    sol = s.solve_or_resume(p.ntmax)
    assert sol.computed_steps == m["computed_steps"]
    check_rows(sol.timedata.data(), z["rows"], p.N)
    assert float(sol.delt[-1]) > p.delt and abs(s.delt - m["delt_final"]) <= 1e-12 * m["delt_final"]
    st = p.N // 64
    assert np.abs(sol.U[::st, ::st] - z["U_sample"]).max() <= U_TOL


@pytest.mark.parametrize("kw,steps", [(dict(jitter=0.004), 40), (dict(adaptive_time=True, delt_max=1.2e-9), 560),
                                      (dict(jitter=0.003, adaptive_time=True, delt_max=1.2e-9), 530)])
def test_slab_jitter_adaptive_forced_vs_oracle(kw, steps):
    """The same features on the forced slab path at N=128 against the oracle, with a re-entry."""
    import ch_oracle as orc
    import chsimpy_b200 as ch
    p = ch.Parameters()
    p.N, p.no_gui, p.full_sim, p.kappa_tilde, p.seed, p.ntmax = 128, True, True, 2.7e-4, 5, steps
    for k, v in kw.items():
        setattr(p, k, v)
    s = ch.Solver(p, _force_slab=True)
    s.prepare()
    s.solve_or_resume(steps - 7)
    sol = s.solve_or_resume(7)
    o = orc.run_default(N=128, nsteps=steps - 7, seed=5, kappa_tilde=2.7e-4, full_sim=True, **kw)
    o.run(7)
    assert sol.computed_steps == o.computed_steps == steps
    check_rows(sol.timedata.data(), o.rows, 128)
    assert np.abs(sol.U - o.U).max() <= U_TOL
    assert abs(s.delt - o.delt) <= 1e-12 * o.delt


def test_slab_path_energy_stop():
    """The slab path (default for N > 1024, where full_sim defaults to False) must honour the device-side
    stop flag: the 8280-step run to the energy stop, 128 steps queued per host poll, forced onto the slab
    kernels -- computed_steps, tau0, t0, the row count and U are those of the stopping step."""
    import chsimpy_b200 as ch
    z, m = load("n64_cinit089_stop")
    p = make_params(m)
    s = ch.Solver(p, _force_slab=True)
    s.prepare()
    sol = s.solve_or_resume(p.ntmax)
    assert sol.stop_reason == "energy" and sol.computed_steps == m["computed_steps"] == 8280
    assert sol.tau0 == m["tau0"] and abs(sol.t0 - m["t0"]) <= 1e-12 * m["t0"]
    assert sol.timedata.data().shape[0] == z["rows"].shape[0]
    check_rows(sol.timedata.data(), z["rows"], 64)
    assert np.abs(sol.U - z["U"]).max() <= U_TOL


def test_cli_and_simulator_chunked(tmp_path, monkeypatch):
    """`python -m chsimpy_b200`-style flow (reference __main__.py:8-25) and the chunked
    `update_every` path of Simulator.solve (simulator.py:56-87) incl. csv/yaml export."""
    import chsimpy_b200 as ch
    monkeypatch.chdir(tmp_path)
    p = ch.CLIParser().get_parameters(["-N", "128", "-n", "61", "--full-sim", "--no-gui", "-f", "t1",
                                       "--export-csv", "U,E2", "--yaml", "-K", "3e-4"])
    sim = ch.Simulator(p)
    sol = sim.solve()
    assert sol.computed_steps == 61 and sol.stop_reason == "None"
    sim.export()
    U = ch.utils.csv_import_matrix("t1.solution.U.csv")
    assert np.allclose(U, sol.U, rtol=0, atol=1e-15)
    assert os.path.exists("t1.solution.yaml") and os.path.exists("t1.solution.E2.csv")
    # the module entry point itself, in-process
    from chsimpy_b200 import __main__ as cli_main
    sol_cli = cli_main.run(["-N", "128", "-n", "61", "--full-sim", "--no-gui", "-K", "3e-4"])
    assert sol_cli.computed_steps == 61 and np.array_equal(sol_cli.E, sol.E)
    # chunked: a view is "required" (png) -> headless stand-in, solve in chunks of 20
    p2 = ch.CLIParser().get_parameters(["-N", "128", "-n", "60", "--full-sim", "--no-gui", "--png", "--update-every", "20",
                                        "-K", "3e-4"])
    with pytest.warns(UserWarning):
        sim2 = ch.Simulator(p2)
    sol2 = sim2.solve()
    assert sim2.steps_total == 60 and sol2.computed_steps == 60
    # without jitter a chunked run equals the unchunked one to rounding (re-entry recomputes hat_U)
    assert np.abs(sol2.E - sol.E[:60]).max() / abs(sol.E[0]) < 1e-12


def test_gemm_path_n100_golden():
    """DCT-as-GEMM path (FP64 tensor cores, chs_gemm.cuh) at the reference's benchmark smoke size."""
    from chsimpy_b200 import _lib
    assert _lib.load().chs_uses_gemm(100, 1) == 1 and _lib.load().chs_uses_gemm(512, 1) == 0
    run_case("n100_k100")


@pytest.mark.parametrize("N,kw", [(40, dict()), (72, dict(jitter=0.004)), (56, dict(adaptive_time=True, delt_max=3e-9)),
                                  (24, dict(time_max=0.5))])
def test_gemm_path_vs_oracle(N, kw):
    """Arbitrary N, with jitter / adaptive dt / time limit, in-kernel time loop vs the oracle."""
    import ch_oracle as orc
    import chsimpy_b200 as ch
    steps = 620 if kw.get("adaptive_time") else 80
    p = ch.Parameters()
    p.N, p.ntmax, p.full_sim, p.no_gui, p.kappa_tilde, p.seed = N, steps, True, True, 2.7e-4, 5
    for k, v in kw.items():
        setattr(p, k, v)
    s = ch.Solver(p)
    s.prepare()
    sol = s.solve_or_resume(steps)
    o = orc.run_default(N=N, nsteps=steps, seed=5, kappa_tilde=2.7e-4, full_sim=True, **kw)
    check_rows(sol.timedata.data(), o.rows, N)
    assert np.abs(sol.U - o.U).max() <= U_TOL
    assert sol.stop_reason == o.stop_reason and sol.computed_steps == o.computed_steps
    assert abs(s.delt - o.delt) <= 1e-12 * o.delt


def test_gemm_dctn_matches_scipy():
    from chsimpy_b200 import _lib
    from chsimpy_b200.solver import BatchStepper
    for N in (8, 30, 100):
        ps = _lib.Params(RT=1, BRT=1, B=1, A0=1, A1=1, Amr=1, kappa_tilde=1, L=2, delx=2 / (N - 1), delt=1e-8,
                         delt_max=1e-8, M_tilde=1, threshold=0.5, time_limit_s=0, jitter=0, full_sim=1, adaptive_time=0)
        st = BatchStepper(N, [ps] * 2)
        x = np.random.default_rng(N).random((2, N, N)) - 0.3
        ref = np.stack([fp.dctn(x[i], norm="ortho") for i in range(2)])
        assert np.linalg.norm(st.dctn(x) - ref) / np.linalg.norm(ref) <= 1e-13
        assert np.abs(st.dctn(ref, inverse=True) - x).max() <= 1e-13


def test_device_pcg64_is_numpy_bit_exact():
    """The jitter noise is numpy's PCG64 stream reproduced on the device (no PCIe traffic)."""
    from chsimpy_b200 import _lib
    from chsimpy_b200.solver import BatchStepper
    N = 64
    ps = _lib.Params(RT=1, BRT=1, B=1, A0=1, A1=1, Amr=1, kappa_tilde=1, L=2, delx=2 / (N - 1), delt=1e-8,
                     delt_max=1e-8, M_tilde=1, threshold=0.5, time_limit_s=0, jitter=0, full_sim=1, adaptive_time=0)
    st = BatchStepper(N, [ps])
    g = np.random.Generator(np.random.PCG64(2023))
    g.random((N, N))                                   # the U_init draw comes first (solver.py:78-82)
    noise, mean = st.pcg64_noise(g.bit_generator.state, 5)
    ref = np.stack([g.random((N, N)) for _ in range(5)])
    assert np.array_equal(noise.cpu().numpy(), ref)
    assert np.abs(mean.cpu().numpy() - ref.reshape(5, -1).mean(axis=1)).max() < 1e-15


def test_experiment_main_end_to_end(tmp_path, monkeypatch):
    """`python -m chsimpy_b200.experiment` flow (reference experiment.py:129-233): factor table,
    batched solve to the energy stop, per-run exports, results + aggregate CSVs."""
    import pandas as pd
    from chsimpy_b200 import experiment as ex
    monkeypatch.chdir(tmp_path)
    ex.main(["-N", "64", "-R", "6", "--A-seed", "85972", "--cinit", "0.89", "--threshold", "0.89", "-n", "400",
             "-f", "exp", "--export-csv", "E2,SA", "-P", "2"])
    df = pd.read_csv("exp-results.csv", index_col=0)
    assert list(df.columns) == ex.RESULT_COLUMNS and len(df) == 6 and sorted(df["id"]) == list(range(6))
    want = np.random.Generator(np.random.PCG64(85972)).uniform(0.995, 1.005, size=(6, 2))
    assert np.allclose(df.sort_values("id")[["fac_A0", "fac_A1"]].values, want, rtol=0, atol=1e-15)
    assert (df["ca"] < df["cb"]).all() and (df["sa"] < df["sb"]).all() and (df["tau0"] >= 0).all()
    agg = pd.read_csv("exp-results-agg.csv", index_col=0)
    assert "cv" in agg.columns and "mean" in agg.columns
    assert os.path.exists("exp-run3.solution.yaml") and os.path.exists("exp-run3.solution.E2.csv")
    assert os.path.exists("exp-metadata.csv")
    # values, not only shapes: every member against its own single-Solver run and the oracle's run of the same member
    import ch_oracle as orc
    import chsimpy_b200 as ch
    rows = df.sort_values("id").reset_index(drop=True)
    for rid in (0, 3, 5):
        f0, f1 = float(want[rid, 0]), float(want[rid, 1])
        r = rows.iloc[rid]
        assert abs(r["A0"] - ch.utils.A0(923.15) * f0) <= 1e-12 * abs(r["A0"]) and abs(r["A1"] - ch.utils.A1(923.15) * f1) <= 1e-12 * abs(r["A1"])
        o = orc.run_default(N=64, c0=0.89, fac_A0=f0, fac_A1=f1, nsteps=400)
        assert int(r["tau0"]) == int(o.tau0) and abs(r["t0"] - o.t0) <= 1e-12 * max(1.0, o.t0), (rid, r["tau0"], o.tau0)
        assert int(r["tsep"]) == int(np.argmax(o.rows[:, 2]))
        ca, cb = ch.utils.get_miscibility_gap(ch.Parameters().R, 923.15, ch.Parameters().B, r["A0"], r["A1"])
        assert abs(r["ca"] - float(ca)) < 1e-12 and abs(r["cb"] - float(cb)) < 1e-12
        e2 = ch.utils.csv_import_matrix(f"exp-run{rid}.solution.E2.csv")
        ref = o.rows[:, 2]
        assert e2.shape == ref.shape and np.abs(e2 - ref).max() <= 1e-9 * np.abs(ref).max()
    assert abs(agg.loc["tau0", "mean"] - rows["tau0"].mean()) < 1e-9 and abs(agg.loc["ca", "cv"] - rows["ca"].std() / rows["ca"].mean()) < 1e-12


def test_uinit_file_and_nan_field(tmp_path):
    """--Uinit-file restart path (simulator.py:21-22) and the NaN assertion (timedata.py:10) when
    the field leaves (0,1)."""
    import chsimpy_b200 as ch
    p = ch.Parameters()
    p.N, p.no_gui, p.kappa_tilde, p.full_sim, p.ntmax = 64, True, 3e-4, True, 30
    s = ch.Simulator(p)
    sol = s.solve()
    f = str(tmp_path / "U.csv")
    ch.utils.csv_export_matrix(sol.U, f)
    p2 = p.deepcopy()
    p2.Uinit_file, p2.ntmax = f, 10
    s2 = ch.Simulator(p2)
    assert s2.solver.create_rand is None and np.abs(s2.solver.U_init - sol.U).max() < 1e-15
    sol2 = s2.solve()
    assert sol2.computed_steps == 10
    bad = sol.U.copy()
    bad[5, 7] = 1.25                          # log(1-U) of a negative number
    s3 = ch.Solver(p, U_init=bad)
    with pytest.raises(AssertionError):
        s3.prepare()


@pytest.mark.gpu
def test_plain_cpp_client_of_the_c_abi(tmp_path):
    """examples/c_abi_demo.cpp drives libchs_b200.so from C++ alone (dlopen + cudaMalloc): the
    reference's `-g lcg` run of tests/golden/n64_lcg_k100.npz, row for row."""
    import subprocess
    from test_abi import build_c_demo
    from chsimpy_b200 import _lib
    z, m = load("n64_lcg_k100")
    exe = build_c_demo(tmp_path)
    args = [exe, _lib.LIB_PATH, str(m["N"]), str(m["ntmax"]), str(m["seed"])] + \
           [repr(float(v)) for v in (m["RT"], m["BRT"], 12.86, m["A0"], m["A1"], m["Amr"], m["kappa_tilde"])]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = np.array([[float(x) for x in ln.split()] for ln in r.stdout.strip().splitlines()])
    assert "computed_steps=100" in r.stderr
    check_rows(rows, z["rows"], m["N"])


@pytest.mark.gpu
def test_slab_two_gpus_peer_memory_and_nccl():
    """Row-slab decomposition over two GPUs (needs a box with >= 2): tools/slab_check.py replays the
    reference fixtures n2048_k10 and n8192_k4 on both ranks, once with the peer-memory transposes,
    once with the NCCL all-to-all route and once with the copy-engine exchange (two row chunks)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in (dict(CHS_SLAB_P2P="1"), dict(CHS_SLAB_P2P="0"), dict(CHS_SLAB_CE="1", CHS_SLAB_CHUNKS="2")):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", "29577",
                            os.path.join(root, "tools", "slab_check.py"), "2048", "5"],
                           capture_output=True, text=True, timeout=900, env=env, cwd=root)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert r.stdout.count("[slab parity] world=2") == 2


@pytest.mark.gpu
def test_fast_log_accuracy_on_device():
    """The device's table-driven log (csrc/fastlog.cuh) against 120-bit mpmath: <= 1.25 ulp over (0.005, 0.995)."""
    from chsimpy_b200.solver import BatchStepper, make_params_struct
    import chsimpy_b200 as ch
    from test_emu_kernels import log_errors
    p = ch.Parameters()
    p.N, p.kappa_tilde = 32, 3e-4
    st = BatchStepper(32, [make_params_struct(p, ch.Solution(p))])
    rng = np.random.default_rng(7)
    x = rng.uniform(0.005, 0.995, 4000)
    ulp_err, abs_err = log_errors(st, x)
    big = np.abs(np.log(x)) >= 2.0 ** -7
    assert ulp_err[big].max() <= 1.25 and abs_err[~big].max() <= 1e-18
    y = st.debug_log(np.array([0.0, -1.0, np.nan]))
    assert y[0] == -np.inf and np.isnan(y[1]) and np.isnan(y[2])


def test_device_lcg_n512_bitexact():
    """k_lcg_fill (one device thread, float64 recurrence) == the host restatement of reference mport.py:8-32,
    bit for bit, at the size of `-g lcg -N 512` (and the reference's 5 x 4 known-answer vector)."""
    from chsimpy_b200 import mport
    from chsimpy_b200.solver import _CudaBackend, lcg_sample
    be = _CudaBackend()
    assert np.array_equal(lcg_sample(be, 512, 512, 2023), mport.matlab_lcg_sample(512, 512, 2023))
    known0 = [0.5475444293336684, 0.29257702841077793, 0.3117376865408093, 0.9844947126621821]
    assert np.allclose(lcg_sample(be, 5, 4, 2023)[0], known0, rtol=0, atol=1e-15)


@pytest.mark.parametrize("N,kw,steps", [(200, dict(), 120), (300, dict(jitter=0.003, adaptive_time=True, delt_max=3.4e-10), 524),
                                        (768, dict(), 12), (105, dict(time_max=0.4), 60)])
def test_arbitrary_n_path_vs_oracle(N, kw, steps):
    """Any N (reference cli_parser.py:27): sizes the FFT kernels do not take run as FP64 tensor-core GEMMs
    (chs_big.cuh, BigEngine), incl. jitter / adaptive dt / time limit, against the oracle with a re-entry."""
    import ch_oracle as orc
    import chsimpy_b200 as ch
    from chsimpy_b200.slab import BigEngine
    p = ch.Parameters()
    p.N, p.no_gui, p.full_sim, p.kappa_tilde, p.seed, p.ntmax = N, True, True, 2.7e-4, 5, steps
    for k, v in kw.items():
        setattr(p, k, v)
    s = ch.Solver(p)
    assert isinstance(s._stepper, BigEngine)
    s.prepare()
    s.solve_or_resume(steps - 6)
    sol = s.solve_or_resume(6)
    o = orc.run_default(N=N, nsteps=steps - 6, seed=5, kappa_tilde=2.7e-4, full_sim=True, **kw)
    o.run(6)
    assert sol.computed_steps == o.computed_steps and sol.stop_reason == o.stop_reason
    check_rows(sol.timedata.data(), o.rows, N)
    assert np.abs(sol.U - o.U).max() <= U_TOL
    assert abs(s.delt - o.delt) <= 1e-12 * o.delt


@pytest.mark.gpu
def test_mixed_launches_and_one_tile_per_sm_kernels_bit_identical_on_device():
    """The scheduling variants of chs_steps on the device: mixed column+row launches (k_mix, throughput kernels) against
    the default, which for 5 members of N=64 (40 tiles <= #SMs) runs the unrolled `_LL` instantiations on exclusive SMs --
    TimeData rows, states and fields must agree bit for bit (same test body as on the host emulation)."""
    from test_emu_kernels import test_mixed_launches_are_bit_identical
    from chsimpy_b200.solver import _CudaBackend
    test_mixed_launches_are_bit_identical(_CudaBackend(None))
