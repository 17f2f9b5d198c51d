"""Host-side mirror of the reference interface: Parameters / TimeData / Solution scalars /
CLI / YAML+CSV round trips (modelled on reference tests/test.py:40-119) -- no GPU."""
import json
import os

import numpy as np
import pytest

import chsimpy_b200 as ch
from chsimpy_b200 import utils

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def meta(name):
    return json.loads(str(np.load(os.path.join(GOLD, name + ".npz"))["meta"]))


def test_parameter_defaults_match_reference():
    p = ch.Parameters()
    assert (p.N, p.L, p.XXX, p.threshold, p.seed, p.generator) == (512, 2, 0.875, 0.875, 2023, "uniform")
    assert p.temp == 650 + 273.15 and p.B == 12.86 and p.R == 0.0083144626181532
    assert (p.delt, p.delt_max, p.M_tilde, p.ntmax) == (3e-8, 9e-8, 1.71e-8, 1000000)
    assert p.kappa_tilde is None and p.jitter is None and p.time_max is None and p.update_every == 100
    assert not (p.full_sim or p.adaptive_time or p.no_gui or p.yaml or p.png)
    assert p.func_A0(923.15) == utils.A0(923.15) and p.func_A1(923.15) == utils.A1(923.15)


def test_parameters_yaml_roundtrip(tmp_path):
    f = str(tmp_path / "p.yaml")
    p1 = ch.Parameters()
    p1.func_A0 = lambda temp: 1 + 2 * temp            # ignored: not a scalar
    p1.N, p1.jitter = 256, 0.01
    p1.yaml_export_scalars(f)
    p2 = ch.Parameters()
    p2.yaml_import_scalars(f)
    assert p1.is_scalarwise_equal_with(p2)
    p1.N = 128
    assert not p1.is_scalarwise_equal_with(p2) and p2.N == 256


def test_solution_scalars_match_reference_default():
    m = meta("n512_stop")
    p = ch.Parameters()
    s = ch.Solution(p)
    assert s.kappa_tilde == m["kappa_tilde"] and float(s.kappa_base) == m["kappa_base"]
    assert (s.A0, s.A1, s.RT, s.BRT, s.Amr, s.delx) == (m["A0"], m["A1"], m["RT"], m["BRT"], m["Amr"], m["delx"])
    CH, S = s.CHeig, s.Seig
    assert CH.shape == (512, 512) and CH[0, 0] == 1.0 and S[0, 0] == 0.0


@pytest.mark.parametrize("name", ["n512_corner_lo_lo", "n512_corner_lo_hi", "n512_corner_hi_lo",
                                  "n512_corner_hi_hi", "n256_T900_k300", "n64_cinit089_stop"])
def test_kappa_matches_reference(name):
    """kappa_tilde inherits 7-digit nsolve rounding (quirk Q12): must be identical, not close."""
    m = meta(name)
    p = ch.Parameters()
    for k, v in m["params"].items():
        setattr(p, k, v)
    if m["fac"]:
        f0, f1 = m["fac"]
        p.func_A0 = lambda T: utils.A0(T) * f0
        p.func_A1 = lambda T: utils.A1(T) * f1
    assert ch.Solution(p).kappa_tilde == m["kappa_tilde"]


def test_miscibility_gap_and_spinodal():
    p = ch.Parameters()
    ca, cb = utils.get_miscibility_gap(p.R, p.temp, p.B, utils.A0(p.temp), utils.A1(p.temp))
    assert abs(float(ca) - 0.8121353) < 5e-7 and abs(float(cb) - 0.9723917) < 5e-7
    sa, sb = utils.get_roots_of_EPP(p.R, p.temp, utils.A0(p.temp), utils.A1(p.temp))
    assert abs(float(sa) - 0.854591765123637) < 1e-12 and abs(float(sb) - 0.949088448398765) < 1e-12


def test_timedata_semantics():
    td = ch.TimeData(capacity=2)
    for i, e2 in enumerate([1.0, 3.0, 2.0, 0.5]):
        td.insert(it=i, delt=3e-8, E=-1.0, E2=e2, SA=0.5, domtime=i, Ra=0.1, L2=0.2, PS=0.3)
    assert td.data().shape == (4, 9) and list(td.E2) == [1.0, 3.0, 2.0, 0.5]
    assert list(td.it_range) == [0, 1, 2, 3] and td.delt[0] == 3e-8
    assert td.energy_falls(2) is True and td.energy_falls(3) is False and td.energy_falls(1) is False
    with pytest.raises(AssertionError):
        td.insert(it=4, delt=3e-8, E=float("nan"), E2=1, SA=0, domtime=0, Ra=0, L2=0, PS=0)
    assert td.data().shape == (5, 9)                  # the NaN row is in the table when the assertion fires


def test_csv_roundtrip(tmp_path):
    a = np.random.default_rng(0).random((54, 33))
    for ext in ("csv", "csv.bz2"):
        f = str(tmp_path / ("m." + ext))
        utils.csv_export_matrix(a, f)
        assert np.allclose(a, utils.csv_import_matrix(f))


def test_cli_parser_flags_and_ranges():
    c = ch.CLIParser()
    p = c.get_parameters(["-N", "256", "-n", "2000", "--full-sim", "--no-gui", "-j", "0.01", "-a",
                          "--A0", "-150", "-t", "1.5", "-g", "lcg", "-s", "7", "--dt", "1e-8"])
    assert (p.N, p.ntmax, p.full_sim, p.no_gui, p.jitter, p.adaptive_time) == (256, 2000, True, True, 0.01, True)
    assert p.func_A0(1.0) == -150 and p.time_max == 1.5 and p.generator == "lcg" and p.seed == 7 and p.delt == 1e-8
    with pytest.raises(SystemExit):
        ch.CLIParser().get_parameters(["--cinit", "0.5"])
    with pytest.raises(SystemExit):
        ch.CLIParser().get_parameters(["--update-every", "1"])


def test_solver_needs_cuda_no_fallback():
    """Without a GPU the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    p = ch.Parameters()
    p.N, p.kappa_tilde = 64, 3e-4
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ch.Solver(p)


# ---- the reference's own YAML / CSV unit tests (reference tests/test.py:40-118), ported -------------
def test_ref_dump_parameters_scalars_roundtrip(tmp_path):
    f = str(tmp_path / "test-dump-parameters.yaml")
    p1 = ch.Parameters()
    p1.func_A0 = lambda temp: 1+2*temp  # is ignored as it is non-scalar
    p1.yaml_export_scalars(f)
    p2 = ch.utils.yaml_import(f)
    assert isinstance(p2, ch.Parameters) and p1.is_scalarwise_equal_with(p2)


def test_ref_dump_parameters_roundtrip_mismatch(tmp_path):
    f = str(tmp_path / "test-dump-parameters.yaml")
    p1 = ch.Parameters()
    p1.N = 512
    p1.yaml_export_scalars(f)
    p2 = ch.utils.yaml_import(f)
    p1.N = 256
    assert p1 != p2 and p2.N == 512 and p1.N == 256


def test_ref_dump_solution_scalars_roundtrip(tmp_path):
    f = str(tmp_path / "test-dump-solution.yaml")
    params = ch.Parameters()
    params.kappa_tilde = 2.989112919661156e-4          # (skips the 0.5 s sympy solve; any Parameters works)
    s1 = ch.Solution(params)
    s1.tau0, s1.t0, s1.computed_steps, s1.stop_reason = 1674, 2934.2, 1674, 'energy'
    s1.yaml_export_scalars(f)
    s2 = ch.utils.yaml_import(f)
    assert isinstance(s2, ch.Solution) and s1.is_scalarwise_equal_with(s2)
    assert s2.params.N == 512 and s2.tau0 == 1674 and s2.stop_reason == 'energy' and s2.A0 == s1.A0


def test_ref_dump_csv_roundtrip_and_compress(tmp_path):
    from chsimpy_b200 import mport
    f = str(tmp_path / "test-dump-lcg_matrix.csv")
    m = mport.matlab_lcg_sample(55, 34, 2023)
    ch.utils.csv_export_matrix(m, fname=f)
    assert np.allclose(m, ch.utils.csv_import_matrix(f))
    fz = str(tmp_path / "test-matrix.csv.bz2")
    r = np.random.default_rng(3).random((54, 33))
    ch.utils.csv_export_matrix(r, fz)
    assert np.allclose(r, ch.utils.csv_import_matrix(fz))
