"""The CPU oracle (oracle/ch_oracle.py) against the golden fixtures frozen from the
unmodified reference (tests/golden/make_golden.py), plus the reference's own LCG
known-answer vector (reference tests/test.py:19-37)."""
import glob
import json
import os

import numpy as np
import pytest

import ch_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FAST = ["n32_k60", "n64_k200", "n128_k200", "n64_lcg_k100", "n128_sobol_k100", "n128_chunked_jitter",
        "n256_T900_k300", "n100_k100", "n512_jitter_stop", "n512_timelimit"]


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return z, json.loads(str(z["meta"]))


def test_every_fixture_was_pinned_bit_exact():
    """make_golden.py ran oracle and reference side by side; every case must have matched."""
    names = sorted(glob.glob(os.path.join(GOLD, "*.npz")))
    assert len(names) >= 25
    for f in names:
        z = np.load(f)
        if "meta" not in z.files:            # factor_tables.npz: not a solver case
            continue
        m = json.loads(str(z["meta"]))
        assert m["oracle_bitexact"] is True, m["name"]


@pytest.mark.parametrize("name", FAST)
def test_oracle_replays_fixture(name):
    z, m = load(name)
    k = orc.Consts.from_params(N=m["N"], temp=m["temp"], delt=m["delt"], delt_max=m["delt_max"],
                               threshold=m["threshold"], kappa_tilde=m["kappa_tilde"], A0=m["A0"], A1=m["A1"])
    U0, draw = orc.initial_field(m["N"], m["XXX"], m["generator"], m["seed"])
    assert abs(U0.sum() - m["U_init_sum"]) < 1e-9
    s = orc.OracleSolver(k, U0, full_sim=m["full_sim"], adaptive_time=m["adaptive_time"], jitter=m["jitter"],
                         time_max=m["time_max"], create_rand=draw)
    s.prepare()
    for c in (m["chunks"] or [m["ntmax"]]):
        s.run(c)
    same_build = (m["numpy"] == np.__version__)
    if same_build:
        assert np.array_equal(s.rows, z["rows"])
    else:
        assert np.allclose(s.rows, z["rows"], rtol=1e-10, atol=0)
    assert s.stop_reason == m["stop_reason"] and s.computed_steps == m["computed_steps"]
    assert s.tau0 == m["tau0"] and abs(s.t0 - m["t0"]) <= 1e-12 * max(1.0, m["t0"])
    if "U" in z:
        assert np.abs(s.U - z["U"]).max() <= (0 if same_build else 1e-12)


def test_lcg_known_answer():
    """The only golden vector the reference's own tests hold (tests/test.py:25-37)."""
    want = [[0.5475444293336684, 0.29257702841077793, 0.3117376865408093, 0.9844947126621821],
            [0.8031704429551821, 0.03775238992541674, 0.37862920778739695, 0.5387215616827465],
            [0.7217314246677474, 0.7984879318617694, 0.8011069301520972, 0.8502945903922872],
            [0.5455620291389348, 0.34767496602035824, 0.8863348965003783, 0.8019890788951838],
            [0.9676096443867356, 0.12967026239711338, 0.008214473728190397, 0.4722352030092083]]
    assert np.allclose(orc.lcg_field(5, 4, 2023), want)
    from chsimpy_b200 import mport
    assert np.allclose(mport.matlab_lcg_sample(5, 4, 2023), want)
    assert np.array_equal(mport.matlab_lcg_sample(55, 34, 2023), orc.lcg_field(55, 34, 2023))


def test_survey_headline_values():
    """SURVEY.md section 6 numbers are what the frozen default run holds."""
    z, m = load("n512_stop")
    assert m["stop_reason"] == "energy" and m["computed_steps"] == 1674 and m["tau0"] == 1674
    assert abs(m["t0"] - 2935.0877192982052) < 1e-9 and m["argmax_E2"] == 1672
    assert m["U_init_sha"] == "0cd95153c30bce0e"
    assert abs(m["kappa_tilde"] - 0.0002989112919661156) < 1e-18
    for name, tau in (("lo_lo", 1695), ("lo_hi", 1557), ("hi_lo", 1985), ("hi_hi", 1668)):
        assert load("n512_corner_" + name)[1]["tau0"] == tau
    assert load("n512_adaptive_default_nan")[1]["nan_row"] == 504
    assert load("n512_jitter_stop")[1]["computed_steps"] == 3
