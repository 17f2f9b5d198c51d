"""The REAL kernel sources compiled for the host (csrc/emu.h: one OS thread per CUDA thread)
and driven through the REAL C ABI: validates the index logic of the FFT/DCT kernels, the
fused step and the control path without a GPU.  Small sizes only (the harness is slow)."""
import json
import os

import numpy as np
import pytest
import scipy.fftpack as fp

import chsimpy_b200 as ch
from chsimpy_b200 import _lib
from chsimpy_b200.solver import BatchStepper

from emu_lib import EmuBackend

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def be():
    return EmuBackend()


def unit_params(N):
    return _lib.Params(RT=1, BRT=1, B=1, A0=1, A1=1, Amr=1, kappa_tilde=1, L=2, delx=2 / (N - 1), delt=1e-8,
                       delt_max=1e-8, M_tilde=1, threshold=0.5, time_limit_s=0, jitter=0, full_sim=1, adaptive_time=0)


@pytest.mark.parametrize("N", [32, 64, 128, 256, 512])
def test_dctn_idctn(be, N):
    st = BatchStepper(N, [unit_params(N)] * 2, backend=be)
    x = np.random.default_rng(N).random((2, N, N)) - 0.4
    ref = np.stack([fp.dctn(x[i], norm="ortho") for i in range(2)])
    assert np.abs(st.dctn(x) - ref).max() < 1e-13
    assert np.abs(st.dctn(ref, inverse=True) - x).max() < 1e-13


def run_fixture(be, name, limit=None):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    m = json.loads(str(z["meta"]))
    p = ch.Parameters()
    p.no_gui = True
    for k, v in m["params"].items():
        setattr(p, k, v)
    if m["kappa_override"] is not None:
        p.kappa_tilde = m["kappa_override"]
    s = ch.Solver(p, _backend=be)
    s.prepare()
    chunks = m["chunks"] or [p.ntmax if limit is None else limit]
    for c in chunks:
        sol = s.solve_or_resume(c)
    rows, ref = sol.timedata.data(), z["rows"][:len(sol.timedata.data())]
    rel = np.abs(rows - ref) / np.maximum(np.abs(ref), 1e-300)
    rel[ref == 0] = np.abs(rows[ref == 0])
    assert rel.max() < 1e-11, rel.max(axis=0)
    return sol, z, m


@pytest.mark.parametrize("name,steps,env", [("n256_k200", 6, {}), ("n512_full2000", 3, {"CHS_ONE_PER_SM": "0"}),
                                            ("n512_full2000", 3, {"CHS_ONE_PER_SM": "1"}), ("n512_full2000", 3, {"CHS_LL_MAX": "2"})])
def test_step_n256_n512(be, name, steps, env, monkeypatch):
    """The tile geometry of the sizes that matter (N = 256, 512), a few steps against the fixtures: the step kernels
    of the throughput build, the instantiations a launch of at most one tile per SM gets on the GPU (CHS_ONE_PER_SM=1:
    unrolled unit / butterfly loops), and the optional 8-points-per-thread build (chs_ll.cu, CHS_LL_MAX)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    sol, z, m = run_fixture(be, name, limit=steps)
    assert sol.computed_steps == steps


def test_step_n32_full(be):
    sol, z, m = run_fixture(be, "n32_k60")
    assert sol.computed_steps == 60 and np.abs(sol.U - z["U"]).max() < 1e-13


def test_step_n64_prefix(be):
    sol, z, m = run_fixture(be, "n64_k200", limit=25)
    assert sol.computed_steps == 25


def test_chunked_jitter_n128_prefix(be):
    """jitter + re-entry (quirks Q2/Q3) through the emulated kernels: first chunk only."""
    z = np.load(os.path.join(GOLD, "n128_chunked_jitter.npz"))
    m = json.loads(str(z["meta"]))
    p = ch.Parameters()
    p.no_gui = True
    for k, v in m["params"].items():
        setattr(p, k, v)
    s = ch.Solver(p, _backend=be)
    s.prepare()
    sol = s.solve_or_resume(10)
    sol = s.solve_or_resume(5)
    assert sol.computed_steps == 15 and sol.tau0 == 7            # stop test fired at 7, full_sim keeps going
    # not comparable with the 20/20/20 fixture beyond row 10 (different re-entry points, Q2); rows 0..9 are
    rel = np.abs(sol.timedata.data()[:10] - z["rows"][:10]) / np.maximum(np.abs(z["rows"][:10]), 1e-300)
    rel[z["rows"][:10] == 0] = 0
    assert rel.max() < 1e-11


def test_slab_path_emulated(be):
    """Row-slab path (chs_slab.cuh: transposes + row kernels + elementwise update + control
    kernel) on one emulated rank, with a re-entry, against the reference fixture."""
    from chsimpy_b200.slab import SlabEngine
    z = np.load(os.path.join(GOLD, "n64_k200.npz"))
    m = json.loads(str(z["meta"]))
    p = ch.Parameters()
    p.no_gui = True
    for k, v in m["params"].items():
        setattr(p, k, v)
    s = ch.Solver(p, _backend=be, _force_slab=True)
    assert isinstance(s._stepper, SlabEngine)
    s.prepare()
    s.solve_or_resume(8)
    s._stepper._nchunks = 2                                      # row-chunked passes (the pipelined exchange's launches)
    sol = s.solve_or_resume(4)
    assert sol.computed_steps == 12
    rows, ref = sol.timedata.data(), z["rows"][:12]
    rel = np.abs(rows - ref) / np.maximum(np.abs(ref), 1e-300)
    rel[ref == 0] = np.abs(rows[ref == 0])
    assert rel.max() < 1e-11, rel.max(axis=0)
    assert abs(sol.U.mean() - s.U_init.mean()) < 1e-14


def test_slab_peer_route_single_rank(be):
    """The peer-memory route of the slab path (one exchange launch over all peers, sums stored into the ranks'
    gather buffers, gathered control kernel) with the rank as its own peer, against the reference fixture."""
    z = np.load(os.path.join(GOLD, "n64_k200.npz"))
    m = json.loads(str(z["meta"]))
    p = ch.Parameters()
    p.no_gui = True
    for k, v in m["params"].items():
        setattr(p, k, v)
    s = ch.Solver(p, _backend=be, _force_slab=True, _selfpeer=True)
    assert s._stepper._peer is not None
    s.prepare()
    sol = s.solve_or_resume(9)
    rows, ref = sol.timedata.data(), z["rows"][:9]
    rel = np.abs(rows - ref) / np.maximum(np.abs(ref), 1e-300)
    rel[ref == 0] = np.abs(rows[ref == 0])
    assert sol.computed_steps == 9 and rel.max() < 1e-11, rel.max(axis=0)


@pytest.mark.parametrize("kw,steps,selfpeer", [(dict(jitter=0.004), 14, False),
                                               (dict(jitter=0.003, adaptive_time=True, delt_max=3e-9), 508, True)])
def test_slab_jitter_and_adaptive_vs_oracle(be, kw, steps, selfpeer):
    """--jitter / --adaptive-time on the slab path (reference solver.py:177-193, 210-211): the rank's rows of the
    PCG64 noise stream, stencil gradient energy with halo rows, all-rank column sums -> delt; against the oracle,
    with a re-entry in the middle (quirks Q2/Q3; adaptive: the multipliers revert to params.delt on re-entry)."""
    import ch_oracle as orc
    p = ch.Parameters()
    p.N, p.no_gui, p.full_sim, p.kappa_tilde, p.seed, p.ntmax = 64, True, True, 2.7e-4, 5, steps
    for k, v in kw.items():
        setattr(p, k, v)
    s = ch.Solver(p, _backend=be, _force_slab=True, _selfpeer=selfpeer)
    s.prepare()
    first = steps - 5
    s.solve_or_resume(first)
    sol = s.solve_or_resume(5)
    o = orc.run_default(N=64, nsteps=first, seed=5, kappa_tilde=2.7e-4, full_sim=True, **kw)
    o.run(5)
    assert sol.computed_steps == o.computed_steps == steps
    rows, ref = sol.timedata.data(), o.rows
    rel = np.abs(rows - ref) / np.maximum(np.abs(ref), 1e-300)
    rel[ref == 0] = np.abs(rows[ref == 0])
    assert rel.max() < 1e-9, rel.max(axis=0)
    assert np.abs(sol.U - o.U).max() < 1e-11
    assert abs(s.delt - o.delt) <= 1e-12 * o.delt
    if "adaptive_time" in kw:
        assert o.delt > p.delt                   # the adaptive branch really ran


def test_slab_path_honours_stop_flag(be):
    """A time limit hit at step ~10 while the host has 40 steps queued (full_sim, so no poll inside the
    chunk): the slab kernels must freeze the state behind the device-side flag exactly like the batched
    path -- computed_steps, row count, time_passed and U are those of the stopping step (oracle)."""
    import ch_oracle as orc
    kw = dict(N=64, seed=2023, kappa_tilde=3e-4, full_sim=True, time_max=0.3)
    o = orc.run_default(nsteps=40, **kw)
    assert o.stop_reason == "time-limit" and 5 < o.computed_steps < 30
    for force in (False, True):
        p = ch.Parameters()
        p.N, p.no_gui, p.full_sim, p.kappa_tilde, p.ntmax, p.time_max = 64, True, True, 3e-4, 40, 0.3
        s = ch.Solver(p, _backend=be, _force_slab=force)
        s.prepare()
        sol = s.solve_or_resume(40)
        assert sol.stop_reason == "time-limit" and sol.computed_steps == o.computed_steps, (force, sol.computed_steps)
        assert sol.timedata.data().shape == o.rows.shape
        assert abs(s.time_passed - o.time_passed) <= 1e-12 * o.time_passed
        assert sol.tau0 == o.tau0 and sol.t0 == o.t0
        assert np.abs(sol.U - o.U).max() < 1e-13


def test_gemm_path_emulated_vs_oracle(be):
    """DCT-as-GEMM path (chs_gemm.cuh) at a non-power-of-two N, in-kernel time loop with a
    re-entry, against the oracle (the tensor-core MMA is replaced by scalar dot products in
    the host build; indexing and control flow are the same)."""
    import ch_oracle as orc
    assert be.lib.chs_uses_gemm(20, 1) == 1
    p = ch.Parameters()
    p.N, p.no_gui, p.full_sim, p.kappa_tilde, p.ntmax = 20, True, True, 3e-4, 15
    s = ch.Solver(p, _backend=be)
    s.prepare()
    s.solve_or_resume(10)
    sol = s.solve_or_resume(5)
    o = orc.run_default(N=20, nsteps=10, kappa_tilde=3e-4, full_sim=True)
    o.run(5)
    rel = np.abs(sol.timedata.data() - o.rows) / np.maximum(np.abs(o.rows), 1e-300)
    rel[o.rows == 0] = 0
    assert sol.timedata.data().shape == o.rows.shape and rel.max() < 1e-11
    assert np.abs(sol.U - o.U).max() < 1e-13


def test_pcg64_device_stream_emulated(be):
    from chsimpy_b200 import _lib as L
    st = BatchStepper(32, [unit_params(32)], backend=be)
    g = np.random.Generator(np.random.PCG64(7))
    g.random((32, 32))
    noise, mean = st.pcg64_noise(g.bit_generator.state, 2)
    ref = np.stack([g.random((32, 32)) for _ in range(2)])
    assert np.array_equal(noise, ref)


def slot_freqs(N):
    """Frequency held by each slot of an x-spectral row (the `kof` table of chs_api.cu)."""
    M = N // 2
    lg = M.bit_length() - 1
    rad = ([1 << (lg % 3)] if lg % 3 else []) + [8] * (lg // 3)
    kof = np.empty(N, np.int64)
    for k in range(M):
        pos, Lb, kk = 0, M, k
        for r in rad:
            pos += (kk % r) * (Lb // r)
            kk //= r
            Lb //= r
        kof[2 * pos] = k
        kof[2 * pos + 1] = M if k == 0 else N - k
    return kof


@pytest.mark.parametrize("N", [2048, 4096, 8192, 16384])
def test_slab_row_kernels_large_rows(be, N):
    """Row kernels of the slab path at the BASELINE config-5 row lengths (line-major tile with
    bank skew for N >= 4096): forward, inverse and the fused inverse -> physics -> forward pass
    on two tiles of rows, against scipy's DCT and the chemical potential in numpy."""
    import ctypes as C
    lib = be.lib
    R = 2 * lib.chs_slab_row_granularity(N)
    ps = unit_params(N)
    U, A, B, D = (be.empty((R, N)) for _ in range(4))
    rows = be.empty((4, 9))
    wbytes = lib.chs_slab_workspace_bytes(N, R)
    work = be.empty((wbytes,), "u1")
    lam = np.ascontiguousarray(ch.utils.laplace_spectrum_1d(N))
    rb = 8                                                 # global index of the first local row
    h = lib.chs_slab_create(0, N, R, rb, N // R, 0, C.byref(ps), be.ptr(U), be.ptr(rows), 4, be.ptr(work), wbytes,
                            lam.ctypes.data, be.stream_handle())
    assert h
    try:
        kof = slot_freqs(N)
        u = 0.85 + 0.1 * (np.random.default_rng(N).random((R, N)) - 0.5)
        be.upload(U, u)
        assert lib.chs_slab_row(h, 0, be.ptr(U), be.ptr(A), R, rb, 0, 0.0) == 0          # S_FWD
        a = be.download(A)
        ref = fp.dct(u, axis=1, norm="ortho")
        assert np.abs(a - ref[:, kof]).max() < 1e-12
        assert lib.chs_slab_row(h, 2, be.ptr(A), be.ptr(B), R, rb, 0, 0.0) == 0          # S_INV
        assert np.abs(be.download(B) - u).max() < 1e-13
        be.upload(U, np.zeros((R, N)))
        assert lib.chs_slab_row(h, 3, be.ptr(A), be.ptr(D), R, rb, 1, float(u.mean())) == 0   # S_STEP
        assert np.abs(be.download(U) - u).max() < 1e-13                                 # the field it stores
        d = 1 - 2 * u
        mu = np.log(u / (1 - u)) - 1 + (1 + d) * d - 2 * u * (1 - u)                     # unit_params: RT=BRT=A0=A1=1
        assert np.abs(be.download(D) - fp.dct(mu, axis=1, norm="ortho")[:, kof]).max() < 1e-11
        # y pass: natural-order forward transform, then the fused DCT -> spectral update -> IDCT
        H = be.empty((R, N))
        assert lib.chs_slab_row(h, 4, be.ptr(B), be.ptr(H), R, rb, 0, 0.0) == 0          # S_YFWD (B holds u)
        assert np.abs(be.download(H) - ref).max() < 1e-12
        w = np.random.default_rng(N + 1).random((R, N)) - 0.5
        be.upload(B, w)
        slot_base = rb
        assert lib.chs_slab_update(h, be.ptr(H), be.ptr(B), R, slot_base) == 0
        delx2 = (2 / (N - 1)) ** 2
        lam1 = 1e-8 / delx2
        lam2 = lam1 / delx2
        leig = lam[None, :] + lam[kof[slot_base:slot_base + R]][:, None]
        Hn = (ref + lam1 * leig * fp.dct(w, axis=1, norm="ortho")) / (1 + lam2 * leig ** 2)
        assert np.abs(be.download(H) - Hn).max() < 1e-12 * np.abs(Hn).max()
        assert np.abs(be.download(B) - fp.idct(Hn, axis=1, norm="ortho")).max() < 1e-12
    finally:
        lib.chs_slab_destroy(h)


def log_errors(stepper, x):
    """(error in ulp of the result, absolute error) of the kernels' table-driven log against 120-bit mpmath."""
    import mpmath as mpm
    y = stepper.debug_log(x)
    with mpm.workprec(120):
        exact = [mpm.log(mpm.mpf(float(v))) for v in x]
        abs_err = np.array([abs(float(mpm.mpf(float(a)) - b)) for a, b in zip(y, exact)])
        ulp = np.spacing(np.abs(np.array([float(b) for b in exact])))
    return abs_err / ulp, abs_err


def test_fast_log_accuracy(be):
    """csrc/fastlog.cuh stands in for np.log at solver.py:173,220: within 1.25 ulp over the range of a
    concentration field and its complement, within 1e-18 absolutely next to 1, libm outside."""
    st = BatchStepper(32, [unit_params(32)], backend=be)
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(0.005, 0.995, 6000), np.exp(rng.uniform(-40, 40, 2000))])
    ulp_err, abs_err = log_errors(st, x)
    big = np.abs(np.log(x)) >= 2.0 ** -7
    assert ulp_err[big].max() <= 1.25 and abs_err[~big].max() <= 1e-18
    _, abs_err = log_errors(st, 1 + rng.uniform(-2e-3, 2e-3, 2000))
    assert abs_err.max() <= 1e-18
    y = st.debug_log(np.array([0.0, -1.0, np.inf, np.nan, 5e-324]))
    assert y[0] == -np.inf and np.isnan(y[1]) and y[2] == np.inf and np.isnan(y[3]) and abs(y[4] - np.log(5e-324)) < 1e-12


def test_device_lcg_is_the_reference_generator(be):
    """k_lcg_fill against the reference's only golden vector (reference tests/test.py:25-37: 5 x 4, seed 2023)
    and bit for bit against the host restatement at the size of `-g lcg -N 64`."""
    from chsimpy_b200 import mport
    from chsimpy_b200.solver import lcg_sample
    known = [[0.5475444293336684, 0.29257702841077793, 0.3117376865408093, 0.9844947126621821],
             [0.8031704429551821, 0.03775238992541674, 0.37862920778739695, 0.5387215616827465],
             [0.7217314246677474, 0.7984879318617694, 0.8011069301520972, 0.8502945903922872],
             [0.5455620291389348, 0.34767496602035824, 0.8863348965003783, 0.8019890788951838],
             [0.9676096443867356, 0.12967026239711338, 0.008214473728190397, 0.4722352030092083]]
    got = lcg_sample(be, 5, 4, 2023)
    assert np.allclose(got, known, rtol=0, atol=1e-15)
    assert np.array_equal(got, mport.matlab_lcg_sample(5, 4, 2023))
    assert np.array_equal(lcg_sample(be, 64, 64, 2023), mport.matlab_lcg_sample(64, 64, 2023))


@pytest.mark.parametrize("kw,steps", [(dict(), 12), (dict(jitter=0.004), 8)])
def test_arbitrary_n_path_vs_oracle(be, kw, steps):
    """N = 120 (neither a power of two nor <= 104): BigEngine -- the transforms as tiled GEMMs (scalar dot products in
    the host build), np.gradient stencil energy, the slab path's reduction / control kernels -- against the oracle,
    with a re-entry."""
    import ch_oracle as orc
    from chsimpy_b200.slab import BigEngine
    assert be.lib.chs_supports_n(120) == 0 and be.lib.chs_slab_supports_n(120) == 0 and be.lib.chs_big_supports_n(120) == 1
    p = ch.Parameters()
    p.N, p.no_gui, p.full_sim, p.kappa_tilde, p.seed, p.ntmax = 120, True, True, 2.7e-4, 5, steps
    for k, v in kw.items():
        setattr(p, k, v)
    s = ch.Solver(p, _backend=be)
    assert isinstance(s._stepper, BigEngine)
    s.prepare()
    s.solve_or_resume(steps - 4)
    sol = s.solve_or_resume(4)
    o = orc.run_default(N=120, nsteps=steps - 4, seed=5, kappa_tilde=2.7e-4, full_sim=True, **kw)
    o.run(4)
    assert sol.computed_steps == o.computed_steps == steps
    rows, ref = sol.timedata.data(), o.rows
    rel = np.abs(rows - ref) / np.maximum(np.abs(ref), 1e-300)
    rel[ref == 0] = np.abs(rows[ref == 0])
    assert rel.max() < 1e-9, rel.max(axis=0)
    assert np.abs(sol.U - o.U).max() < 1e-11


def test_mixed_launches_are_bit_identical(be):
    """chs_steps as mixed launches (k_mix: column half-step of one half of the members + row half-step of the
    other half per launch, the halves skewed by half a step) against one column + one row launch per iteration:
    the same kernels on the same data in the same per-member order, so TimeData rows, fields and states must be
    identical bit for bit -- with an odd member count, two calls in a row, the `last` flag and a member that
    reaches its time limit in the middle of a call (device stop flag; its half keeps stepping)."""
    N, B = 64, 5
    structs = []
    for i in range(B):
        p = unit_params(N)
        p.delt = 1e-8 * (1 + 0.1 * i)
        p.A0 = 1 + 0.05 * i
        p.full_sim = 1
        if i == 3:
            p.time_limit_s = 4.5e-8 * 1.3           # halts after its 4th accounting step
        structs.append(p)
    U0 = 0.5 + 0.2 * (np.random.default_rng(7).random((N, N)) - 0.5)
    out = {}
    for mode in (0, 1):
        st = BatchStepper(N, structs, rows_cap=32, backend=be)
        st.set_mix(mode)
        st.set_U(U0)
        st.prepare()
        st.begin()
        l0 = st.launch_count()
        st.steps(4)
        st.steps(5, last=True)
        n_launch = st.launch_count() - l0
        running, stop, cs, rw = st.poll()
        rows = st.take_rows()
        st.end()
        out[mode] = (n_launch, stop.copy(), cs.copy(), rw.copy(), rows, np.stack([st.get_U(i) for i in range(B)]))
    assert out[0][0] == 2 * 9 and out[1][0] == (2 * 4 + 1) + (2 * 5 + 1)
    assert (out[0][1] == out[1][1]).all() and (out[0][2] == out[1][2]).all() and (out[0][3] == out[1][3]).all()
    assert out[0][2][3] < 10 and out[0][2][0] == 10, out[0][2]        # (prepare counts as step 1) member 3 stopped early, the others ran on
    for a, b in zip(out[0][4], out[1][4]):
        assert a.shape == b.shape and (a.view(np.int64) == b.view(np.int64)).all()
    assert (out[0][5].view(np.int64) == out[1][5].view(np.int64)).all()
