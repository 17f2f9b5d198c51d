"""Solver: drop-in for reference chsimpy/solver.py (`Solver(params, U_init)`, `prepare()`,
`solve_or_resume(nsteps)`), with the time loop running as hand-written sm_100a CUDA
kernels behind the C ABI of include/chs_b200.h.  PyTorch only owns the device buffers.

`BatchStepper` is the same machinery for a batch of independent simulations (the A0/A1
ensemble of reference chsimpy/experiment.py); `Solver` is its batch-of-one face.

There is no CPU path here: without a CUDA device (or with the library missing and
unbuildable) construction raises."""
import ctypes as C

import numpy as np
from scipy.stats import qmc

from . import _lib, mport, utils
from .solution import Solution
from .timedata import TimeData


class _CudaBackend:
    """Device memory = torch tensors; compute = libchs_b200.so on torch's current stream."""
    name = "cuda"

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("chsimpy_b200 needs a CUDA device: the B200 stepper has no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.lib = _lib.load()

    def empty(self, shape, dtype="f8"):
        dt = {"f8": self.torch.float64, "u1": self.torch.uint8}[dtype]
        return self.torch.empty(shape, dtype=dt, device=self.device)

    def ptr(self, t):
        return t.data_ptr()

    def upload(self, t, arr):
        t.copy_(self.torch.from_numpy(np.require(arr, dtype=np.float64, requirements=['C', 'W'])))

    def upload_broadcast(self, t, arr2d):
        """t[b] = arr2d for every b: ONE host->device copy of the shared field, replicated on the device."""
        one = self.torch.from_numpy(np.require(arr2d, dtype=np.float64, requirements=['C', 'W'])).to(self.device, non_blocking=True)
        t.copy_(one.expand_as(t))

    def to_device(self, arr):
        return self.torch.from_numpy(np.require(arr, requirements=['C', 'W'])).to(self.device)

    def download(self, t):
        return t.detach().cpu().numpy()

    def stream_handle(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def device_index(self):
        return self.device.index


def lcg_sample(be, n1, n2, seed):
    """mport.matlab_lcg_sample(n1, n2, seed) computed by the device kernel k_lcg_fill (one thread: the
    float64 recurrence is serial; ~10 ms for 512 x 512 instead of a 262 144-iteration Python loop)."""
    out = be.empty((n1, n2))
    _lib.check(be.lib, be.lib.chs_lcg_fill(be.ptr(out), int(n1), int(n2), float(seed), be.stream_handle()), "chs_lcg_fill")
    return be.download(out)


def make_params_struct(params, sol):
    """chs_params from a Parameters/Solution pair (reference solution.py:25-50)."""
    jitter = params.jitter if (params.jitter is not None and 0.0 < params.jitter < 0.1) else 0.0   # solver.py:210
    limit = params.time_max * 60 if (params.time_max is not None and params.time_max > 0) else 0.0  # solver.py:148-150
    return _lib.Params(RT=sol.RT, BRT=sol.BRT, B=params.B, A0=float(sol.A0), A1=float(sol.A1), Amr=sol.Amr,
                       kappa_tilde=float(sol.kappa_tilde), L=float(params.L), delx=sol.delx,
                       delt=params.delt, delt_max=params.delt_max, M_tilde=params.M_tilde,
                       threshold=params.threshold, time_limit_s=float(limit), jitter=float(jitter),
                       full_sim=int(bool(params.full_sim)), adaptive_time=int(bool(params.adaptive_time)))


class BatchStepper:
    """`batch` independent N x N simulations stepping in lock-step on one GPU."""

    def __init__(self, N, param_structs, rows_cap=2048, backend=None, device=None):
        self.be = backend if backend is not None else _CudaBackend(device)
        lib = self.lib = self.be.lib
        self.N, self.batch, self.rows_cap = int(N), len(param_structs), int(rows_cap)
        if not lib.chs_supports_n(self.N):
            raise ValueError(f"N={N} is not supported by the batched kernels (FFT path: powers of two 32..1024; "
                             f"tensor-core GEMM path: any N from 8 to 104)")
        b, n = self.batch, self.N
        self.U = self.be.empty((b, n, n))
        self.hatU = self.be.empty((b, n, n))
        self.T = self.be.empty((b, n, n))
        self.rows = self.be.empty((b, self.rows_cap, 9))
        wbytes = lib.chs_workspace_bytes(n, b)
        self.work = self.be.empty((wbytes,), "u1")
        lam = np.ascontiguousarray(utils.laplace_spectrum_1d(n), dtype=np.float64)
        self._h = _lib.check(lib, lib.chs_create(self.be.device_index(), n, b, self.be.ptr(self.U),
                                                  self.be.ptr(self.hatU), self.be.ptr(self.T),
                                                  self.be.ptr(self.rows), self.rows_cap, self.be.ptr(self.work),
                                                  wbytes, lam.ctypes.data, self.be.stream_handle()), "chs_create")
        for i, ps in enumerate(param_structs):
            _lib.check(lib, lib.chs_set_params(self._h, i, C.byref(ps)), "chs_set_params")
        self._stop = np.zeros(b, np.int32)
        self._cs = np.zeros(b, np.int64)
        self._rw = np.zeros(b, np.int64)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.chs_destroy(h)

    # -- fields ---------------------------------------------------------------------------
    def set_U(self, U):
        U = np.asarray(U, dtype=np.float64)
        if U.ndim == 2:                                   # one field shared by all members (the ensemble, quirk Q11)
            assert U.shape == (self.N, self.N)
            if hasattr(self.be, "upload_broadcast"):
                self.be.upload_broadcast(self.U, U)
            else:
                self.be.upload(self.U, np.broadcast_to(U, (self.batch,) + U.shape))
            self._meanU = np.full(self.batch, U.mean())
            return
        assert U.shape == (self.batch, self.N, self.N)
        self.be.upload(self.U, U)
        self._meanU = np.ascontiguousarray(U.reshape(self.batch, -1).mean(axis=1)) if U.strides[0] else \
            np.full(self.batch, U[0].mean())

    def get_U(self, sim=None):
        u = self.be.download(self.U if sim is None else self.U[sim])
        return u

    # -- control --------------------------------------------------------------------------
    def prepare(self):
        """chs_prepare; returns row 0 of every sim, shape (batch, 9)."""
        m = np.ascontiguousarray(self._meanU, dtype=np.float64)
        _lib.check(self.lib, self.lib.chs_prepare(self._h, m.ctypes.data), "chs_prepare")
        self.poll()
        return self.be.download(self.rows[:, 0, :])

    def begin(self):
        _lib.check(self.lib, self.lib.chs_begin(self._h), "chs_begin")

    def steps(self, n, noise=None, noise_mean=None, last=False):
        pn = self.be.ptr(noise) if noise is not None else None
        pm = self.be.ptr(noise_mean) if noise_mean is not None else None
        _lib.check(self.lib, self.lib.chs_steps(self._h, int(n), pn, pm, int(bool(last))), "chs_steps")

    def poll(self):
        running = _lib.check(self.lib, self.lib.chs_poll(self._h, self._stop.ctypes.data, self._cs.ctypes.data,
                                                         self._rw.ctypes.data), "chs_poll")
        return running, self._stop, self._cs, self._rw

    def take_rows(self):
        """Rows written since the last rewind, per sim; rewinds the device cursor."""
        top = int(self._rw.max())
        out = [np.empty((0, 9))] * self.batch
        if top > 0:
            block = self.be.download(self.rows[:, :min(top, self.rows_cap), :])
            out = [block[i, :int(self._rw[i])].copy() for i in range(self.batch)]
            _lib.check(self.lib, self.lib.chs_rewind_rows(self._h), "chs_rewind_rows")
        return out

    def end(self):
        _lib.check(self.lib, self.lib.chs_end(self._h), "chs_end")

    def get_state(self, sim=0):
        st = _lib.State()
        _lib.check(self.lib, self.lib.chs_get_state(self._h, sim, C.byref(st)), "chs_get_state")
        return st

    def set_state(self, sim, st):
        _lib.check(self.lib, self.lib.chs_set_state(self._h, sim, C.byref(st)), "chs_set_state")

    def set_timing(self, on):
        _lib.check(self.lib, self.lib.chs_set_timing(self._h, int(bool(on))), "chs_set_timing")

    def get_timing(self):
        """({'col': ms, 'row': ms, 'diag': ms}, iterations) accumulated since the last call."""
        ms = (C.c_double * 3)()
        n = C.c_int64(0)
        _lib.check(self.lib, self.lib.chs_get_timing(self._h, ms, C.byref(n)), "chs_get_timing")
        return {"col": ms[0], "row": ms[1], "diag": ms[2]}, int(n.value)

    def set_mix(self, mode):
        """-1: mixed launches (k_mix) when enough members run (default), 0: never, 1: whenever >= 2 members run."""
        _lib.check(self.lib, self.lib.chs_set_mix(self._h, int(mode)), "chs_set_mix")

    def get_timing_mix(self):
        """({'mix': ms, 'solo': ms}, {'mix': launches, 'solo': launches}, iterations) of the calls that ran as
        mixed launches (k_mix) since the last call."""
        ms = (C.c_double * 2)()
        n2 = (C.c_int64 * 2)()
        n = C.c_int64(0)
        _lib.check(self.lib, self.lib.chs_get_timing_mix(self._h, ms, n2, C.byref(n)), "chs_get_timing_mix")
        return {"mix": ms[0], "solo": ms[1]}, {"mix": int(n2[0]), "solo": int(n2[1])}, int(n.value)

    def pcg64_noise(self, bit_generator_state, n):
        """(noise [n][N][N], per-step means [n]) on the device: the next n*N*N doubles of a numpy
        PCG64 generator whose `bit_generator.state` is given -- bit-identical to rng.random()."""
        st = bit_generator_state["state"]
        s128, i128 = int(st["state"]), int(st["inc"])
        m64 = (1 << 64) - 1
        count = n * self.N * self.N
        noise = self.be.empty((n, self.N, self.N))
        mean = self.be.empty((n,))
        _lib.check(self.lib, self.lib.chs_pcg64_fill(self._h, s128 >> 64, s128 & m64, i128 >> 64, i128 & m64, 0,
                                                     self.be.ptr(noise), count), "chs_pcg64_fill")
        _lib.check(self.lib, self.lib.chs_row_means(self._h, self.be.ptr(noise), n, self.N * self.N, self.be.ptr(mean)),
                   "chs_row_means")
        return noise, mean

    def debug_log(self, x):
        """Device fast_log of a host array (self-test of csrc/fastlog.cuh)."""
        x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        src, dst = self.be.to_device(x), self.be.empty((x.size,))
        _lib.check(self.lib, self.lib.chs_debug_log(self._h, self.be.ptr(src), self.be.ptr(dst), x.size), "chs_debug_log")
        return self.be.download(dst)

    def launch_count(self):
        return int(self.lib.chs_launch_count(self._h))

    def dctn(self, x, inverse=False):
        """2-D orthonormal DCT-II / DCT-III of a (batch, N, N) host array (test entry point)."""
        src = self.be.to_device(np.asarray(x, dtype=np.float64).reshape(self.batch, self.N, self.N))
        dst = self.be.empty((self.batch, self.N, self.N))
        fn = self.lib.chs_idctn if inverse else self.lib.chs_dctn
        _lib.check(self.lib, fn(self._h, self.be.ptr(src), self.be.ptr(dst)), "chs_dctn")
        return self.be.download(dst)

    # -- the loop -------------------------------------------------------------------------
    def run(self, iters, draw_noise=None, poll_every=None):
        """`iters` iterations of the reference loop body for every sim (begin .. end).
        draw_noise(n) -> (n, N, N) uniform draws for the next n iterations, or None.
        Returns (rows_per_sim, iterations_completed_per_sim)."""
        rows = [[] for _ in range(self.batch)]
        done_iters = np.zeros(self.batch, np.int64)
        if iters <= 0:
            return [np.empty((0, 9))] * self.batch, done_iters
        chunk = self.rows_cap if poll_every is None else min(self.rows_cap, int(poll_every))
        self.begin()
        done, running = 0, self.batch
        while done < iters and running > 0:
            n = min(chunk, iters - done)
            noise = nmean = None
            if draw_noise is not None:
                drawn = draw_noise(n)
                if isinstance(drawn, tuple):              # already on the device (PCG64 kernel)
                    noise, nmean = drawn
                else:
                    noise = self.be.to_device(drawn)
                    nmean = self.be.to_device(drawn.reshape(n, -1).mean(axis=1))
            self.steps(n, noise, nmean, last=(done + n == iters))
            running, _, _, _ = self.poll()
            got = self.take_rows()
            for i, r in enumerate(got):
                if len(r):
                    rows[i].append(r)
                    done_iters[i] += len(r)
            done += n
        self.end()
        return [np.concatenate(r) if r else np.empty((0, 9)) for r in rows], done_iters


class Solver:
    """Cahn-Hilliard integrator (DCT, Flory-Huggins energy) -- API of reference
    chsimpy/solver.py:45-252."""

    def __init__(self, params=None, U_init=None, _backend=None, _world=None, _force_slab=False, _selfpeer=False):
        self.params = params
        self.solution = Solution(self.params)
        N = params.N
        self.skip_check = False
        self.time_delta_sum = 0.0
        self.time_passed = 0.0
        self._prepared = False
        self.delt = self.params.delt
        self.create_rand = None
        self.U_init = None
        self._rng = None
        self._sobol = None
        self._sobol_drawn = 0
        # initial concentration field, reference solver.py:56-82
        if U_init is not None:
            if U_init.shape == (params.N, params.N):
                self.U_init = U_init
            else:
                print("U_init has wrong shape, must match parameters.N")
                exit(1)
        elif params.generator == 'lcg':
            lcg = lcg_sample(_backend if _backend is not None else _CudaBackend(), N, N, params.seed)
            self.U_init = params.XXX + (params.XXX * 0.01 * lcg)
        elif params.generator == 'sobol':
            self._sobol = qmc.Sobol(d=N, seed=params.seed)
            self.create_rand = self._draw_sobol
        elif params.generator == 'simplex':
            import opensimplex   # optional dependency, as in the reference
            self.create_rand = lambda n: opensimplex.noise2array(np.linspace(0, 48, n), np.linspace(0, 48, n))
        else:
            self._rng = np.random.Generator(np.random.PCG64(params.seed))
            self.create_rand = lambda n: self._rng.random((n, n))
        if self.U_init is None:
            self.U_init = params.XXX + (params.XXX * 0.01 * (self.create_rand(N) - 0.5))
        # engine: the batched tile kernels for N <= 1024 on one GPU, the row-slab path for a
        # larger domain or one that is decomposed over several ranks (_world = (rank, size))
        ps = make_params_struct(params, self.solution)
        be = _backend if _backend is not None else _CudaBackend()
        if _world is None and not _force_slab and be.lib.chs_supports_n(N):
            self._stepper = BatchStepper(N, [ps], backend=be)
        elif be.lib.chs_slab_supports_n(N):
            from .slab import SlabEngine
            self._stepper = SlabEngine(N, ps, backend=be, world=_world, _selfpeer=_selfpeer)
        elif _world is None and be.lib.chs_big_supports_n(N):
            from .slab import BigEngine              # any other N up to 2048: the transforms as tensor-core GEMMs
            self._stepper = BigEngine(N, ps, backend=be)
        else:
            raise ValueError(f"N={N}: supported are powers of two 32..16384 (FFT kernels; 64..16384 over several GPUs) "
                             f"and any N from 8 to 2048 on one GPU (GEMM kernels)")

    def _draw_sobol(self, n):
        self._sobol_drawn += n
        return self._sobol.random(n)

    # ---------------------------------------------------------------------------------
    def prepare(self):
        """Row 0 of the diagnostics and state reset (reference solver.py:84-135)."""
        st = self._stepper
        U = np.array(self.U_init, dtype=np.float64, copy=True)
        assert U.shape == (self.params.N, self.params.N)
        self._push_host_state(reset=False)
        st.set_U(U)
        row0 = st.prepare()[0]
        data = TimeData()
        data.extend(row0[None, :])
        self.solution.U = U
        self.solution.timedata = data
        self.solution.tau0 = 0.0
        self.solution.t0 = 0.0
        self.solution.stop_reason = 'None'
        self.solution.computed_steps = 1
        self._prepared = True

    def _push_host_state(self, reset):
        """The attributes a caller may have touched between calls (delt, time_delta_sum,
        skip_check persist across prepare(), quirk Q16) are the device's initial state."""
        s = self._stepper.get_state(0)
        s.delt = float(self.delt)
        s.time_delta_sum = float(self.time_delta_sum)
        s.time_passed = float(self.time_passed)
        s.skip_check = int(bool(self.skip_check))
        self._stepper.set_state(0, s)

    def _pull_device_state(self):
        s = self._stepper.get_state(0)
        self.delt = s.delt
        self.time_delta_sum = s.time_delta_sum
        self.time_passed = s.time_passed
        self.skip_check = bool(s.skip_check)
        sol = self.solution
        sol.computed_steps = int(s.computed_steps)
        sol.tau0 = int(s.tau0) if s.tau0 != 0 else 0.0        # a count once set (solver.py:243), 0.0 after prepare
        sol.t0 = float(s.t0)
        if s.stop_reason in (0, 1, 2):
            sol.stop_reason = _lib.STOP_NAMES[s.stop_reason]
        return s

    def solve_or_resume(self, nsteps=None):
        """Runs the time loop (reference solver.py:137-252) and returns the Solution."""
        assert (self._prepared is True)
        p, sol, st = self.params, self.solution, self._stepper
        if nsteps is None:
            nsteps = max(p.ntmax, 0)
        first = 1 if sol.computed_steps == 1 else 0            # prepare() did the first row (Q3)
        iters = int(nsteps) - first
        jitter_on = p.jitter is not None and 0.0 < p.jitter < 0.1
        if jitter_on and self.create_rand is None and iters > 0:
            raise TypeError("'NoneType' object is not callable")     # reference quirk Q7
        draw = None
        snap = None
        if jitter_on:
            snap = self._rng.bit_generator.state if self._rng is not None else None
            sob0 = self._sobol_drawn
            N = p.N

            def draw(n):
                if self._rng is not None and hasattr(st, "pcg64_noise"):
                    # numpy's PCG64 stream reproduced on the device; the host generator is only advanced
                    dev = st.pcg64_noise(self._rng.bit_generator.state, n)
                    self._rng.bit_generator.advance(n * N * N)
                    return dev
                out = np.empty((n, N, N))
                for i in range(n):
                    out[i] = self.create_rand(N)
                return out
        poll = 64 if jitter_on else (128 if not p.full_sim or p.time_max else None)
        rows, done = st.run(iters, draw_noise=draw, poll_every=poll)
        if jitter_on and iters > 0:
            self._rewind_noise(snap, sob0, int(done[0]))
        self._pull_device_state()
        if done[0] > 0 or iters > 0:
            sol.U = st.get_U(0)
        sol.timedata.extend(rows[0])           # raises AssertionError on a NaN row (timedata.py:10)
        return sol

    def _rewind_noise(self, snap, sob0, consumed):
        """Noise is drawn a chunk ahead; give back what a stopped run did not consume so the
        generator continues exactly where the reference's would."""
        N = self.params.N
        if self._rng is not None and snap is not None:
            self._rng.bit_generator.state = snap
            self._rng.bit_generator.advance(consumed * N * N)
        elif self._sobol is not None:
            self._sobol.reset()
            self._sobol.fast_forward(sob0 + consumed * N)
            self._sobol_drawn = sob0 + consumed * N
