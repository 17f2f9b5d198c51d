"""Large single domain: one N x N simulation row-slab decomposed over the GPUs of a box
(BASELINE config 5), or run on one GPU when it does not fit the batched tile path (N > 1024).

Every transform pass of the stepper works on contiguous rows; the change of direction is a
transpose.  With P ranks each rank owns R = N/P rows of U and R x-spectral rows of hat_U and
one CH step is (chs_slab.cuh):

    B = rowIDCT(H = (H + Seig*rowDCT(B))/CHeig)      one kernel (B = transpose(A) of the last step)
    A = transpose(B)          P > 1: tiled transposes that write straight into the peer ranks'
                              buffers over NVLink (symmetric memory) + a device-side barrier
                              (optionally per row chunk on a second stream, CHS_SLAB_CHUNKS > 1:
                              measured no gain, the row kernels fill the GPU); NCCL all-to-all
                              with pack/unpack when peer mapping is unavailable
    U, A = rowIDCT(A) -> physics, diagnostics -> rowDCT(mu)
    B = transpose(A)          same exchange, for the next step
    7 diagnostic sums: all-reduce (NCCL) -> device-side control kernel (TimeData row, stop test)

No host synchronisation happens inside a chunk of steps; the host polls the stop flag every
`poll_every` steps exactly like the batched path.  `SlabEngine` has the same interface as
`BatchStepper` (batch of one), so `Solver` drives either.
"""
import ctypes as C
import os

import numpy as np

from . import _lib, utils


class SlabEngine:
    S_FWD, S_MU, S_INV, S_STEP, S_YFWD = 0, 1, 2, 3, 4

    def __init__(self, N, param_struct, rows_cap=1024, backend=None, device=None, world=None, _selfpeer=False):
        from .solver import _CudaBackend
        self.be = backend if backend is not None else _CudaBackend(device)
        lib = self.lib = self.be.lib
        self.N = int(N)
        self.rank, self.P = (0, 1) if world is None else (int(world[0]), int(world[1]))
        if not lib.chs_slab_supports_n(self.N):
            raise ValueError(f"N={N} is not supported by the slab path (powers of two, 64..16384)")
        gran = lib.chs_slab_row_granularity(self.N)
        if self.N % self.P or (self.N // self.P) % gran or self.N // self.P < 2:
            raise ValueError(f"N={N} cannot be split into {self.P} slabs of a multiple of {gran} rows")
        self.R = R = self.N // self.P
        self.row_base = self.rank * R
        self.batch, self.rows_cap = 1, int(rows_cap)
        n = self.N
        self.U = self.be.empty((R, n))
        self.H = self.be.empty((R, n))
        self.Uh = None                                   # [R+2][N] halo copy, prepare() only
        self.rows = self.be.empty((self.rows_cap, 9))
        self._peer = None                                # peer-mapped base pointers of the A|B exchange buffer
        self._nchunks = 1
        self._ce = False                                 # copy-engine exchange (CHS_SLAB_CE, see _pass_ce)
        self._main = self._side = None
        if self.P > 1 and self.be.name == "cuda" and os.environ.get("CHS_SLAB_P2P", "1") != "0":
            self._setup_peer_buffers()
        elif self.P == 1 and _selfpeer:
            # test hook: the peer-memory route with this rank as its own (only) peer -- exercises the one-launch
            # exchange, the gathered sums and the gathered control kernel without a second GPU
            RN = R * n
            flat = self.be.empty((2 * RN + 16,))
            self._ab, self._hdl = flat, type("NoBarrier", (), {"barrier": staticmethod(lambda: None)})()
            self.A, self.B = flat[:RN].reshape(R, n), flat[RN:2 * RN].reshape(R, n)
            self._gather = flat[2 * RN:].reshape(2, 1, 8)
            self._peer, self._ab_base, self._parity = [self.be.ptr(flat)], self.be.ptr(flat), 0
        if self._peer is None:
            self.A = self.be.empty((R, n))
            self.B = self.be.empty((R, n))
            if self.P > 1:
                self.send = self.be.empty((self.P, R, R))
                self.recv = self.be.empty((self.P, R, R))
        wbytes = lib.chs_slab_workspace_bytes(n, R)
        self.work = self.be.empty((wbytes,), "u1")
        lam = np.ascontiguousarray(utils.laplace_spectrum_1d(n), dtype=np.float64)
        self._ps = param_struct
        self._h = _lib.check(lib, lib.chs_slab_create(self.be.device_index(), n, R, self.row_base, self.P, self.rank,
                                                       C.byref(param_struct), self.be.ptr(self.U), self.be.ptr(self.rows),
                                                       self.rows_cap, self.be.ptr(self.work), wbytes, lam.ctypes.data,
                                                       self.be.stream_handle()), "chs_slab_create")
        vec_ptr = lib.chs_slab_vec(self._h)
        off = vec_ptr - self.be.ptr(self.work)
        # the 7 rank-local sums as a tensor torch.distributed can all-reduce in place (NCCL on the
        # device; the host-emulated test backend hands out numpy memory -> gloo)
        self._vec = self._tensor(self.work[off:off + 56]).view(self._torch().float64)
        self._full = None
        self._mean = 0.0
        self._prof = None                                # bench.py: list of (start, end) CUDA events around the exchanges
        self._cols = self._cols_scr = self._halo = None  # adaptive dt: all-rank column sums; jitter: neighbours' rows
        self._cs = 1                                     # host mirror of computed_steps (which iterations update delt)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.chs_slab_destroy(h)

    def _setup_peer_buffers(self):
        """A and B live in one symmetric allocation that every rank of the box maps (NVLink peer
        memory, torch.distributed._symmetric_memory): the transposes then WRITE their blocks
        straight into the destination rank's buffer -- no pack buffer, no all-to-all, no unpack."""
        try:
            import torch
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            # one symmetric allocation: A | B | gather buffer of the diagnostic sums [2 (step parity)][P][8]
            RN = self.R * self.N
            flat = symm.empty((2 * RN + 16 * self.P,), dtype=torch.float64, device=self.U.device)
            hdl = symm.rendezvous(flat, dist.group.WORLD)
            peers = [int(x) for x in hdl.buffer_ptrs]
            if len(peers) != self.P or int(hdl.rank) != self.rank:
                raise RuntimeError("unexpected symmetric-memory group")
            off = int(getattr(hdl, "offset", 0) or 0)
            self._ab, self._hdl = flat, hdl
            self.A, self.B = flat[:RN].view(self.R, self.N), flat[RN:2 * RN].view(self.R, self.N)
            self._gather = flat[2 * RN:].view(2, self.P, 8)
            self._gather.zero_()
            self._peer = [x + off for x in peers]
            self._ab_base = self.be.ptr(flat)
            self._parity = 0
            hdl.barrier()
            n = int(os.environ.get("CHS_SLAB_CHUNKS", "1"))   # measured on 2 GPUs, N=8192: 1.59 / 1.67 / 1.66 ms for 1 / 2 / 4
            ce = os.environ.get("CHS_SLAB_CE", "0") != "0"
            gran = self.lib.chs_slab_row_granularity(self.N)
            if (n > 1 or ce) and self.R % (n * gran) == 0 and self.R // n >= 128 and (self.R // n) % 128 == 0:
                self._nchunks = n
                self._main = torch.cuda.current_stream()
                self._side = torch.cuda.Stream()
                self._ev = [torch.cuda.Event() for _ in range(n)]
                self._ev_done = torch.cuda.Event()
                if ce:
                    # copy-engine exchange: two staging buffers of one row chunk each (plain device memory)
                    self._ce = True
                    self._stage = [torch.empty((self.R // n) * self.N, dtype=torch.float64, device=self.U.device) for _ in range(2)]
                    self._copied = [torch.cuda.Event() for _ in range(n)]
        except Exception as e:                           # noqa: BLE001 -- any failure: NCCL all-to-all path
            if self.rank == 0:
                print(f"[chsimpy_b200.slab] peer-memory transposes unavailable ({type(e).__name__}: {e}); "
                      f"using NCCL all-to-all", flush=True)
            self._peer = None
            self._nchunks = 1

    # -- helpers ----------------------------------------------------------------------------
    def _torch(self):
        import torch
        return torch

    def _tensor(self, x):
        """Backend array -> torch tensor sharing its memory."""
        return x if self.be.name == "cuda" else self._torch().from_numpy(x)

    def _ck(self, rc, what):
        return _lib.check(self.lib, rc, what)

    def _row(self, mode, src, dst, diag=0):
        self._ck(self.lib.chs_slab_row(self._h, mode, self.be.ptr(src), self.be.ptr(dst), self.R, self.row_base,
                                       int(diag), float(self._mean)), "chs_slab_row")

    def _allreduce_vec(self):
        if self.P > 1:
            import torch.distributed as dist
            dist.all_reduce(self._vec)                  # 7 doubles, NCCL; identical on every rank

    def _transpose(self, src, dst, r0=0, rc=None, sync=True):
        if self._prof is None:
            return self._transpose_impl(src, dst, r0, rc, sync)
        torch = self._torch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self._transpose_impl(src, dst, r0, rc, sync)
        e1.record()
        self._prof.append((e0, e1))

    def _transpose_impl(self, src, dst, r0=0, rc=None, sync=True):
        """dst[c_local][r_global] = src[r_local][c_global] for the local rows r0 .. r0+rc (works in
        both directions).  P > 1: the block for rank p is written into p's buffer; with `sync` a
        cross-rank barrier follows (all writes have landed, all reads of the old dst are over)."""
        lib, h, be, R, N, P = self.lib, self._h, self.be, self.R, self.N, self.P
        rc = R if rc is None else rc
        esz = 8
        if P == 1 and self._peer is None:
            self._ck(lib.chs_slab_transpose(h, be.ptr(src) + r0 * N * esz, be.ptr(dst) + r0 * esz, rc, N, N, N),
                     "chs_slab_transpose")
            return
        if self._peer is not None:
            # out_p[c_local][rank*R + r_local] = src[r_local][p*R + c_local], over NVLink into rank p's dst:
            # ONE launch, blockIdx.z = peer (the local block first, the link load spread)
            doff = be.ptr(dst) - self._ab_base
            dsts = (C.c_uint64 * P)(*[self._peer[p] + doff + (self.rank * R + r0) * esz for p in range(P)])
            self._ck(lib.chs_slab_transpose_peers(h, be.ptr(src) + r0 * N * esz, dsts, rc, R, N, N), "chs_slab_transpose_peers")
            if sync:
                self._hdl.barrier()
            return
        import torch.distributed as dist
        assert r0 == 0 and rc == R
        for p in range(P):                               # block p of my rows, transposed, goes to rank p
            self._ck(lib.chs_slab_transpose(h, be.ptr(src) + p * R * esz, be.ptr(self.send) + p * R * R * esz,
                                            R, R, N, R), "chs_slab_transpose")
        dist.all_to_all_single(self._tensor(self.recv), self._tensor(self.send))     # NCCL over NVLink / NVSwitch
        # recv[q] = [c_local][r_local of rank q]  ->  dst[c_local][q*R + r_local]
        self._tensor(dst).view(R, P, R).copy_(self._tensor(self.recv).permute(1, 0, 2))

    def _chunks(self):
        """Row chunks of the pipelined exchange (peer-memory path): the transposes of chunk k run
        on a second stream while chunk k+1 is transformed."""
        n = self._nchunks
        rc = self.R // n
        return [(k * rc, rc) for k in range(n)]

    def _pass_ce(self, compute, src, dst, barrier=True):
        """One direction of a step with the COPY ENGINES driving NVLink: per row chunk compute(r0, rc), then one
        launch transposes the chunk's P blocks into a local staging buffer (this rank's own block straight into
        `dst`), then one pitched device-to-device copy per peer on the side stream moves block p into rank p's
        `dst` -- while the SMs already transform the next chunk.  Ends with the exchange barrier."""
        lib, h, be, R, N, P = self.lib, self._h, self.be, self.R, self.N, self.P
        main, side = self._main, self._side
        doff = be.ptr(dst) - self._ab_base
        chunks = self._chunks()
        for k, (r0, rc) in enumerate(chunks):
            compute(r0, rc)
            if k >= 2:
                main.wait_event(self._copied[k - 2])     # the copies out of this staging buffer are over
            stage = self._stage[k % 2]
            col0 = (self.rank * R + r0) * 8              # this rank's columns of the destination rows
            self._ck(lib.chs_slab_transpose_stage(h, be.ptr(src) + r0 * N * 8, be.ptr(stage), be.ptr(dst) + col0, N, rc, R, N),
                     "chs_slab_transpose_stage")
            self._ev[k].record(main)
            side.wait_event(self._ev[k])
            dsts = (C.c_uint64 * P)(*[self._peer[p] + doff + col0 for p in range(P)])
            self._ck(lib.chs_slab_copy_blocks(h, dsts, N * 8, be.ptr(stage), rc, R, side.cuda_stream), "chs_slab_copy_blocks")
            self._copied[k].record(side)
        main.wait_event(self._copied[len(chunks) - 1])   # the side stream is in order: the last copy implies all
        if barrier:
            self._hdl.barrier()

    def _pass(self, compute, src, dst):
        """One direction of a step: compute(r0, rc) on every row chunk of `src`, each followed by its
        transposes into `dst` (on the side stream when pipelined), then the exchange barrier."""
        if self._ce:
            return self._pass_ce(compute, src, dst)
        if self._nchunks == 1:
            compute(0, self.R)
            self._transpose(src, dst)
            return
        main, side = self._main, self._side
        for k, (r0, rc) in enumerate(self._chunks()):
            compute(r0, rc)
            if side is None:                             # one stream (single rank): chunked, not overlapped
                self._transpose(src, dst, r0, rc, sync=False)
                continue
            self._ev[k].record(main)
            side.wait_event(self._ev[k])
            self._ck(self.lib.chs_slab_set_stream(self._h, side.cuda_stream), "chs_slab_set_stream")
            self._transpose(src, dst, r0, rc, sync=False)
            self._ck(self.lib.chs_slab_set_stream(self._h, main.cuda_stream), "chs_slab_set_stream")
        if side is not None:
            self._ev_done.record(side)
            main.wait_event(self._ev_done)
            self._hdl.barrier()

    # -- BatchStepper interface -------------------------------------------------------------
    def set_U(self, U):
        U = np.asarray(U, dtype=np.float64)
        if U.ndim == 3:
            U = U[0]
        assert U.shape == (self.N, self.N)
        self._full = U
        self._mean = float(U.mean())
        r0, R = self.row_base, self.R
        self.be.upload(self.U, U[r0:r0 + R])

    def prepare(self):
        U, r0, R, N = self._full, self.row_base, self.R, self.N
        halo = np.empty((R + 2, N))
        halo[1:R + 1] = U[r0:r0 + R]
        halo[0] = U[r0 - 1] if r0 > 0 else U[r0]
        halo[R + 1] = U[r0 + R] if r0 + R < N else U[r0 + R - 1]
        Uh = self.be.to_device(halo)
        self._ck(self.lib.chs_slab_prepare(self._h, self.be.ptr(Uh), self._mean), "chs_slab_prepare")
        self._allreduce_vec()
        self._ck(self.lib.chs_slab_control(self._h, 0, 2), "chs_slab_control")
        self.get_state(0)
        return self.be.download(self.rows[:1, :])

    def begin(self):
        lib, h = self.lib, self._h
        # the conserved mean (Q4) is that of the field hat_U is recomputed from (solver.py:159): after a jittered
        # call that is the jittered field (quirk Q2), not U_init
        m = self._tensor(self.U).sum(dtype=self._torch().float64)
        if self.P > 1:
            import torch.distributed as dist
            m = m.reshape(1).clone()
            dist.all_reduce(m)
        self._mean = float(m) / (self.N * self.N)
        self._ck(lib.chs_slab_begin(h), "chs_slab_begin")
        self._row(self.S_FWD, self.U, self.A)            # x-transform of U
        self._transpose(self.A, self.B)
        self._row(self.S_YFWD, self.B, self.H)           # y-transform -> hat_U' (solver.py:159)
        self._row(self.S_MU, self.U, self.A)             # mu(U) -> x-transform, ||mu||^2
        self._ck(lib.chs_slab_clear_yedge(h), "chs_slab_clear_yedge")
        self._ck(lib.chs_slab_reduce(h, self.R, 0), "chs_slab_reduce")
        self._allreduce_vec()
        cols = None
        if self._ps.adaptive_time:
            if self._want_cols(self._cs, False):         # re-entry: the first iteration may already update delt
                self._colsum()
            elif self._cols is None:
                self._cols = self.be.empty((self.N,))
                self._cols_scr = self.be.empty((16 * self.N,))
            cols = self.be.ptr(self._cols)
        self._ck(lib.chs_slab_control_dyn(h, 0, 0, None, cols), "chs_slab_control_dyn")
        if self._peer is not None:
            self._hdl.barrier()                          # every rank is done reading B before it is rewritten
        self._transpose(self.A, self.B)                  # every step starts from B = transpose(A)

    def _want_cols(self, cs_next, last):
        """Will the iteration that follows update delt (solver.py:177-181)?  cs_next = computed_steps it starts with."""
        return bool(self._ps.adaptive_time) and not last and cs_next > 500 and cs_next % 2 == 0

    def _colsum(self):
        """All-rank column sums of delt_max/sqrt(1+62.5 mu(U)^2) -> self._cols (adaptive dt, solver.py:182-183)."""
        if self._cols is None:
            self._cols = self.be.empty((self.N,))
            self._cols_scr = self.be.empty((16 * self.N,))
        self._ck(self.lib.chs_slab_colsum(self._h, self.be.ptr(self._cols), self.be.ptr(self._cols_scr)), "chs_slab_colsum")
        if self.P > 1:
            import torch.distributed as dist
            dist.all_reduce(self._tensor(self._cols))    # sum over the row slabs; identical bits on every rank

    def _halo_rows(self):
        """Device pointers to the boundary rows of the neighbouring slabs (own rows at the domain edges)."""
        be, R, N = self.be, self.R, self.N
        if self.P == 1:
            return be.ptr(self.U), be.ptr(self.U)
        import torch
        import torch.distributed as dist
        U = self._tensor(self.U)
        mine = torch.stack([U[0], U[R - 1]]).contiguous()
        allb = [torch.empty_like(mine) for _ in range(self.P)]
        dist.all_gather(allb, mine)
        self._halo = allb                                # keep alive until the kernel has run
        top = allb[self.rank - 1][1] if self.rank > 0 else U[0]
        bot = allb[self.rank + 1][0] if self.rank < self.P - 1 else U[R - 1]
        return top.data_ptr(), bot.data_ptr()

    def _step(self, last, noise=None, noise_mean=None):
        """One CH step.  noise / noise_mean: device pointers to this step's draws for the rank's rows and to the
        mean of the whole draw (jitter), or None."""
        lib, h, be, R, N = self.lib, self._h, self.be, self.R, self.N
        esz = 8 * N
        jit = noise is not None
        adaptive = bool(self._ps.adaptive_time)

        def y_pass(r0, rc):      # hat_mu' = DCT(B); H = (H + Seig*hat_mu')/CHeig; B = IDCT(H): one kernel
            self._ck(lib.chs_slab_update(h, be.ptr(self.H) + r0 * esz, be.ptr(self.B) + r0 * esz, rc,
                                         self.row_base + r0), "chs_slab_update")

        def x_pass(r0, rc):      # (jitter,) U_new stored, diagnostics, mu, x-transform
            self._ck(lib.chs_slab_step_x(h, be.ptr(self.A) + r0 * esz, be.ptr(self.A) + r0 * esz, rc, self.row_base + r0,
                                         float(self._mean), noise, noise_mean), "chs_slab_step_x")

        def after_x():           # what needs the complete stored field of this step
            if jit:              # stencil gradient energy of the jittered field (1-row halo from the neighbours)
                top, bot = self._halo_rows()
                self._ck(lib.chs_slab_grad(h, top, bot), "chs_slab_grad")
            if self._want_cols(self._cs + 1, last):
                self._colsum()

        cols = be.ptr(self._cols) if (adaptive and self._cols is not None) else None
        edges = (0, 0) if jit else (int(self.rank == 0), int(self.rank == self.P - 1))
        self._pass(y_pass, self.B, self.A)
        if self._peer is not None and (self._nchunks == 1 or self._ce):
            # peer-memory route: x pass, its exchange and the 7 sums (stored into every rank's gather buffer by
            # the sums kernel) share ONE device-side barrier; the control kernel adds the ranks' sums in rank
            # order -- 8 launches per step, no collective
            if self._ce:
                self._pass_ce(x_pass, self.A, self.B, barrier=False)
                after_x()
            else:
                x_pass(0, R)
                after_x()
                self._transpose(self.A, self.B, sync=False)
            cols = be.ptr(self._cols) if (adaptive and self._cols is not None) else None
            par = self._parity
            self._parity ^= 1
            goff = be.ptr(self._gather) - self._ab_base + (par * self.P + self.rank) * 64
            slots = (C.c_uint64 * self.P)(*[self._peer[p] + goff for p in range(self.P)])
            self._ck(lib.chs_slab_sums_peers(h, edges[0], edges[1], slots), "chs_slab_sums_peers")
            self._hdl.barrier()
            self._ck(lib.chs_slab_control_dyn(h, int(bool(last)), 1, be.ptr(self._gather) + par * self.P * 64, cols),
                     "chs_slab_control_dyn")
            self._cs += 1
            return
        if jit or adaptive:      # the whole field of the step must be stored before the stencil / column sums
            x_pass(0, R)
            after_x()
            cols = be.ptr(self._cols) if (adaptive and self._cols is not None) else None
            self._transpose(self.A, self.B)
        else:
            self._pass(x_pass, self.A, self.B)           # B for the next step (its barrier also covers the sums)
        # per-tile partials + y-edge terms of the stored field -> the 7 local sums, one launch
        self._ck(lib.chs_slab_sums(h, edges[0], edges[1]), "chs_slab_sums")
        self._allreduce_vec()
        self._ck(lib.chs_slab_control_dyn(h, int(bool(last)), 1, None, cols), "chs_slab_control_dyn")
        self._cs += 1

    def get_state(self, sim=0):
        st = _lib.State()
        rw, halted = C.c_int64(0), C.c_int32(0)
        self._ck(self.lib.chs_slab_get_state(self._h, C.byref(st), C.byref(rw), C.byref(halted)), "chs_slab_get_state")
        self._rw, self._halted = int(rw.value), int(halted.value)
        return st

    def set_state(self, sim, st):
        self._ck(self.lib.chs_slab_set_state(self._h, C.byref(st)), "chs_slab_set_state")

    def pcg64_noise(self, bit_generator_state, n):
        """(noise [n][R][N] for this rank's rows, global per-step means [n]) on the device: the next n*N*N doubles of
        a numpy PCG64 generator, bit-identical to rng.random((N, N)) per step (each rank generates only its rows)."""
        st = bit_generator_state["state"]
        s128, i128 = int(st["state"]), int(st["inc"])
        m64 = (1 << 64) - 1
        R, N = self.R, self.N
        noise = self.be.empty((n, R, N))
        mean = self.be.empty((n,))
        for i in range(n):
            self._ck(self.lib.chs_slab_pcg64_fill(self._h, s128 >> 64, s128 & m64, i128 >> 64, i128 & m64,
                                                  i * N * N + self.row_base * N, self.be.ptr(noise) + i * R * N * 8, R * N),
                     "chs_slab_pcg64_fill")
        self._ck(self.lib.chs_slab_row_means(self._h, self.be.ptr(noise), n, R * N, self.be.ptr(mean)), "chs_slab_row_means")
        if self.P > 1:
            import torch.distributed as dist
            mt = self._tensor(mean)
            dist.all_reduce(mt)
            mt /= self.P
        return noise, mean

    def run(self, iters, draw_noise=None, poll_every=None):
        done = np.zeros(1, np.int64)
        if iters <= 0:
            return [np.empty((0, 9))], done
        chunk = self.rows_cap if poll_every is None else min(self.rows_cap, int(poll_every))
        if draw_noise is not None:                       # noise of a whole chunk is resident: bound it to ~1 GiB
            chunk = max(1, min(chunk, (1 << 30) // (self.R * self.N * 8)))
        out = []
        self._cs = int(self.get_state(0).computed_steps)
        self.begin()
        n_done = 0
        self.get_state(0)
        while n_done < iters and not self._halted:
            n = min(chunk, iters - n_done)
            noise = nmean = None
            if draw_noise is not None:
                drawn = draw_noise(n)
                if isinstance(drawn, tuple):              # generated on the device (PCG64 kernel), this rank's rows
                    noise, nmean = drawn
                else:                                     # host draws (Sobol): this rank's rows are uploaded
                    r0 = self.row_base
                    noise = self.be.to_device(np.ascontiguousarray(drawn[:, r0:r0 + self.R, :]))
                    nmean = self.be.to_device(drawn.reshape(n, -1).mean(axis=1))
            for i in range(n):
                self._step(last=(n_done + i + 1 == iters),
                           noise=None if noise is None else self.be.ptr(noise) + i * self.R * self.N * 8,
                           noise_mean=None if nmean is None else self.be.ptr(nmean) + i * 8)
            self.get_state(0)
            if self._rw:
                out.append(self.be.download(self.rows[:self._rw, :]))
                done[0] += self._rw
                self._ck(self.lib.chs_slab_rewind_rows(self._h), "chs_slab_rewind_rows")
            n_done += n
        return [np.concatenate(out) if out else np.empty((0, 9))], done

    def get_U(self, sim=None):
        """Full field on the host (gathered over the ranks)."""
        if self.P == 1:
            return self.be.download(self.U)
        import torch
        import torch.distributed as dist
        mine = self._tensor(self.U)
        parts = [torch.empty_like(mine) for _ in range(self.P)]
        dist.all_gather(parts, mine)
        return torch.cat(parts, dim=0).cpu().numpy()

    def launch_count(self):
        return int(self.lib.chs_slab_launch_count(self._h))


class BigEngine(SlabEngine):
    """Arbitrary N (the reference accepts any N): sizes that are neither a power of two (FFT kernels) nor <= 104
    (one-CTA GEMM kernel).  One simulation on one GPU; the 2-D DCT-II / DCT-III are FP64 tensor-core GEMMs
    C.X.C^T / C^T.Y.C over global memory (csrc/chs_big.cuh), the rest is the slab path's elementwise, reduction
    and control machinery.  Same interface as SlabEngine / BatchStepper."""

    def __init__(self, N, param_struct, rows_cap=1024, backend=None, device=None):
        from .solver import _CudaBackend
        self.be = backend if backend is not None else _CudaBackend(device)
        lib = self.lib = self.be.lib
        self.N = n = int(N)
        if not lib.chs_big_supports_n(n):
            raise ValueError(f"N={N} is not supported by the arbitrary-N path (8..2048)")
        self.rank, self.P, self.R, self.row_base = 0, 1, n, 0
        self.batch, self.rows_cap = 1, int(rows_cap)
        self.n8 = n8 = (n + 7) // 8 * 8
        self.ld = n8
        self.U = self.be.empty((n, n))
        self.rows = self.be.empty((self.rows_cap, 9))
        # Cm | Ct | H (hat_U) | A (mu) | T | M (hat_mu, then the new field): zero-padded n8 x ld operands
        self.mats = self.be.empty((6, n8, self.ld))
        if self.be.name == "cuda":
            self.mats.zero_()
        else:
            self.mats[...] = 0.0
        k = np.arange(n, dtype=np.longdouble)[:, None]
        x = np.arange(n, dtype=np.longdouble)[None, :]
        Cm = np.sqrt(np.longdouble(2) / n) * np.cos(np.pi * k * (2 * x + 1) / (2 * np.longdouble(n)))
        Cm[0, :] = np.sqrt(np.longdouble(1) / n)
        pad = np.zeros((2, n8, self.ld))
        pad[0, :n, :n] = Cm.astype(np.float64)               # orthonormal DCT-II matrix (scipy.fftpack.dctn norm='ortho')
        pad[1, :n, :n] = Cm.T.astype(np.float64)
        self.be.upload(self.mats[:2], pad)
        self._peer = None
        self._nchunks = 1
        self._main = self._side = None
        wbytes = lib.chs_slab_workspace_bytes(n, n)
        self.work = self.be.empty((wbytes,), "u1")
        lam = np.ascontiguousarray(utils.laplace_spectrum_1d(n), dtype=np.float64)
        self._ps = param_struct
        self._h = _lib.check(lib, lib.chs_slab_create(self.be.device_index(), n, n, 0, 1, 0, C.byref(param_struct),
                                                       self.be.ptr(self.U), self.be.ptr(self.rows), self.rows_cap,
                                                       self.be.ptr(self.work), wbytes, lam.ctypes.data,
                                                       self.be.stream_handle()), "chs_slab_create")
        self._full = None
        self._mean = 0.0
        self._prof = None
        self._cols = self._cols_scr = self._halo = None
        self._cs = 1

    def _m(self, i):
        return self.be.ptr(self.mats) + i * self.n8 * self.ld * 8

    def _gemm(self, a, b, d):
        self._ck(self.lib.chs_big_gemm(self._h, self._m(a), self._m(b), self._m(d), self.n8, self.ld), "chs_big_gemm")

    CM, CT, H_, A_, T_, M_ = range(6)

    def begin(self):
        lib, h, be = self.lib, self._h, self.be
        m = self._tensor(self.U).sum(dtype=self._torch().float64)
        self._mean = float(m) / (self.N * self.N)            # conserved mean of the field hat_U is recomputed from (Q2, Q4)
        self._ck(lib.chs_slab_begin(h), "chs_slab_begin")
        self._ck(lib.chs_big_copy(h, be.ptr(self.U), self.N, self._m(self.M_), self.ld, 0), "chs_big_copy")
        self._gemm(self.CM, self.M_, self.T_)                # hat_U = C . U . C^T          (solver.py:159)
        self._gemm(self.T_, self.CT, self.H_)
        self._ck(lib.chs_big_phys(h, None, self._m(self.A_), self.ld, float(self._mean), None, None, 0, 1), "chs_big_phys")
        self._ck(lib.chs_big_sums(h), "chs_big_sums")
        cols = None
        if self._ps.adaptive_time:
            if self._want_cols(self._cs, False):
                self._colsum()
            elif self._cols is None:
                self._cols = self.be.empty((self.N,))
                self._cols_scr = self.be.empty((16 * self.N,))
            cols = be.ptr(self._cols)
        self._ck(lib.chs_slab_control_dyn(h, 0, 0, None, cols), "chs_slab_control_dyn")

    def _step(self, last, noise=None, noise_mean=None):
        lib, h, be = self.lib, self._h, self.be
        self._gemm(self.A_, self.CT, self.T_)                # hat_mu = C . mu . C^T        (solver.py:201)
        self._gemm(self.CM, self.T_, self.M_)
        self._ck(lib.chs_big_update(h, self._m(self.H_), self._m(self.M_), self.ld), "chs_big_update")
        self._gemm(self.H_, self.CM, self.T_)                # U = C^T . hat_U . C          (solver.py:208)
        self._gemm(self.CT, self.T_, self.M_)
        self._ck(lib.chs_big_phys(h, self._m(self.M_), self._m(self.A_), self.ld, float(self._mean), noise, noise_mean,
                                  1, 0), "chs_big_phys")
        self._ck(lib.chs_slab_grad(h, be.ptr(self.U), be.ptr(self.U)), "chs_slab_grad")     # np.gradient energy (one rank: no halo)
        if self._want_cols(self._cs + 1, last):
            self._colsum()
        cols = be.ptr(self._cols) if (self._ps.adaptive_time and self._cols is not None) else None
        self._ck(lib.chs_big_sums(h), "chs_big_sums")
        self._ck(lib.chs_slab_control_dyn(h, int(bool(last)), 1, None, cols), "chs_slab_control_dyn")
        self._cs += 1
