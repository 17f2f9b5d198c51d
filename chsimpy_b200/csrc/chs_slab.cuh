// Slab path: one large N x N domain, row-slab decomposed over P ranks (P = 1: one GPU).
// Every transform pass works on CONTIGUOUS rows, so the same row machinery serves both
// directions; the direction change is a transpose -- a local tiled transpose for P = 1, an
// NCCL all-to-all (issued by the host layer, chsimpy_b200/slab.py) plus the local
// pack/unpack kernels below for P > 1.  One CH step (reference solver.py:165-249):
//
//   A  = rowDCT(mu(U))                       k_slab_row<S_STEP> of the previous step / <S_MU>
//   B  = transpose(A)                        pack -> all-to-all -> unpack   (x-slot rows, y cols)
//   B  = rowIDCT(H = (H + Seig*rowDCT(B))/CHeig), spectral gradient energy
//                                            k_slab_row<S_YSTEP>: ONE pass, the spectral update sits
//                                            between the last forward and the first inverse stage
//                                            (H = hat_U': x-slot rows, natural ky columns)
//   A  = transpose(B)                        pack -> all-to-all -> unpack
//   U, A = rowIDCT(A), physics, rowDCT(mu)   k_slab_row<S_STEP>   (per-tile partial sums)
//   sums -> all-reduce -> k_slab_control     (TimeData row, NaN flag, stop test, time accounting)
#pragma once
#include "chs_kernels.cuh"

namespace CHS_NS {

enum { S_FWD = 0, S_MU = 1, S_INV = 2, S_STEP = 3, S_YFWD = 4, S_YSTEP = 5 };
// reduced vector layout (all-reduced over ranks)
enum { R_GE = 0, R_EDGE, R_F, R_ABS, R_MU2, R_CNT, R_RA, R_NVAL };

struct SlabArgs {
    const double* src;
    double* dst;
    double* Uout;               // S_STEP: U_new rows are stored here
    int rows;                   // rows of this launch (multiple of LINES)
    int tile0, tiles_total;     // position of the launch's first tile among the rank's tiles (partial sums)
    int row_base;               // global index of local row 0
    int diag;                   // S_MU / S_STEP: accumulate the diagnostics of the field
    double mean_u;              // conserved mean of U (PS)
    double* part;               // [R_NVAL][rows/LINES] per-tile partial sums
    double* H;                  // S_YFWD / S_YSTEP: hat_U' rows (x-slot row_base + r, natural ky)
    double* part_ge;            // S_YSTEP: [rows/LINES] spectral gradient-energy sums
    const double* lam;
    const double* gsin;
    const double2* lamg;        // packed per-item lambda / g table (KArgs::lamg)
    const int* kof;
    Sim* S;
    const double* noise;        // S_STEP with jitter: uniform draws of this step for the launch's rows [rows][N], else null
    const double* noise_mean;   // ... and the mean of the WHOLE N x N draw (device scalar)
    const double2* tw;          // natural table (point-major geometry), or the per-stage tables (line-major)
    const double2* om;
    const double2* logtab;
};

template <int N, int MODE>
CHS_KERNEL void __launch_bounds__(Geo<N>::NT, Geo<N>::MINB) k_slab_row(SlabArgs a) {
    CHS_PDL_TRIGGER();
    CHS_PDL_WAIT();               // programmatic dependent launch: nothing is read before the predecessor is complete
    // device-side stop flag (energy stop, time limit, NaN): a stopped simulation is frozen -- U, H, A, B
    // keep the state of the stopping step whatever the host still has queued (solver.py:199,247 break)
    if ((MODE == S_STEP || MODE == S_YSTEP) && a.S->halted) return;
    using G = Geo<N>;
    constexpr int M = G::M, LINES = G::LINES, TPL = G::TPL, NT = G::NT;
    constexpr int NST = Rad<M>::nst;
    constexpr int R0 = Rad<M>::radix(0), ST0 = M / R0, NB0 = 16 / R0;
    CHS_SMEM_DECL
    double* sm = reinterpret_cast<double*>(CHS_SMEM_PTR);
    double2* sc = reinterpret_cast<double2*>(sm);
    double* edge = sm + G::OFF_EDGE;
    double* ra_scr = sm + G::OFF_RA;
    const int tid = threadIdx.x, l = G::line_of(tid), t = G::t_of(tid);
    double2* scl = sc + l * G::LOFF;
    const int ntiles = a.rows / LINES;
    const bool physics_on = (MODE == S_MU) || (MODE == S_STEP);
    const bool diag = physics_on && a.diag;
    const double2* ltab = physics_on ? stage_logtab<G>(sm, a.logtab, tid) : nullptr;
    const int ra_row = N / 2 + 1;
    CHS_TILE_LOOP(tile, ntiles) {
        const int row0 = tile * LINES;
        const size_t goff = (size_t)row0 * N;
        if (MODE == S_YFWD || MODE == S_YSTEP) {
            // y pass: rows are x-slots, the row index runs over y.  Forward DCT, then (S_YSTEP) the
            // spectral update against hat_U' and the inverse DCT without leaving the tile.
            constexpr int CM = (MODE == S_YFWD) ? COL_FWD : COL_STEP;
            row_tile_load_phys_async<N>(sm, a.src + goff, tid);
            double lam1 = 0, lam2 = 0, lamx = 0, gxs = 0;
            if (MODE == S_YSTEP) {
                const double* hp = a.H + goff;                        // consumed mid-tile: pull into L2 now
                for (int i = tid * 16; i < N * LINES; i += NT * 16) CHS_PREFETCH_L2(hp + i);
                const double delx2 = a.S->p.delx * a.S->p.delx;
                lam1 = a.S->delt_coef / delx2;                        // utils.py:41-42
                lam2 = a.S->p.kappa_tilde * lam1 / delx2;
                const int kx = a.kof[a.row_base + row0 + l];
                lamx = a.lam[kx];
                gxs = a.gsin[kx];
            }
            chs_cp_async_wait_all();
            __syncthreads();
            fft_fwd_range<N, 0, NST - 1>(scl, t, a.tw);
            ColMid<N, CM> mid;
            mid.om = a.om; mid.lamg = a.lamg;
            mid.hstride = 1;
            mid.hat = a.H + goff + (size_t)l * N;
            mid.hat_in = mid.hat;
            mid.hat00 = nullptr;
            mid.ge = 0; mid.lam1 = lam1; mid.lam2 = lam2; mid.lamx = lamx; mid.gx = gxs;
            fused_units<N, true, MODE == S_YSTEP>(scl, t, mid);
            if (MODE == S_YSTEP) {
                const double v[1] = {mid.ge};
                reduce_stage<1>(v, sm + G::OFF_RED, tid);
                __syncthreads();
                fft_inv_range<N, 0, NST - 1>(scl, t, a.tw);
                if (tid == 0) {
                    double s1[1];
                    reduce_final<1>(s1, sm + G::OFF_RED, NT);
                    a.part_ge[a.tile0 + tile] = s1[0];
                }
                row_tile_store_phys<N>(sm, a.dst + goff, tid);
            }
        } else {
        if (MODE == S_INV || MODE == S_STEP) row_tile_load_slots_async<N>(sc, a.src + goff, tid);
        else row_tile_load_phys_async<N>(sm, a.src + goff, tid);
        const bool ra_line = diag && (a.row_base + row0 + l == ra_row);
        const bool ra_tile = diag && (ra_row >= a.row_base + row0) && (ra_row < a.row_base + row0 + LINES);
        chs_cp_async_wait_all();
        __syncthreads();
        if (MODE == S_INV || MODE == S_STEP) {
            if (ra_line && t == 0) ra_scr[0] = scl[G::idx(0)].x * sqrt(1.0 / N);
            {
                RowPre<N> pre{a.om};
                fused_units<N, false, true>(scl, t, pre);
            }
            __syncthreads();
            fft_inv_range<N, 0, NST - 1>(scl, t, a.tw);        // includes stage 0: the field is stored below
        }
        if (MODE == S_INV) {
            row_tile_store_phys<N>(sm, a.dst + goff, tid);
        } else {
            const bool jit = (MODE == S_STEP) && (a.noise != nullptr);
            if (jit) {
                // jitter (solver.py:210-211): U += jitter*(2*noise - 1) on the field in shared memory; the jittered
                // field is what gets stored, diagnosed and fed to mu (hat_U' is NOT touched, quirk Q2)
                const double jv = a.S->p.jitter;
                const double* nz = a.noise + goff;
                constexpr int CNT = LINES * N / NT;
#pragma unroll 4
                for (int j = 0; j < CNT; ++j) {
                    const int i = tid + j * NT;
                    double* q = sm + real_off<N>(mk_pos<N>(i % N)) + 2 * G::LOFF * (i / N);
                    *q += jv * (2.0 * nz[i] - 1.0);
                }
                __syncthreads();
                if (ra_tile) {                                          // mean of the jittered Ra row
                    if (ra_line) {
                        double sj = 0;
                        for (int i = 0; i < 16; ++i) {
                            const double2 v = scl[G::idx(t + i * TPL)];
                            sj += v.x + v.y;
                        }
                        ra_scr[2 + t] = sj;
                    }
                    __syncthreads();
                    if (ra_line && t == 0) {
                        double sj = 0;
                        for (int jj = 0; jj < TPL; ++jj) sj += ra_scr[2 + jj];
                        ra_scr[0] = sj / (double)N;
                    }
                    __syncthreads();
                }
            }
            if (MODE == S_STEP) {
                row_tile_store_phys<N>(sm, a.Uout + goff, tid);  // reads only; no barrier needed before the in-place physics
                __syncthreads();
            }
            if (physics_on) {
                const chs_params p = a.S->p;
                PhysK pk;
                pk.th.RT = p.RT; pk.th.mBRT = -p.BRT; pk.th.A0 = p.A0; pk.th.A1 = p.A1; pk.th.m2A1 = -2.0 * p.A1; pk.th.B = p.B;
                pk.threshold = p.threshold;
                // conserved mean (Q4); the jitter shifts it by jitter*(2*mean(noise) - 1)
                pk.meanU = a.mean_u + (jit ? p.jitter * (2.0 * a.noise_mean[0] - 1.0) : 0.0);
                const double ra_mean = ra_line ? ra_scr[0] : 0.0;
                RowAcc acc = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
                for (int i = 0; i < NB0; ++i) {
                    const int j = t + i * TPL;
                    double xr[R0], xi[R0];
#pragma unroll
                    for (int q = 0; q < R0; ++q) {
                        const double2 v = scl[G::idx(j) + q * G::step(ST0)];
                        xr[q] = v.x; xi[q] = v.y;
                    }
                    // the Ra row is one row of the whole domain: keep its per-value work out of the common copy
                    if (ra_line) physics<N, R0, G::LOG_STRIDE, true>(xr, xi, j, pk, ltab, diag, true, ra_mean, acc, edge + 4 * l);
                    else physics<N, R0, G::LOG_STRIDE, true>(xr, xi, j, pk, ltab, diag, false, 0.0, acc, edge + 4 * l);
                    dft<R0, false>(xr, xi);
#pragma unroll
                    for (int q = 1; q < R0; ++q) {
                        const double2 w = G::STAGED_TABLES ? __ldg(a.tw + (q - 1) * ST0 + j) : tab_tw<N>(a.tw, j * q);
                        const double x = xr[q], y = xi[q];
                        xr[q] = x * w.x - y * w.y;
                        xi[q] = x * w.y + y * w.x;
                    }
#pragma unroll
                    for (int q = 0; q < R0; ++q) scl[G::idx(j) + q * G::step(ST0)] = make_double2(xr[q], xi[q]);
                }
                if (ra_line) ra_scr[2 + t] = acc.ra;
                const double v[4] = {chs_fma(pk.th.RT, acc.fa + acc.fb, acc.fp), acc.ab, acc.mu2, (double)acc.cnt};
                reduce_stage<4>(v, sm + G::OFF_RED, tid);
                __syncthreads();
                if (tid == 0) {
                    double s4[4];
                    reduce_final<4>(s4, sm + G::OFF_RED, NT);
                    double* pp = a.part + a.tile0 + tile;
                    const int nt_all = a.tiles_total;
                    double e = 0, ra = 0;
                    if (diag) {
                        for (int l2 = 0; l2 < LINES; ++l2) {
                            const double* eg = edge + 4 * l2;
                            e += (eg[1] - eg[0]) * (eg[1] - eg[0]) + (eg[3] - eg[2]) * (eg[3] - eg[2]);
                        }
                        if (ra_tile) {
                            for (int jj = 0; jj < TPL; ++jj) ra += ra_scr[2 + jj];
                            ra /= (double)N;
                        }
                    }
                    pp[R_GE * nt_all] = 0;
                    pp[R_EDGE * nt_all] = jit ? 0.0 : 0.75 * e;        // with jitter the gradient energy is the stencil sum of k_slab_grad
                    pp[R_F * nt_all] = s4[0];
                    pp[R_ABS * nt_all] = s4[1];
                    pp[R_MU2 * nt_all] = s4[2];
                    pp[R_CNT * nt_all] = s4[3];
                    pp[R_RA * nt_all] = ra;
                }
            } else {
                fft_stage<N, 0, false>(scl, t, a.tw);
                __syncthreads();
            }
            fft_fwd_range<N, 1, NST - 1>(scl, t, a.tw);
            {
                RowPost<N> post{a.om};
                fused_units<N, true, false>(scl, t, post);
            }
            __syncthreads();
            row_tile_store_slots<N>(sc, a.dst + goff, tid);
        }
        }
        __syncthreads();
    }
    chs_cp_async_wait_all();
}

// out[c][r] = in[r][c] for an R x C row-major block, through a 32x33 shared tile.
// in_ld / out_ld: leading dimensions; used for the local transpose (P = 1) and for the
// per-peer blocks around the all-to-all (P > 1).
CHS_KERNEL void k_slab_transpose(const double* in, double* out, int R, int C, int in_ld, int out_ld) {
    CHS_PDL_TRIGGER();
    CHS_PDL_WAIT();
    CHS_SMEM_DECL
    double* tile = reinterpret_cast<double*>(CHS_SMEM_PTR);       // 32*33 doubles
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;       // 256 threads: 32 x 8
    for (int k = ty; k < 32; k += 8)
        if (by + k < R && bx + tx < C) tile[k * 33 + tx] = in[(size_t)(by + k) * in_ld + bx + tx];
    __syncthreads();
    for (int k = ty; k < 32; k += 8)
        if (bx + k < C && by + tx < R) out[(size_t)(bx + k) * out_ld + by + tx] = tile[tx * 33 + k];
}

// The exchange of one pass as ONE launch over all peers: blockIdx.z = i handles the block for rank
// p = (rank + i) % P (local block first, link load spread): out_p[c][r] = in[r][p*C + c], written straight
// into rank p's buffer over NVLink (dst.p[p] = peer-mapped destination, offset to this rank's columns).
struct PeerPtrs { double* p[8]; };
CHS_KERNEL void k_slab_transpose_peers(PeerPtrs dst, const double* in, int R, int C, int in_ld, int out_ld, int rank, int P) {
    CHS_PDL_TRIGGER();
    CHS_PDL_WAIT();
    CHS_SMEM_DECL
    double* tile = reinterpret_cast<double*>(CHS_SMEM_PTR);       // 32*33 doubles
    const int peer = (rank + (int)blockIdx.z) % P;
    const double* src = in + (size_t)peer * C;
    double* out = dst.p[peer];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;       // 256 threads: 32 x 8
    for (int k = ty; k < 32; k += 8)
        if (by + k < R && bx + tx < C) tile[k * 33 + tx] = src[(size_t)(by + k) * in_ld + bx + tx];
    __syncthreads();
    for (int k = ty; k < 32; k += 8)
        if (bx + k < C && by + tx < R) out[(size_t)(bx + k) * out_ld + by + tx] = tile[tx * 33 + k];
}

// The same exchange with 4096-element tiles whose transposed rows leave the SM as BULK asynchronous stores
// (cp.async.bulk shared -> global, SASS UBLKCP): one 1 KB copy per output row segment instead of 8-byte stores of a
// warp -- fewer, larger write packets on NVLink, and no LSU instruction for the remote side.  Needs R % SLAB_TR == 0
// and C % SLAB_TC == 0.
// Tile in shared memory: out row c at pitch TP doubles (16-byte aligned rows for the bulk copies; the 2-way bank
// conflicts of the transposing 8-byte writes are far below what the links can take).
// Tile: SLAB_TC input columns (= output rows) x SLAB_TR input rows (= bytes/8 of one bulk store).
#ifndef CHS_SLAB_TR
#define CHS_SLAB_TR 128
#endif
constexpr int SLAB_TR = CHS_SLAB_TR, SLAB_TC = 4096 / SLAB_TR, SLAB_TP = SLAB_TR + 2;
static_assert(SLAB_TC >= 32 && SLAB_TC % 32 == 0 && SLAB_TR % 16 == 0, "tile shape");
struct PeerDst { double* p[8]; int ld[8]; };                      // destination base and leading dimension per rank
CHS_KERNEL void k_slab_transpose_bulk(PeerDst dst, const double* in, int R, int C, int in_ld, int rank, int P) {
    CHS_PDL_TRIGGER();
    CHS_PDL_WAIT();
    CHS_SMEM_DECL
    double* tile = reinterpret_cast<double*>(CHS_SMEM_PTR);       // SLAB_TC * SLAB_TP doubles
    const int peer = (rank + (int)blockIdx.z) % P;
    const double* src = in + (size_t)peer * C;
    double* out = dst.p[peer];
    const int out_ld = dst.ld[peer];
    const int bx = blockIdx.x * SLAB_TC, by = blockIdx.y * SLAB_TR;
    const int tid = threadIdx.x;                                  // 256 threads
    // in[by + r][bx + c]: a warp reads 32 consecutive columns of one row (256 contiguous bytes) and writes them to 32
    // tile rows (pitch = 2 mod 32 doubles, i.e. 4 banks apart: 2-way conflicts); the 16 loads of a thread are in
    // flight together
    constexpr int RS = 256 / SLAB_TC;                             // rows covered by one pass of the CTA
    const int c = tid % SLAB_TC, r0 = tid / SLAB_TC;
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = src[(size_t)(by + r0 + RS * k) * in_ld + bx + c];
#pragma unroll
    for (int k = 0; k < 16; ++k) tile[c * SLAB_TP + r0 + RS * k] = v[k];
    chs_fence_async_smem();
    __syncthreads();
    if (tid < SLAB_TC) {
        chs_bulk_s2g(out + (size_t)(bx + tid) * out_ld + by, tile + tid * SLAB_TP, SLAB_TR * 8);
        chs_bulk_commit_wait();
    }
}

// y-edge terms of the gradient energy from two stored rows of U: 3/4 * sum_x (U[r1][x]-U[r0][x])^2
CHS_KERNEL void k_slab_yedge(const double* r0, const double* r1, int N, double* out, int accumulate) {
    CHS_SMEM_DECL
    double* red = reinterpret_cast<double*>(CHS_SMEM_PTR);
    double s = 0;
    for (int x = threadIdx.x; x < N; x += blockDim.x) { const double d = r1[x] - r0[x]; s += d * d; }
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (unsigned j = 0; j < blockDim.x; ++j) t += red[j];
        out[0] = (accumulate ? out[0] : 0.0) + 0.75 * t;
    }
}

// adaptive dt (solver.py:182-183): per-column sums of delt_max/sqrt(1 + 62.5 mu(U)^2) over this rank's rows,
// in two fixed-order levels (gridDim.y row chunks -> part[chunk][N], then k_slab_colsum_final): deterministic
CHS_KERNEL void k_slab_colsum(const double* U, int rows, int N, const Sim* S, const double2* logtab, double* part) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= N) return;
    const chs_params p = S->p;
    ThermoK k;
    k.RT = p.RT; k.mBRT = -p.BRT; k.A0 = p.A0; k.A1 = p.A1; k.m2A1 = -2.0 * p.A1; k.B = p.B;
    const int per = (rows + gridDim.y - 1) / gridDim.y;
    const int y0 = blockIdx.y * per, y1 = (y0 + per < rows) ? y0 + per : rows;
    double s = 0, fa = 0, fb = 0, fp = 0;
    for (int y = y0; y < y1; ++y) {
        const double mu = thermo_acc<1, true>(U[(size_t)y * N + x], k, logtab, fa, fb, fp);
        s += p.delt_max / sqrt(1.0 + 62.5 * (mu * mu));
    }
    part[(size_t)blockIdx.y * N + x] = s;
}
CHS_KERNEL void k_slab_colsum_final(const double* part, int nchunks, int N, double* colsum) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= N) return;
    double s = 0;
    for (int c = 0; c < nchunks; ++c) s += part[(size_t)c * N + x];
    colsum[x] = s;
}

// gradient energy of the stored (jittered) field by np.gradient's stencils (solver.py:213-217): raw sum of
// (gy^2 + gx^2) h^2 over this rank's rows; top / bot = the neighbouring ranks' boundary rows (ignored at the
// domain edges); per-block partial sums in part[gridDim.x]
CHS_KERNEL void k_slab_grad(const double* U, const double* top, const double* bot, int rows, int row_base, int N, double* part) {
    CHS_SMEM_DECL
    double* red = reinterpret_cast<double*>(CHS_SMEM_PTR);
    double v = 0;
    const size_t total = (size_t)rows * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / N), x = (int)(i % N);
        const int gy_ = row_base + y;
        const double c = U[i];
        const double up = (y > 0) ? U[i - N] : top[x], dn = (y < rows - 1) ? U[i + N] : bot[x];
        double gy, gx;
        if (gy_ == 0) gy = dn - c;
        else if (gy_ == N - 1) gy = c - up;
        else gy = 0.5 * (dn - up);
        if (x == 0) gx = U[i + 1] - c;
        else if (x == N - 1) gx = c - U[i - 1];
        else gx = 0.5 * (U[i + 1] - U[i - 1]);
        v += gy * gy + gx * gx;
    }
    red[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (unsigned j = 0; j < blockDim.x; ++j) t += red[j];
        part[blockIdx.x] = t;
    }
}

// Solver.prepare() on a slab with one halo row above and below (Uh = [rows+2][N], row 0 and
// row rows+1 are the neighbours' rows; ignored at the domain edges): np.gradient stencils,
// free energy, |U - mean|, and Ra of global row N/2+1; per-block partial sums [4][gridDim.x].
CHS_KERNEL void k_slab_prepare(const double* Uh, int rows, int row_base, int N, double mean_u, const Sim* S,
                               const double2* logtab, double* part) {
    CHS_SMEM_DECL
    double* red = reinterpret_cast<double*>(CHS_SMEM_PTR);      // 4 * blockDim.x
    const chs_params p = S->p;
    const double* U = Uh + N;                                    // local row 0
    const int ra_row = N / 2 + 1 - row_base;                     // local index (may be out of range)
    double ra_mean = 0;
    if (ra_row >= 0 && ra_row < rows) {                          // every block recomputes the row mean (cheap, deterministic)
        double s = 0;
        for (int x = 0; x < N; ++x) s += U[(size_t)ra_row * N + x];
        ra_mean = s / (double)N;
    }
    double v[4] = {0, 0, 0, 0};
    const size_t total = (size_t)rows * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / N), x = (int)(i % N);
        const int gy_ = row_base + y;
        const double c = U[i];
        double gy, gx;
        if (gy_ == 0) gy = U[i + N] - c;
        else if (gy_ == N - 1) gy = c - U[i - N];
        else gy = 0.5 * (U[i + N] - U[i - N]);
        if (x == 0) gx = U[i + 1] - c;
        else if (x == N - 1) gx = c - U[i - 1];
        else gx = 0.5 * (U[i + 1] - U[i - 1]);
        double f, mu;
        thermo<1>(c, p, logtab, f, mu);
        v[0] += gy * gy + gx * gx;
        v[1] += f;
        v[2] += fabs(c - mean_u);
        if (y == ra_row) v[3] += fabs(c - ra_mean);
    }
    for (int k = 0; k < 4; ++k) red[k * blockDim.x + threadIdx.x] = v[k];
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0;
        for (unsigned j = 0; j < blockDim.x; ++j) s += red[threadIdx.x * blockDim.x + j];
        part[threadIdx.x * gridDim.x + blockIdx.x] = s;
    }
}

// prepare: vec from the k_slab_prepare block sums
CHS_KERNEL void k_slab_reduce_prepare(const double* part, int nblk, int N, double* vec) {
    if (threadIdx.x != 0) return;
    double s[4] = {0, 0, 0, 0};
    for (int k = 0; k < 4; ++k)
        for (int i = 0; i < nblk; ++i) s[k] += part[k * nblk + i];
    for (int v = 0; v < R_NVAL; ++v) vec[v] = 0;
    vec[R_GE] = s[0]; vec[R_F] = s[1]; vec[R_ABS] = s[2]; vec[R_RA] = s[3] / (double)N;
}

// vec[v] = sum over tiles of part[v][*] (+ the update kernel's GE blocks, + y-edge); one group
// of 32 threads per value, lane-strided partial sums added in fixed order (deterministic)
CHS_KERNEL void k_slab_reduce(const double* part, int ntiles, const double* part_ge, int nge, const double* yedge,
                              double* vec) {
    CHS_SMEM_DECL
    double* red = reinterpret_cast<double*>(CHS_SMEM_PTR);      // R_NVAL * 32
    const int v = threadIdx.x / 32, lane = threadIdx.x % 32;
    double s = 0;
    if (v < R_NVAL) {
        for (int i = lane; i < ntiles; i += 32) s += part[v * ntiles + i];
        if (v == R_GE) for (int i = lane; i < nge; i += 32) s += part_ge[i];
        red[v * 32 + lane] = s;
    }
    __syncthreads();
    if (v < R_NVAL && lane == 0) {
        double t = 0;
        for (int j = 0; j < 32; ++j) t += red[v * 32 + j];
        if (v == R_EDGE) t += yedge[0];
        vec[v] = t;
    }
}

// k_slab_reduce + both k_slab_yedge calls of a step in ONE launch of 1024 threads: 7 groups of 128
// lanes sum the per-tile partials (lane-strided, fixed order), the last 128 threads the y-edge
// terms of the stored field (top: rows 0/1 of the domain, bottom: rows N-2/N-1; null = not mine).
CHS_KERNEL void k_slab_sums(const double* part, int ntiles, const double* part_ge, int nge, const double* top0,
                            const double* top1, const double* bot0, const double* bot1, int N, double* vec,
                            PeerPtrs peers, int P) {
    CHS_PDL_TRIGGER();
    CHS_PDL_WAIT();
    CHS_SMEM_DECL
    double* red = reinterpret_cast<double*>(CHS_SMEM_PTR);      // (R_NVAL + 1) * 128
    const int v = threadIdx.x / 128, lane = threadIdx.x % 128;
    double s = 0;
    if (v < R_NVAL) {
        for (int i = lane; i < ntiles; i += 128) s += part[v * ntiles + i];
        if (v == R_GE) for (int i = lane; i < nge; i += 128) s += part_ge[i];
    } else {
        if (top0) for (int x = lane; x < N; x += 128) { const double d = top1[x] - top0[x]; s += d * d; }
        if (bot0) for (int x = lane; x < N; x += 128) { const double d = bot1[x] - bot0[x]; s += d * d; }
    }
    red[v * 128 + lane] = s;
    __syncthreads();
    if (v < R_NVAL && lane == 0) {
        double t = 0;
        for (int j = 0; j < 128; ++j) t += red[v * 128 + j];
        if (v == R_EDGE) {
            double e = 0;
            for (int j = 0; j < 128; ++j) e += red[R_NVAL * 128 + j];
            t += 0.75 * e;
        }
        vec[v] = t;
        // P > 0: this rank's sums also go to slot [rank] of every rank's gather buffer (peer-mapped stores);
        // k_slab_control adds the P slots in rank order after the exchange barrier -- no all-reduce launch
        for (int r = 0; r < P; ++r) peers.p[r][v] = t;
    }
}

// step_control() of the tile path, fed with the rank-reduced sums (one thread; every rank runs
// it on identical inputs and so keeps an identical Sim image).  post = 0: prologue.
CHS_KERNEL void k_slab_control(Sim* S, double* vec, double* rows, long long rows_cap, int N, int last, int post,
                               const double* allvec, int P, const double* colsum) {
    CHS_PDL_TRIGGER();
    CHS_PDL_WAIT();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (S->halted && post != 2) return;          // stopped: no further rows, no time accounting
    if (allvec) {                                // sums of all ranks, gathered by k_slab_sums: fixed order -> identical on every rank
        for (int v = 0; v < R_NVAL; ++v) {
            double t = 0;
            for (int r = 0; r < P; ++r) t += allvec[r * 8 + v];
            vec[v] = t;
        }
    }
    const chs_params& p = S->p;
    if (post == 2) {                     // Solver.prepare(): row 0 (solver.py:117-135)
        const double N2 = (double)N * (double)N, L2sq = p.L * p.L;
        const double E2 = 0.5 * p.Amr * p.kappa_tilde * L2sq * ((vec[R_GE] / (p.delx * p.delx)) / N2);
        const double E = p.Amr * L2sq * (vec[R_F] / N2) + E2;
        const double PS = vec[R_ABS] / N2;
        S->ra = vec[R_RA];
        double* r = rows;
        r[CHS_COL_IT] = 0; r[CHS_COL_E] = E; r[CHS_COL_E2] = E2; r[CHS_COL_SA] = 0; r[CHS_COL_DOMTIME] = 0;
        r[CHS_COL_RA] = S->ra; r[CHS_COL_L2] = 0; r[CHS_COL_PS] = PS; r[CHS_COL_DELT] = S->delt;
        S->rows_written = 1;
        S->e2_first = E2; S->e2_prev = E2;
        S->tau0 = 0; S->t0 = 0;
        S->stop_reason = ((E != E) || (E2 != E2) || (PS != PS)) ? CHS_STOP_NAN : CHS_STOP_NONE;
        S->computed_steps = 1;
        S->halted = 0;
        return;
    }
    if (post) {
        const double N2 = (double)N * (double)N, L2sq = p.L * p.L;
        const double grad2 = (vec[R_GE] + vec[R_EDGE]) / (p.delx * p.delx);
        const double E2 = 0.5 * p.Amr * p.kappa_tilde * L2sq * (grad2 / N2);
        const double E = p.Amr * L2sq * (vec[R_F] / N2) + E2;
        const double PS = vec[R_ABS] / N2, L2 = sqrt(S->mu2_pending) / N2, SA = vec[R_CNT] / N2;
        const double domtime = pow(S->time_passed, 1.0 / 3.0);
        S->mu2_pending = vec[R_MU2];
        S->ra = vec[R_RA];
        const long long rw = S->rows_written;
        if (rw < rows_cap) {
            double* r = rows + rw * CHS_NCOLS;
            r[CHS_COL_IT] = (double)S->computed_steps;
            r[CHS_COL_E] = E; r[CHS_COL_E2] = E2; r[CHS_COL_SA] = SA; r[CHS_COL_DOMTIME] = domtime;
            r[CHS_COL_RA] = S->ra; r[CHS_COL_L2] = L2; r[CHS_COL_PS] = PS; r[CHS_COL_DELT] = S->delt;
        }
        S->rows_written = rw + 1;
        S->u_stale = 0;
        if ((E != E) || (E2 != E2) || (SA != SA) || (domtime != domtime) || (S->ra != S->ra) || (L2 != L2) || (PS != PS)) {
            S->stop_reason = CHS_STOP_NAN;
            S->halted = 1;
            return;
        }
        S->computed_steps += 1;
        const bool falls = (S->e2_prev > E2) && (E2 > S->e2_first);
        S->e2_prev = E2;
        if (!S->skip_check && falls) {
            S->tau0 = (double)S->computed_steps;
            S->t0 = S->time_passed;
            if (!p.full_sim) { S->stop_reason = CHS_STOP_ENERGY; S->halted = 1; return; }
            S->skip_check = 1;
        }
    } else {
        S->mu2_pending = vec[R_MU2];
    }
    if (last) return;
    // ---- "pre" part of the next iteration: adaptive dt (solver.py:177-193) from the all-rank column sums
    const long long cs_next = S->computed_steps;
    if (colsum && p.adaptive_time && cs_next > 500 && (cs_next % 2) == 0) {
        double m = colsum[0];
        for (int x = 1; x < N; ++x) {
            const double v = colsum[x];
            if (v != v) m = v; else if (m == m && v < m) m = v;      // NaN propagates like np.min
        }
        const double dnew = (m > p.delt) ? m : p.delt;               // Python max(params.delt, dyn): NaN loses
        if (dnew / S->delt > 1.15) S->delt = 0.75 * S->delt + 0.25 * dnew;
        else S->delt = dnew;
        S->delt_coef = S->delt;
        sim_derive_lam(*S);
    }
    S->time_delta_sum += S->delt;
    S->time_passed = S->time_delta_sum / p.M_tilde;
    if (p.time_limit_s > 0.0 && S->time_passed > p.time_limit_s) { S->stop_reason = CHS_STOP_TIME; S->halted = 1; }
}

}  // namespace CHS_NS
