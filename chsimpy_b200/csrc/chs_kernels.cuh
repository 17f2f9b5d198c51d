// Kernels of the B200 Cahn-Hilliard stepper.  One time step of the reference loop
// (chsimpy/solver.py:165-249) is two launches over the whole batch of simulations:
//
//   k_col<STEP> : per tile of 16 columns:  column DCT-II of T1 (= row-DCT of mu)
//                 -> spectral update  hat_U = (hat_U + Seig*hat_mu)/CHeig  (solver.py:201-206,
//                 multipliers generated on the fly from the 1-D lambda table, utils.py:34-49)
//                 -> column DCT-III -> T2, plus the y-part of the gradient energy
//                 (Parseval over x: sum_x (d_y U)^2 == sum_kx (d_y T2)^2).
//   k_row<STEP> : per tile of 16 rows:  row DCT-III of T2 -> U_new (solver.py:208)
//                 [+ jitter, :210-211] -> x-gradient, free energy, PS, SA, Ra (:213-228)
//                 -> chemical potential mu(U_new) of the NEXT step (:166-175)
//                 -> row DCT-II -> T1.   The last CTA of a simulation then runs
//                 step_control(): TimeData row (:231-240), NaN flag (timedata.py:10),
//                 energy stop test (:242-249, timedata.py:63) and the "pre" part of the
//                 next iteration: adaptive dt (:177-193), time accounting / limit (:195-199).
//
// U itself is not written per step (it is idctn(hat_U), materialised by chs_end) unless
// jitter is on.  Everything is IEEE float64.
#pragma once
#include "dct_core.cuh"
#include "fastlog.cuh"
#include "../../include/chs_b200.h"

namespace chs {

// per-CTA partial sums, reduced in fixed order by the last CTA (deterministic)
enum { P_GY2 = 0, P_GX2, P_F, P_ABS, P_MU2, P_CNT, P_SUMU, P_NSLOT };

// device-resident image of one simulation
struct Sim {
    chs_params p;
    double delt, delt_coef, time_delta_sum, time_passed, tau0, t0;
    double e2_first, e2_prev, mu2_pending, ra;
    long long computed_steps;
    long long rows_written;
    int skip_check, stop_reason, halted, u_stale;
    unsigned ticket;
    int pad_;
};

enum { COL_FWD = 0, COL_STEP = 1, COL_INV = 2 };
enum { ROW_FWD_U = 0, ROW_FWD_MU = 1, ROW_STEP = 2, ROW_INV = 3 };
enum { DIAG_PREPARE = 0, DIAG_JITTER = 1 };

struct KArgs {
    Sim* sims;
    const int* sim_index;        // blockIdx.y -> simulation (compacted list of running sims), or null
    double* U;                   // [batch][N][N]
    double* hatU;
    double* T;
    const double* src;           // stand-alone transforms: input  (else null)
    double* dst;                 //                         output
    double* rows;                // [batch][rows_cap][9]
    long long rows_cap;
    double* part;                // [batch][P_NSLOT][NTILES]
    double* colpart;             // [batch][NTILES][N]   adaptive-dt column sums
    const double2* tw;           // exp(-2 pi i m / M), m < M
    const double2* om;           // exp(-i pi m / (2N)), m < N
    const double* lam;           // 2 cos(pi k/(N-1)) - 2
    const double2* logtab;       // fast_log table {1/c, log c}
    const double* noise;         // [N][N] uniform draws of this step, or null
    const double* noise_mean;    // mean of that draw
    const double* mean_host;     // prepare: [batch] mean(U)
    int store_U;                 // row step: write U_new to the U buffer
    int last;                    // no "pre" part after this iteration
    int iter_in_call;            // index of the iteration inside this chs_steps call
};

#ifdef CHS_EMU
#define CHS_LDCG(p) (*(p))
#else
#define CHS_LDCG(p) __ldcg(p)
#endif

// ---------------------------------------------------------------------------------------
// Block reduction of NV doubles; result valid in thread 0.  GPU: warp shuffles then one
// value per warp through shared memory, summed in fixed order.
template <int NV>
CHS_DEV void block_reduce(double (&v)[NV], double* scratch, int tid, int nthreads) {
#ifdef CHS_EMU
    for (int i = 0; i < NV; ++i) scratch[tid * NV + i] = v[i];
    __syncthreads();
    if (tid == 0) {
        for (int i = 0; i < NV; ++i) {
            double s = 0;
            for (int j = 0; j < nthreads; ++j) s += scratch[j * NV + i];
            v[i] = s;
        }
    }
    __syncthreads();
#else
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        v[i] = x;
    }
    const int warp = tid >> 5, lane = tid & 31, nw = (nthreads + 31) >> 5;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) scratch[warp * NV + i] = v[i];
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0;
            for (int w = 0; w < nw; ++w) s += scratch[w * NV + i];
            v[i] = s;
        }
    }
    __syncthreads();
#endif
}

// scratch after the tile: [0,2) last-CTA flag | [2, 2+2*TPL) Ra partials | reduction scratch
// (the host emulation needs NT*NV doubles there; emu::launch() allocates that slack)
#define CHS_FLAG_PTR(G, sm) (reinterpret_cast<int*>((sm) + G::TILE_DOUBLES))
#define CHS_RA_SCRATCH(G, sm) ((sm) + G::TILE_DOUBLES + 2)
#define CHS_RED_SCRATCH(G, sm) ((sm) + G::TILE_DOUBLES + 2 + 2 * G::TPL + 2 * LOG_TABLE_N)
#define CHS_LOGTAB(G, sm) (reinterpret_cast<double2*>((sm) + G::TILE_DOUBLES + 2 + 2 * G::TPL))

// copies the 2 KB fast_log table into shared memory (visible after the next barrier)
template <class G>
CHS_DEV double2* stage_logtab(double* sm, const double2* __restrict__ g, int tid) {
    double2* t = CHS_LOGTAB(G, sm);
    for (int i = tid; i < LOG_TABLE_N; i += G::NT) t[i] = g[i];
    return t;
}

// ---------------------------------------------------------------------------------------
// thermodynamics of one value (solver.py:166-175 and :218-221)
CHS_DEV void thermo(double u, const chs_params& p, const double2* __restrict__ ltab, double& f, double& mu) {
    const double ui = 1.0 - u;
    const double lu = fast_log(u, ltab), li = fast_log(ui, ltab);
    const double d = ui - u;
    const double uui = u * ui;
    f = p.RT * (u * (lu - p.B) + ui * li) + (p.A0 + p.A1 * d) * uui;
    mu = p.RT * (lu - li) - p.BRT + (p.A0 + p.A1 * d) * d - 2.0 * p.A1 * uui;
}

// ---------------------------------------------------------------------------------------
// End-of-iteration control, executed by every thread of the LAST CTA of a simulation.
// part/colpart of all CTAs are visible (threadfence + ticket).  rows: this sim's table.
template <int N>
CHS_DEV void step_control(Sim* S, const double* part, const double* colpart, double* rows, long long rows_cap,
                          int last, bool post, double* scratch, int tid, int nthreads) {
    using G = Geo<N>;
    constexpr int NTILES = G::NTILES;
    // ---- adaptive column minimum (needs all threads) ----------------------------------
    // the test of solver.py:177-181 is made with the value computed_steps will have at the
    // start of the next iteration, i.e. after the increment below when post is true
    const long long cs_next = S->computed_steps + (post ? 1 : 0);
    const bool want_dyn = !last && S->p.adaptive_time && cs_next > 500 && (cs_next % 2) == 0;
    __syncthreads();                                  // thread 0 changes computed_steps below
    double dyn = 0.0;
    if (want_dyn) {                                   // block-uniform
        double mn = 1e300;
        for (int x = tid; x < N; x += nthreads) {
            double s = 0;
            for (int tl = 0; tl < NTILES; ++tl) s += CHS_LDCG(colpart + (size_t)tl * N + x);
            mn = s < mn ? s : mn;
            if (s != s) mn = s;                       // NaN propagates like np.min
        }
        scratch[tid] = mn;
        __syncthreads();
        if (tid == 0) {
            double m = scratch[0];
            for (int j = 1; j < nthreads; ++j) {
                const double v = scratch[j];
                if (v != v) m = v; else if (m == m && v < m) m = v;
            }
            dyn = m;
        }
        __syncthreads();
    }
    if (tid != 0) return;
    const chs_params& p = S->p;
    if (post) {
        double acc[P_NSLOT];
        for (int s = 0; s < P_NSLOT; ++s) {
            double a = 0;
            for (int tl = 0; tl < NTILES; ++tl) a += CHS_LDCG(part + s * NTILES + tl);
            acc[s] = a;
        }
        const double N2 = (double)N * (double)N;
        const double L2sq = p.L * p.L;
        const double E2 = 0.5 * p.Amr * p.kappa_tilde * L2sq * ((acc[P_GX2] + acc[P_GY2]) / N2);
        const double E = p.Amr * L2sq * (acc[P_F] / N2) + E2;
        const double PS = acc[P_ABS] / N2;
        const double L2 = sqrt(S->mu2_pending) / N2;
        const double SA = acc[P_CNT] / N2;
        const double domtime = pow(S->time_passed, 1.0 / 3.0);
        S->mu2_pending = acc[P_MU2];
        const long long rw = S->rows_written;
        if (rw < rows_cap) {
            double* r = rows + rw * CHS_NCOLS;
            r[CHS_COL_IT] = (double)S->computed_steps;
            r[CHS_COL_E] = E; r[CHS_COL_E2] = E2; r[CHS_COL_SA] = SA; r[CHS_COL_DOMTIME] = domtime;
            r[CHS_COL_RA] = S->ra; r[CHS_COL_L2] = L2; r[CHS_COL_PS] = PS; r[CHS_COL_DELT] = S->delt;
        }
        S->rows_written = rw + 1;
        S->u_stale = 1;
        const bool bad = (E != E) || (E2 != E2) || (SA != SA) || (domtime != domtime) || (S->ra != S->ra) ||
                         (L2 != L2) || (PS != PS) || (S->delt != S->delt);
        if (bad) {                                    // timedata.py:10 fires before the increment
            S->stop_reason = CHS_STOP_NAN;
            S->halted = 1;
            return;
        }
        S->computed_steps += 1;
        // timedata.py:63 with it = computed_steps-1: E2[it-1] > E2[it] > E2[0]
        const bool falls = (S->e2_prev > E2) && (E2 > S->e2_first);
        S->e2_prev = E2;
        if (!S->skip_check && falls) {
            S->tau0 = (double)S->computed_steps;
            S->t0 = S->time_passed;
            if (!p.full_sim) {
                S->stop_reason = CHS_STOP_ENERGY;
                S->halted = 1;
                return;
            }
            S->skip_check = 1;
        }
    }
    else {                                            // prologue (chs_begin): only ||mu||^2 is new
        double a2 = 0;
        for (int tl = 0; tl < NTILES; ++tl) a2 += CHS_LDCG(part + P_MU2 * NTILES + tl);
        S->mu2_pending = a2;
    }
    if (last) return;
    // ---- "pre" part of the next iteration --------------------------------------------
    if (want_dyn) {
        const double dnew = (dyn > p.delt) ? dyn : p.delt;       // Python max(params.delt, dyn): NaN loses
        if (dnew / S->delt > 1.15) S->delt = 0.75 * S->delt + 0.25 * dnew;
        else S->delt = dnew;
        S->delt_coef = S->delt;
    }
    S->time_delta_sum += S->delt;
    S->time_passed = S->time_delta_sum / p.M_tilde;
    if (p.time_limit_s > 0.0 && S->time_passed > p.time_limit_s) {
        S->stop_reason = CHS_STOP_TIME;
        S->halted = 1;
    }
}

// ticket: returns true in every thread of the last CTA of this simulation
CHS_DEV bool last_cta(Sim* S, int ntiles, int* flag_smem, int tid) {
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&S->ticket, 1u);
        const int lastf = (prev == (unsigned)(ntiles - 1));
        if (lastf) S->ticket = 0;
        *flag_smem = lastf;
        __threadfence();
    }
    __syncthreads();
    return *flag_smem != 0;
}


// ---------------------------------------------------------------------------------------
// Tile I/O.  Loads are issued in independent batches of 8 per thread before the first use
// (memory-level parallelism), then scattered into the tile.
//   column tile: all N rows x LINES adjacent columns of a row-major N x N array
//   row tile   : LINES adjacent rows
// PHYS = the array index along the line is a physical coordinate (Makhoul position in smem),
// otherwise it is a spectral index stored at its natural position.
template <int N, bool PHYS>
CHS_DEV int line_pos(int n) { return PHYS ? mk_pos<N>(n) : n; }

template <int N, bool PHYS>
CHS_DEV void col_tile_load(double* sm, const double* __restrict__ g, int l, int t) {
    using G = Geo<N>;
    constexpr int CNT = N / G::TPL, UNR = 8;
#pragma unroll
    for (int j0 = 0; j0 < CNT; j0 += UNR) {
        double v[UNR];
#pragma unroll
        for (int j = 0; j < UNR; ++j) v[j] = g[(size_t)(t + (j0 + j) * G::TPL) * N + l];
#pragma unroll
        for (int j = 0; j < UNR; ++j) sm[line_pos<N, PHYS>(t + (j0 + j) * G::TPL) * G::LP + l] = v[j];
    }
}

template <int N, bool PHYS>
CHS_DEV void col_tile_store(const double* sm, double* __restrict__ g, int l, int t) {
    using G = Geo<N>;
    constexpr int CNT = N / G::TPL;
#pragma unroll 8
    for (int j = 0; j < CNT; ++j)
        g[(size_t)(t + j * G::TPL) * N + l] = sm[line_pos<N, PHYS>(t + j * G::TPL) * G::LP + l];
}

template <int N, bool PHYS>
CHS_DEV void row_tile_load(double* sm, const double* __restrict__ g, int tid) {
    using G = Geo<N>;
    constexpr int CNT = G::LINES * N / G::NT, UNR = 8;
#pragma unroll
    for (int j0 = 0; j0 < CNT; j0 += UNR) {
        double v[UNR];
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            const int i = tid + (j0 + j) * G::NT;
            v[j] = g[(size_t)(i / N) * N + (i % N)];
        }
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            const int i = tid + (j0 + j) * G::NT;
            sm[line_pos<N, PHYS>(i % N) * G::LP + (i / N)] = v[j];
        }
    }
}

template <int N, bool PHYS>
CHS_DEV void row_tile_store(const double* sm, double* __restrict__ g, int tid) {
    using G = Geo<N>;
    constexpr int CNT = G::LINES * N / G::NT;
#pragma unroll 8
    for (int j = 0; j < CNT; ++j) {
        const int i = tid + j * G::NT;
        g[(size_t)(i / N) * N + (i % N)] = sm[line_pos<N, PHYS>(i % N) * G::LP + (i / N)];
    }
}

// =======================================================================================
//  column kernel
// =======================================================================================
template <int N, int MODE>
CHS_KERNEL void __launch_bounds__(Geo<N>::NT, Geo<N>::MINB) k_col(KArgs a) {
    using G = Geo<N>;
    constexpr int M = G::M, LP = G::LP, LINES = G::LINES, TPL = G::TPL, NT = G::NT;
    CHS_SMEM_DECL
    double* sm = reinterpret_cast<double*>(CHS_SMEM_PTR);
    const int tid = threadIdx.x, l = tid % LINES, t = tid / LINES;
    const int sim = a.sim_index ? a.sim_index[blockIdx.y] : (int)blockIdx.y;
    const int tile = blockIdx.x, kx0 = tile * LINES;
    Sim* S = a.sims + sim;
    if (MODE == COL_STEP && S->halted) return;
    const size_t off = (size_t)sim * N * N;
    double* sl = sm + l;

    // -------- load + forward column DCT-II
    if (MODE != COL_INV) {
        col_tile_load<N, true>(sm, a.T + off + kx0, l, t);
        __syncthreads();
        fft_fwd<N>(sl, t, a.tw);
    }
    // -------- spectral middle section (registers <-> global, lanes along kx)
    {
        double* hat = (MODE == COL_FWD && a.dst) ? a.dst + off : a.hatU + off;
        const double* hat_in = (MODE == COL_INV && a.src) ? a.src + off : a.hatU + off;
        double lam1 = 0, lam2 = 0, lamx = 0;
        if (MODE == COL_STEP) {
            const double delx2 = S->p.delx * S->p.delx;
            lam1 = S->delt_coef / delx2;                  // utils.py:41-42
            lam2 = S->p.kappa_tilde * lam1 / delx2;
            lamx = a.lam[kx0 + l];
        }
#pragma unroll
        for (int k0 = 0; k0 < M / 2; k0 += TPL) {
            const int k = k0 + t;
            if (M / 2 < TPL && k >= M / 2) break;
            int idx[4];
            item_index<N>(k, idx);
            double c[4];
            if (MODE != COL_INV) post_item<N>(sl, k, a.om, c);
            if (MODE == COL_FWD) {
#pragma unroll
                for (int j = 0; j < 4; ++j) hat[(size_t)idx[j] * N + kx0 + l] = c[j];
            } else if (MODE == COL_STEP) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const size_t g = (size_t)idx[j] * N + kx0 + l;
                    const double leig = a.lam[idx[j]] + lamx;
                    const double Se = __dmul_rn(lam1, leig);
                    const double CH = __dadd_rn(1.0, __dmul_rn(__dmul_rn(lam2, leig), leig));
                    const double hu = __ddiv_rn(__dadd_rn(hat[g], __dmul_rn(Se, c[j])), CH);
                    hat[g] = hu;
                    c[j] = hu;
                }
                pre_item<N>(sl, k, a.om, c);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) c[j] = hat_in[(size_t)idx[j] * N + kx0 + l];
                pre_item<N>(sl, k, a.om, c);
            }
        }
    }
    if (MODE == COL_FWD) return;
    __syncthreads();
    // -------- inverse column DCT-III
    fft_inv<N>(sl, t, a.tw);
    // -------- y-part of the gradient energy on T2 (np.gradient along axis 0, solver.py:213)
    if (MODE == COL_STEP) {
        constexpr int YPT = N / TPL;
        const int y0 = t * YPT;
        const double ih = 1.0 / S->p.delx, ih2 = 0.5 / S->p.delx;
        double prev = (y0 > 0) ? sl[mk_pos<N>(y0 - 1) * LP] : 0.0;
        double cur = sl[mk_pos<N>(y0) * LP];
        double acc = 0;
#pragma unroll 4
        for (int y = y0; y < y0 + YPT; ++y) {
            const double nxt = (y + 1 < N) ? sl[mk_pos<N>(y + 1) * LP] : 0.0;
            double g;
            if (y == 0) g = (nxt - cur) * ih;
            else if (y == N - 1) g = (cur - prev) * ih;
            else g = (nxt - prev) * ih2;
            acc += g * g;
            prev = cur; cur = nxt;
        }
        double v[1] = {acc};
        block_reduce<1>(v, CHS_RED_SCRATCH(G, sm), tid, NT);
        if (tid == 0) a.part[((size_t)sim * P_NSLOT + P_GY2) * G::NTILES + tile] = v[0];
    }
    // -------- store T2 tile
    col_tile_store<N, true>(sm, a.T + off + kx0, l, t);
}

// =======================================================================================
//  row kernel
// =======================================================================================
template <int N, int MODE>
CHS_KERNEL void __launch_bounds__(Geo<N>::NT, Geo<N>::MINB) k_row(KArgs a) {
    using G = Geo<N>;
    constexpr int M = G::M, LP = G::LP, LINES = G::LINES, TPL = G::TPL, NT = G::NT;
    constexpr int IPT = (M / 2 + TPL - 1) / TPL;
    CHS_SMEM_DECL
    double* sm = reinterpret_cast<double*>(CHS_SMEM_PTR);
    double* ra_scr = CHS_RA_SCRATCH(G, sm);                // 2*TPL doubles
    int* flag = CHS_FLAG_PTR(G, sm);
    const int tid = threadIdx.x, l = tid % LINES, t = tid / LINES;
    const int sim = a.sim_index ? a.sim_index[blockIdx.y] : (int)blockIdx.y;
    const int tile = blockIdx.x, row0 = tile * LINES;
    Sim* S = a.sims + sim;
    if (MODE == ROW_STEP && S->halted) return;
    const size_t off = (size_t)sim * N * N;
    double* sl = sm + l;
    const bool control = (MODE == ROW_STEP) || (MODE == ROW_FWD_MU);
    const double2* ltab = control ? stage_logtab<G>(sm, a.logtab, tid) : nullptr;

    // ================= inverse half: T2 rows -> U rows (Makhoul order in smem)
    if (MODE == ROW_STEP || MODE == ROW_INV) {
        const double* src = (MODE == ROW_INV && a.src) ? a.src + off : a.T + off;
        row_tile_load<N, false>(sm, src + (size_t)row0 * N, tid);
        __syncthreads();
        double c[IPT][4];
#pragma unroll
        for (int it = 0; it < IPT; ++it) {
            const int k = it * TPL + t;
            if (k < M / 2) {
                int idx[4];
                item_index<N>(k, idx);
#pragma unroll
                for (int j = 0; j < 4; ++j) c[it][j] = sl[idx[j] * LP];
            }
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < IPT; ++it) {
            const int k = it * TPL + t;
            if (k < M / 2) pre_item<N>(sl, k, a.om, c[it]);
        }
        __syncthreads();
        fft_inv<N>(sl, t, a.tw);
    } else {
        // ROW_FWD_U / ROW_FWD_MU: physical U rows -> smem (Makhoul order)
        const double* src = (MODE == ROW_FWD_U && a.src) ? a.src + off : a.U + off;
        row_tile_load<N, true>(sm, src + (size_t)row0 * N, tid);
        __syncthreads();
    }

    // ================= jitter (solver.py:210-211): U += jitter*(2*noise - 1)
    if (MODE == ROW_STEP && a.noise != nullptr) {
        const double jit = S->p.jitter;
        for (int i = tid; i < LINES * N; i += NT) {
            const int l2 = i / N, x = i % N;
            const double r = a.noise[(size_t)(row0 + l2) * N + x];
            sm[mk_pos<N>(x) * LP + l2] += jit * (2.0 * r - 1.0);
        }
        __syncthreads();
    }

    // ================= physical-space output
    if (MODE == ROW_INV || (MODE == ROW_STEP && a.store_U)) {
        double* dstU = (MODE == ROW_INV && a.dst) ? a.dst + off : a.U + off;
        row_tile_store<N, true>(sm, dstU + (size_t)row0 * N, tid);
        if (MODE == ROW_INV) return;
    }

    // ================= diagnostics of U_new and chemical potential for the next step
    if (control) {
        constexpr int XPT = N / TPL;
        const int x0 = t * XPT;
        const chs_params p = S->p;
        const bool diag = (MODE == ROW_STEP);
        const int ra_row = N / 2 + 1;                                   // int(N/2)+1, solver.py:226
        const bool ra_line = diag && (row0 + l == ra_row);
        const bool ra_tile = diag && (ra_row >= row0) && (ra_row < row0 + LINES);
        // mean(U_new): hat_U[0,0]/N is conserved by the update (Q4); jitter shifts it
        double meanU = 0;
        if (diag) {
            meanU = a.hatU[off] / (double)N;
            if (a.noise != nullptr) meanU += p.jitter * (2.0 * a.noise_mean[0] - 1.0);
        }
        const double lo = (x0 > 0) ? sl[mk_pos<N>(x0 - 1) * LP] : 0.0;
        const double hi = (x0 + XPT < N) ? sl[mk_pos<N>(x0 + XPT) * LP] : 0.0;
        if (ra_tile) {
            if (ra_line) {
                double s = 0;
                for (int x = x0; x < x0 + XPT; ++x) s += sl[mk_pos<N>(x) * LP];
                ra_scr[t] = s;
            }
        }
        __syncthreads();
        double ra_mean = 0;
        if (ra_line) {
            double s = 0;
            for (int j = 0; j < TPL; ++j) s += ra_scr[j];
            ra_mean = s / (double)N;
        }
        const double ih = 1.0 / p.delx, ih2 = 0.5 / p.delx;
        double v[6] = {0, 0, 0, 0, 0, 0};           // GX2, F, ABS, MU2, CNT, SUMU
        double ra_abs = 0;
        double prev = lo, cur = sl[mk_pos<N>(x0) * LP];
#pragma unroll 4
        for (int x = x0; x < x0 + XPT; ++x) {
            const double nxt = (x + 1 < x0 + XPT) ? sl[mk_pos<N>(x + 1) * LP] : hi;
            double f, mu;
            thermo(cur, p, ltab, f, mu);
            if (diag) {
                double g;
                if (x == 0) g = (nxt - cur) * ih;
                else if (x == N - 1) g = (cur - prev) * ih;
                else g = (nxt - prev) * ih2;
                v[0] += g * g;
                v[1] += f;
                v[2] += fabs(cur - meanU);
                v[4] += (cur < p.threshold) ? 1.0 : 0.0;
                v[5] += cur;
                if (ra_line) ra_abs += fabs(cur - ra_mean);
            }
            v[3] += mu * mu;
            sl[mk_pos<N>(x) * LP] = mu;
            prev = cur; cur = nxt;
        }
        if (ra_line) ra_scr[TPL + t] = ra_abs;
        block_reduce<6>(v, CHS_RED_SCRATCH(G, sm), tid, NT);        // contains barriers
        if (tid == 0) {
            double* pp = a.part + (size_t)sim * P_NSLOT * G::NTILES + tile;
            pp[P_MU2 * G::NTILES] = v[3];
            if (diag) {
                pp[P_GX2 * G::NTILES] = v[0];
                pp[P_F * G::NTILES] = v[1];
                pp[P_ABS * G::NTILES] = v[2];
                pp[P_CNT * G::NTILES] = v[4];
                pp[P_SUMU * G::NTILES] = v[5];
            } else {
                pp[P_GX2 * G::NTILES] = 0; pp[P_F * G::NTILES] = 0; pp[P_ABS * G::NTILES] = 0;
                pp[P_CNT * G::NTILES] = 0; pp[P_SUMU * G::NTILES] = 0;
            }
            if (ra_tile) {
                double s = 0;
                for (int j = 0; j < TPL; ++j) s += ra_scr[TPL + j];
                S->ra = s / (double)N;
            }
        }
        // adaptive dt: column sums of delt_max/sqrt(1 + 62.5 mu^2) over this tile's rows (solver.py:182-183)
        const long long cs_next = S->computed_steps + (diag ? 1 : 0);
        if (p.adaptive_time && !a.last && cs_next > 500 && (cs_next % 2) == 0) {
            for (int x = tid; x < N; x += NT) {
                const double* col = sm + mk_pos<N>(x) * LP;
                double s = 0;
#pragma unroll
                for (int l2 = 0; l2 < LINES; ++l2) {
                    const double m = col[l2];
                    s += p.delt_max / sqrt(1.0 + 62.5 * (m * m));
                }
                a.colpart[((size_t)sim * G::NTILES + tile) * N + x] = s;
            }
            __syncthreads();                                       // mu is transformed in place next
        }
    }

    // ================= forward half: rows -> row DCT-II -> T
    fft_fwd<N>(sl, t, a.tw);
    {
        double c[IPT][4];
#pragma unroll
        for (int it = 0; it < IPT; ++it) {
            const int k = it * TPL + t;
            if (k < M / 2) post_item<N>(sl, k, a.om, c[it]);
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < IPT; ++it) {
            const int k = it * TPL + t;
            if (k < M / 2) {
                int idx[4];
                item_index<N>(k, idx);
#pragma unroll
                for (int j = 0; j < 4; ++j) sl[idx[j] * LP] = c[it][j];
            }
        }
        __syncthreads();
        double* dstT = (MODE == ROW_FWD_U && a.dst) ? a.dst + off : a.T + off;
        row_tile_store<N, false>(sm, dstT + (size_t)row0 * N, tid);
    }

    // ================= control (not in jitter mode: k_diag finishes the iteration there)
    if (control) {
        const bool defer = (MODE == ROW_STEP) && (a.noise != nullptr);
        if (!defer) {
            if (last_cta(S, G::NTILES, flag, tid)) {
                step_control<N>(S, a.part + (size_t)sim * P_NSLOT * G::NTILES,
                                a.colpart + (size_t)sim * G::NTILES * N,
                                a.rows + (size_t)sim * a.rows_cap * CHS_NCOLS, a.rows_cap, a.last,
                                MODE == ROW_STEP, sm, tid, NT);
            }
        }
    }
}

// =======================================================================================
//  k_diag: diagnostics straight from the U buffer.
//    DIAG_PREPARE : all of row 0 (Solver.prepare, solver.py:100-135)
//    DIAG_JITTER  : only the y-part of the gradient energy of the jittered U (Parseval
//                   does not hold once noise is added in physical space), then step_control
// =======================================================================================
template <int N, int MODE>
CHS_KERNEL void __launch_bounds__(Geo<N>::NT) k_diag(KArgs a) {
    using G = Geo<N>;
    constexpr int LINES = G::LINES, NT = G::NT;
    CHS_SMEM_DECL
    double* sm = reinterpret_cast<double*>(CHS_SMEM_PTR);
    int* flag = CHS_FLAG_PTR(G, sm);
    const int tid = threadIdx.x;
    const int sim = a.sim_index ? a.sim_index[blockIdx.y] : (int)blockIdx.y;
    const int tile = blockIdx.x, row0 = tile * LINES;
    Sim* S = a.sims + sim;
    if (MODE == DIAG_JITTER && S->halted) return;
    const chs_params p = S->p;
    const double* U = a.U + (size_t)sim * N * N;
    const double ih = 1.0 / p.delx, ih2 = 0.5 / p.delx;
    const double meanU = (MODE == DIAG_PREPARE) ? a.mean_host[sim] : 0.0;
    const double2* ltab = nullptr;
    if (MODE == DIAG_PREPARE) {
        ltab = stage_logtab<G>(sm, a.logtab, tid);
        __syncthreads();
    }
    double v[4] = {0, 0, 0, 0};                    // GY2, GX2, F, ABS
    for (int i = tid; i < LINES * N; i += NT) {
        const int y = row0 + i / N, x = i % N;
        const double c = U[(size_t)y * N + x];
        double gy;
        if (y == 0) gy = (U[(size_t)(y + 1) * N + x] - c) * ih;
        else if (y == N - 1) gy = (c - U[(size_t)(y - 1) * N + x]) * ih;
        else gy = (U[(size_t)(y + 1) * N + x] - U[(size_t)(y - 1) * N + x]) * ih2;
        v[0] += gy * gy;
        if (MODE == DIAG_PREPARE) {
            double gx;
            if (x == 0) gx = (U[(size_t)y * N + 1] - c) * ih;
            else if (x == N - 1) gx = (c - U[(size_t)y * N + x - 1]) * ih;
            else gx = (U[(size_t)y * N + x + 1] - U[(size_t)y * N + x - 1]) * ih2;
            double f, mu;
            thermo(c, p, ltab, f, mu);
            v[1] += gx * gx;
            v[2] += f;
            v[3] += fabs(c - meanU);
        }
    }
    block_reduce<4>(v, CHS_RED_SCRATCH(G, sm), tid, NT);
    double* pp = a.part + (size_t)sim * P_NSLOT * G::NTILES + tile;
    if (tid == 0) {
        pp[P_GY2 * G::NTILES] = v[0];
        if (MODE == DIAG_PREPARE) {
            pp[P_GX2 * G::NTILES] = v[1];
            pp[P_F * G::NTILES] = v[2];
            pp[P_ABS * G::NTILES] = v[3];
        }
    }
    // Ra of row int(N/2)+1 (solver.py:115-116)
    const int ra_row = N / 2 + 1;
    if (MODE == DIAG_PREPARE && ra_row >= row0 && ra_row < row0 + LINES) {
        double s[1] = {0};
        for (int x = tid; x < N; x += NT) s[0] += U[(size_t)ra_row * N + x];
        block_reduce<1>(s, CHS_RED_SCRATCH(G, sm), tid, NT);
        if (tid == 0) sm[0] = s[0] / (double)N;
        __syncthreads();
        const double m = sm[0];
        double q[1] = {0};
        for (int x = tid; x < N; x += NT) q[0] += fabs(U[(size_t)ra_row * N + x] - m);
        block_reduce<1>(q, CHS_RED_SCRATCH(G, sm), tid, NT);
        if (tid == 0) S->ra = q[0] / (double)N;
    }
    if (!last_cta(S, G::NTILES, flag, tid)) return;
    if (MODE == DIAG_JITTER) {
        step_control<N>(S, a.part + (size_t)sim * P_NSLOT * G::NTILES, a.colpart + (size_t)sim * G::NTILES * N,
                        a.rows + (size_t)sim * a.rows_cap * CHS_NCOLS, a.rows_cap, a.last, true,
                        sm, tid, NT);
        return;
    }
    if (tid != 0) return;
    // ---- Solver.prepare(): row 0 and state reset (solver.py:117-135)
    double acc[4] = {0, 0, 0, 0};
    const int slots[4] = {P_GY2, P_GX2, P_F, P_ABS};
    const double* part = a.part + (size_t)sim * P_NSLOT * G::NTILES;
    for (int s = 0; s < 4; ++s)
        for (int tl = 0; tl < G::NTILES; ++tl) acc[s] += CHS_LDCG(part + slots[s] * G::NTILES + tl);
    const double N2 = (double)N * (double)N, L2sq = p.L * p.L;
    const double E2 = 0.5 * p.Amr * p.kappa_tilde * L2sq * ((acc[0] + acc[1]) / N2);
    const double E = p.Amr * L2sq * (acc[2] / N2) + E2;
    const double PS = acc[3] / N2;
    double* r = a.rows + (size_t)sim * a.rows_cap * CHS_NCOLS;
    r[CHS_COL_IT] = 0; r[CHS_COL_E] = E; r[CHS_COL_E2] = E2; r[CHS_COL_SA] = 0; r[CHS_COL_DOMTIME] = 0;
    r[CHS_COL_RA] = S->ra; r[CHS_COL_L2] = 0; r[CHS_COL_PS] = PS; r[CHS_COL_DELT] = S->delt;
    S->rows_written = 1;
    S->e2_first = E2;
    S->e2_prev = E2;
    S->tau0 = 0; S->t0 = 0;
    S->stop_reason = ((E != E) || (E2 != E2) || (PS != PS) || (S->ra != S->ra)) ? CHS_STOP_NAN : CHS_STOP_NONE;
    S->computed_steps = 1;
    S->halted = 0;
    S->u_stale = 0;
}

// begin(): reset per-call state before the prologue kernels (one thread per sim)
CHS_KERNEL void k_begin(Sim* sims, int batch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    Sim* S = sims + i;
    S->rows_written = 0;
    S->halted = 0;
    S->delt_coef = S->p.delt;          // solver.py:151-152: multipliers of the *initial* delt
    S->ticket = 0;
}

// self-test of fast_log (chs_debug_log)
CHS_KERNEL void k_debug_log(const double* x, double* y, long long n, const double2* tab) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fast_log(x[i], tab);
}

CHS_KERNEL void k_rewind(Sim* sims, int batch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) sims[i].rows_written = 0;
}

}  // namespace chs
