// DCT-as-GEMM variant for small and arbitrary N (e.g. the reference's `benchmark.py -N 100`
// smoke size, and N below the FFT path's minimum): one CTA owns one whole simulation, the
// field lives in shared memory for ALL time steps of a launch, and the orthonormal 2-D
// DCT-II / DCT-III (scipy.fftpack.dctn/idctn at reference solver.py:159,201,208) are the
// matrix products  C.X.C^T  and  C^T.Y.C  on the FP64 tensor cores
// (mma.sync.aligned.m8n8k4 f64, "DMMA"; C[k][n] = f_k cos(pi k (2n+1) / 2N)).
// Because nothing leaves the SM, the reference loop (solver.py:165-249) runs in its own
// order inside the kernel: mu -> [adaptive dt] -> time accounting / limit -> update ->
// [jitter] -> diagnostics (np.gradient stencils straight from shared memory) -> TimeData
// row -> stop test, for n iterations per launch with no host involvement.
//
// Cost: 4 GEMMs = 8 N^3 flop per step (N=100: 8 MFLOP) against ~0.7 MFLOP for the FFT path
// scaled to that size, but no kernel launches, barriers only inside one CTA and zero HBM
// traffic per step; selected per N by measurement (chsimpy_b200/solver.py: GEMM_MAX_N).
#pragma once
#include "chs_kernels.cuh"

namespace CHS_NS {

constexpr int GEMM_NT = 256;          // 8 warps
constexpr int GEMM_MAX_N = 104;       // two N8 x LD fp64 matrices must fit in 227 KB of shared memory
constexpr int GEMM_MIN_N = 8;

struct GemmArgs {
    Sim* sims;
    const int* sim_index;
    double* U;                  // [batch][N][N]
    double* hatU;               // [batch][N][N] natural order
    double* rows;
    long long rows_cap;
    const double* Cm;           // [N8][N8] DCT-II matrix, zero padded
    const double* Ct;           // its transpose
    const double* lam;
    const double2* logtab;
    const double* noise;        // [n_iters][N][N] or null
    const double* mean_host;    // prepare
    int N, N8, LD;
    const double* src;          // modes 3/4: stand-alone transforms
    double* dst;
    int mode;                   // 0 prepare, 1 begin (hat_U = dctn(U)), 2 steps, 3 dctn, 4 idctn
    long long n_iters;
};

#ifndef CHS_EMU
CHS_DEV void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif

// D = A . B for N8 x N8 matrices; every warp produces 8 x 32 output strips.
// A_GLOBAL: A is read from global memory (row pitch N8), B from shared (pitch LD); else the
// other way round.  Output always to shared memory `D` (pitch LD).
template <bool A_GLOBAL>
CHS_DEV void gemm_n8(const double* A, const double* B, double* D, int N8, int LD, int tid) {
    const int warp = tid >> 5, lane = tid & 31, nw = GEMM_NT / 32;
    const int lda = A_GLOBAL ? N8 : LD, ldb = A_GLOBAL ? LD : N8;
    const int tr = N8 / 8, tc = (N8 + 31) / 32;
    for (int tile = warp; tile < tr * tc; tile += nw) {
        const int r0 = (tile / tc) * 8, c0 = (tile % tc) * 32;
        double d[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#ifdef CHS_EMU
        const int r = r0 + lane / 4;
        for (int j = 0; j < 4; ++j)
            for (int e = 0; e < 2; ++e) {
                const int c = c0 + 8 * j + 2 * (lane % 4) + e;
                if (c >= N8) continue;
                double s = 0;
                for (int k = 0; k < N8; ++k) s += A[r * lda + k] * B[k * ldb + c];
                d[j][e] = s;
            }
#else
        const double* ap = A + (r0 + lane / 4) * lda + (lane % 4);
        const double* bp = B + (lane % 4) * ldb + c0 + lane / 4;
        for (int k0 = 0; k0 < N8; k0 += 4) {
            const double a = A_GLOBAL ? __ldg(ap + k0) : ap[k0];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c0 + 8 * j < N8) {                               // warp-uniform
                    const double b = A_GLOBAL ? bp[k0 * ldb + 8 * j] : __ldg(bp + k0 * ldb + 8 * j);
                    dmma884(d[j][0], d[j][1], a, b);
                }
            }
        }
#endif
        const int r_ = r0 + lane / 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 8 * j + 2 * (lane % 4);
            if (c < N8) { D[r_ * LD + c] = d[j][0]; D[r_ * LD + c + 1] = d[j][1]; }
        }
    }
}

// block sum of NV values (fixed order), result in every thread
template <int NV>
CHS_DEV void gemm_block_sum(double (&v)[NV], double* red, int tid) {
    __syncthreads();
    for (int i = 0; i < NV; ++i) red[i * GEMM_NT + tid] = v[i];
    __syncthreads();
    if (tid < NV) {
        double s = 0;
        for (int j = 0; j < GEMM_NT; ++j) s += red[tid * GEMM_NT + j];
        red[NV * GEMM_NT + tid] = s;
    }
    __syncthreads();
    for (int i = 0; i < NV; ++i) v[i] = red[NV * GEMM_NT + i];
}

// diagnostics of the field X (np.gradient stencils, free energy, PS, SA, Ra): solver.py:213-228
CHS_DEV void gemm_diag(const double* X, int N, int LD, const chs_params& p, const double2* ltab, double* red, int tid,
                       double& E, double& E2, double& PS, double& SA, double& Ra) {
    double s[6] = {0, 0, 0, 0, 0, 0};            // sum U, grad2 raw, F, cnt, |.-mean| (2nd pass), ra
    for (int i = tid; i < N * N; i += GEMM_NT) s[0] += X[(i / N) * LD + (i % N)];
    double m1[1] = {s[0]};
    gemm_block_sum<1>(m1, red, tid);
    const double mean = m1[0] / ((double)N * (double)N);
    const int rr = N / 2 + 1;
    double rs[1] = {0};
    for (int x = tid; x < N; x += GEMM_NT) rs[0] += X[rr * LD + x];
    gemm_block_sum<1>(rs, red, tid);
    const double rmean = rs[0] / (double)N;
    double v[5] = {0, 0, 0, 0, 0};
    for (int i = tid; i < N * N; i += GEMM_NT) {
        const int y = i / N, x = i % N;
        const double c = X[y * LD + x];
        double gy, gx;
        if (y == 0) gy = X[(y + 1) * LD + x] - c;
        else if (y == N - 1) gy = c - X[(y - 1) * LD + x];
        else gy = 0.5 * (X[(y + 1) * LD + x] - X[(y - 1) * LD + x]);
        if (x == 0) gx = X[y * LD + 1] - c;
        else if (x == N - 1) gx = c - X[y * LD + x - 1];
        else gx = 0.5 * (X[y * LD + x + 1] - X[y * LD + x - 1]);
        double f, mu;
        thermo(c, p, ltab, f, mu);
        v[0] += gy * gy + gx * gx;
        v[1] += f;
        v[2] += (c < p.threshold) ? 1.0 : 0.0;
        v[3] += fabs(c - mean);
        if (y == rr) v[4] += fabs(c - rmean);
    }
    gemm_block_sum<5>(v, red, tid);
    const double N2 = (double)N * (double)N, L2sq = p.L * p.L;
    E2 = 0.5 * p.Amr * p.kappa_tilde * L2sq * ((v[0] / (p.delx * p.delx)) / N2);
    E = p.Amr * L2sq * (v[1] / N2) + E2;
    SA = v[2] / N2;
    PS = v[3] / N2;
    Ra = v[4] / (double)N;
}

CHS_KERNEL void __launch_bounds__(GEMM_NT, 1) k_gemm(GemmArgs a) {
    CHS_SMEM_DECL
    double* sm = reinterpret_cast<double*>(CHS_SMEM_PTR);
    const int N = a.N, N8 = a.N8, LD = a.LD, tid = threadIdx.x;
    double* X = sm;                               // N8 x LD
    double* Y = sm + N8 * LD;
    double* red = Y + N8 * LD;                    // 8*GEMM_NT
    double2* ltab = reinterpret_cast<double2*>(red + 8 * GEMM_NT);
    const int sim = a.sim_index ? a.sim_index[blockIdx.x] : (int)blockIdx.x;
    Sim* S = a.sims + sim;
    if (a.mode >= 3) {                            // stand-alone dctn / idctn of src -> dst
        const double* in = a.src + (size_t)sim * N * N;
        double* out = a.dst + (size_t)sim * N * N;
        for (int i = tid; i < 2 * N8 * LD; i += GEMM_NT) sm[i] = 0.0;
        __syncthreads();
        for (int i = tid; i < N * N; i += GEMM_NT) X[(i / N) * LD + (i % N)] = in[i];
        __syncthreads();
        if (a.mode == 3) { gemm_n8<true>(a.Cm, X, Y, N8, LD, tid); __syncthreads(); gemm_n8<false>(Y, a.Ct, X, N8, LD, tid); }
        else { gemm_n8<true>(a.Ct, X, Y, N8, LD, tid); __syncthreads(); gemm_n8<false>(Y, a.Cm, X, N8, LD, tid); }
        __syncthreads();
        for (int i = tid; i < N * N; i += GEMM_NT) out[i] = X[(i / N) * LD + (i % N)];
        return;
    }
    if (a.mode == 2 && S->halted) return;
    const chs_params p = S->p;
    double* Ug = a.U + (size_t)sim * N * N;
    double* Hg = a.hatU + (size_t)sim * N * N;
    double* rows = a.rows + (size_t)sim * a.rows_cap * CHS_NCOLS;
    for (int i = tid; i < LOG_TABLE_N; i += GEMM_NT) ltab[i] = a.logtab[i];
    for (int i = tid; i < 2 * N8 * LD; i += GEMM_NT) sm[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < N * N; i += GEMM_NT) X[(i / N) * LD + (i % N)] = Ug[i];
    __syncthreads();
    if (a.mode == 0) {                            // Solver.prepare(): row 0 (solver.py:100-135)
        double E, E2, PS, SA, Ra;
        gemm_diag(X, N, LD, p, ltab, red, tid, E, E2, PS, SA, Ra);
        if (tid == 0) {
            double* r = rows;
            r[CHS_COL_IT] = 0; r[CHS_COL_E] = E; r[CHS_COL_E2] = E2; r[CHS_COL_SA] = 0; r[CHS_COL_DOMTIME] = 0;
            r[CHS_COL_RA] = Ra; r[CHS_COL_L2] = 0; r[CHS_COL_PS] = PS; r[CHS_COL_DELT] = S->delt;
            S->rows_written = 1; S->e2_first = E2; S->e2_prev = E2; S->tau0 = 0; S->t0 = 0; S->ra = Ra;
            S->stop_reason = ((E != E) || (E2 != E2) || (PS != PS) || (Ra != Ra)) ? CHS_STOP_NAN : CHS_STOP_NONE;
            S->computed_steps = 1; S->halted = 0; S->u_stale = 0;
        }
        return;
    }
    if (a.mode == 1) {                            // hat_U = dctn(U)  (solver.py:159)
        gemm_n8<true>(a.Cm, X, Y, N8, LD, tid);   // Y = C . U
        __syncthreads();
        gemm_n8<false>(Y, a.Ct, X, N8, LD, tid);  // X = Y . C^T
        __syncthreads();
        for (int i = tid; i < N * N; i += GEMM_NT) Hg[i] = X[(i / N) * LD + (i % N)];
        if (tid == 0) { S->rows_written = 0; S->halted = 0; S->delt_coef = p.delt; }
        return;
    }
    // ---------------- time loop
    const double delx2 = p.delx * p.delx;
    for (long long it = 0; it < a.n_iters; ++it) {
        // X = U.  mu -> Y, ||mu||^2 (solver.py:166-175, 225)
        double q[1] = {0};
        for (int i = tid; i < N * N; i += GEMM_NT) {
            const int o = (i / N) * LD + (i % N);
            double f, mu;
            thermo(X[o], p, ltab, f, mu);
            Y[o] = mu;
            q[0] += mu * mu;
        }
        gemm_block_sum<1>(q, red, tid);
        const double mu2 = q[0];
        // adaptive dt (solver.py:177-193): min over columns of sum_rows delt_max/sqrt(1+62.5 mu^2)
        double delt = S->delt, delt_coef = S->delt_coef;
        const long long cs = S->computed_steps;
        if (p.adaptive_time && cs > 500 && (cs % 2) == 0) {
            double mn[1];
            double best = 1e300;
            bool nan = false;
            for (int x = tid; x < N; x += GEMM_NT) {
                double s = 0;
                for (int y = 0; y < N; ++y) { const double m = Y[y * LD + x]; s += p.delt_max / sqrt(1.0 + 62.5 * (m * m)); }
                if (s != s) nan = true;
                best = s < best ? s : best;
            }
            __syncthreads();
            red[tid] = nan ? NAN : best;
            __syncthreads();
            if (tid == 0) {
                double m = red[0];
                for (int j = 1; j < GEMM_NT; ++j) { const double v = red[j]; if (v != v) m = v; else if (m == m && v < m) m = v; }
                red[GEMM_NT] = m;
            }
            __syncthreads();
            mn[0] = red[GEMM_NT];
            const double dnew = (mn[0] > p.delt) ? mn[0] : p.delt;
            if (dnew / delt > 1.15) delt = 0.75 * delt + 0.25 * dnew; else delt = dnew;
            delt_coef = delt;
        }
        // time accounting / limit (solver.py:195-199)
        const double tds = S->time_delta_sum + delt;
        const double tpass = tds / p.M_tilde;
        __syncthreads();
        if (tid == 0) { S->delt = delt; S->delt_coef = delt_coef; S->time_delta_sum = tds; S->time_passed = tpass; }
        if (p.time_limit_s > 0.0 && tpass > p.time_limit_s) {
            if (tid == 0) { S->stop_reason = CHS_STOP_TIME; S->halted = 1; }
            break;
        }
        // hat_mu = C . mu . C^T
        gemm_n8<true>(a.Cm, Y, X, N8, LD, tid);   // X = C . mu
        __syncthreads();
        gemm_n8<false>(X, a.Ct, Y, N8, LD, tid);  // Y = hat_mu
        __syncthreads();
        // hat_U = (hat_U + Seig*hat_mu)/CHeig  (solver.py:201-206, utils.py:39-49)
        const double lam1 = delt_coef / delx2, lam2 = p.kappa_tilde * lam1 / delx2;
        for (int i = tid; i < N * N; i += GEMM_NT) {
            const int ky = i / N, kx = i % N, o = ky * LD + kx;
            const double leig = a.lam[ky] + a.lam[kx];
            const double Se = __dmul_rn(lam1, leig);
            const double CH = __dadd_rn(1.0, __dmul_rn(__dmul_rn(lam2, leig), leig));
            const double hu = __ddiv_rn(__dadd_rn(Hg[i], __dmul_rn(Se, Y[o])), CH);
            Hg[i] = hu;
            Y[o] = hu;
        }
        __syncthreads();
        // U = C^T . hat_U . C  (solver.py:208)
        gemm_n8<true>(a.Ct, Y, X, N8, LD, tid);   // X = C^T . hat
        __syncthreads();
        gemm_n8<false>(X, a.Cm, Y, N8, LD, tid);  // Y = U_new
        __syncthreads();
        // jitter (solver.py:210-211)
        if (a.noise != nullptr) {
            const double* nz = a.noise + (size_t)it * N * N;
            for (int i = tid; i < N * N; i += GEMM_NT) Y[(i / N) * LD + (i % N)] += p.jitter * (2.0 * nz[i] - 1.0);
            __syncthreads();
        }
        // U_new -> X (keeps the zero padding of X intact: only the N x N block is copied)
        for (int i = tid; i < N * N; i += GEMM_NT) { const int o = (i / N) * LD + (i % N); X[o] = Y[o]; }
        __syncthreads();
        double E, E2, PS, SA, Ra;
        gemm_diag(X, N, LD, p, ltab, red, tid, E, E2, PS, SA, Ra);
        int stop = 0;
        if (tid == 0) {
            const double N2 = (double)N * (double)N;
            const double L2 = sqrt(mu2) / N2;
            const double domtime = pow(tpass, 1.0 / 3.0);
            const long long rw = S->rows_written;
            if (rw < a.rows_cap) {
                double* r = rows + rw * CHS_NCOLS;
                r[CHS_COL_IT] = (double)S->computed_steps;
                r[CHS_COL_E] = E; r[CHS_COL_E2] = E2; r[CHS_COL_SA] = SA; r[CHS_COL_DOMTIME] = domtime;
                r[CHS_COL_RA] = Ra; r[CHS_COL_L2] = L2; r[CHS_COL_PS] = PS; r[CHS_COL_DELT] = delt;
            }
            S->rows_written = rw + 1;
            S->ra = Ra;
            if ((E != E) || (E2 != E2) || (SA != SA) || (domtime != domtime) || (Ra != Ra) || (L2 != L2) || (PS != PS) || (delt != delt)) {
                S->stop_reason = CHS_STOP_NAN; S->halted = 1; stop = 1;
            } else {
                S->computed_steps += 1;
                const bool falls = (S->e2_prev > E2) && (E2 > S->e2_first);
                S->e2_prev = E2;
                if (!S->skip_check && falls) {
                    S->tau0 = (double)S->computed_steps;
                    S->t0 = tpass;
                    if (!p.full_sim) { S->stop_reason = CHS_STOP_ENERGY; S->halted = 1; stop = 1; }
                    else S->skip_check = 1;
                }
            }
            red[0] = (double)stop;
        }
        __syncthreads();
        stop = (int)red[0];
        __syncthreads();
        if (stop) break;
    }
    // the field goes back to global memory at the end of the launch
    for (int i = tid; i < N * N; i += GEMM_NT) Ug[i] = X[(i / N) * LD + (i % N)];
}

}  // namespace CHS_NS
