// Arbitrary-N path (the reference accepts any N, chsimpy/cli_parser.py:27, solver.py:45-82): one simulation whose
// size is neither a power of two (FFT kernels) nor small enough for the one-CTA GEMM kernel (N <= 104).  The 2-D
// orthonormal DCT-II / DCT-III of solver.py:159,201,208 are the matrix products  C.X.C^T  /  C^T.Y.C  as tiled
// FP64 tensor-core GEMMs (mma.sync.m8n8k4.f64 -> DMMA) over global memory; everything else is elementwise:
//
//   T  = A . C^T ; Mh = C . T                 k_big_gemm x2      (A = mu of the previous step)
//   H  = (H + Seig*Mh)/CHeig                  k_big_update       (natural frequency order, utils.py:34-49)
//   T  = H . C   ; U  = C^T . T               k_big_gemm x2      (solver.py:208)
//   [U += jitter*(2 noise - 1)] ; F, |U-mean|, SA count, Ra, ||mu||^2 ; A = mu(U)      k_big_phys
//   gradient energy by np.gradient's stencils k_slab_grad ; [adaptive column sums k_slab_colsum]
//   k_slab_sums -> k_slab_control             (the slab path's reduction and control kernels, one rank)
//
// 8 N^3 flop per step instead of O(N^2 log N): a completeness path (N = 768: ~3.6 GFLOP per step), not the fast one.
#pragma once
#include "chs_slab.cuh"

namespace CHS_NS {

// D = A . B for n8 x n8 row-major matrices with pitch ld (zero padded to multiples of 8).  CTA = 256 threads = 8
// warps, 64 x 64 output tile: warp w owns rows 16 (w/2) .. +15 and columns 32 (w%2) .. +31 = 2 x 4 DMMA blocks
// (8 accumulator pairs).  The K loop runs in slabs of 16 through shared memory; the next slab is fetched into
// registers while the current one is multiplied (software double buffering).
constexpr int BIG_TILE = 64, BIG_KS = 16;
CHS_KERNEL void __launch_bounds__(256) k_big_gemm(const double* A, const double* B, double* D, int n8, int ld) {
    CHS_SMEM_DECL
    double* sA = reinterpret_cast<double*>(CHS_SMEM_PTR);            // [64][BIG_KS + 1]
    double* sB = sA + BIG_TILE * (BIG_KS + 1);                       // [BIG_KS][64 + 1]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.y * BIG_TILE, c0 = blockIdx.x * BIG_TILE;
    const int wr = (warp >> 1) * 16, wc = (warp & 1) * 32;
    double d[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) d[i][j][0] = d[i][j][1] = 0.0;
    // staging: thread -> 4 elements of the A slab (64 x 16) and 4 of the B slab (16 x 64)
    double pa[4], pb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = tid + e * 256;
            const int ar = i / BIG_KS, ac = i % BIG_KS, br = i / BIG_TILE, bc = i % BIG_TILE;
            pa[e] = (r0 + ar < n8 && k0 + ac < n8) ? A[(size_t)(r0 + ar) * ld + k0 + ac] : 0.0;
            pb[e] = (k0 + br < n8 && c0 + bc < n8) ? B[(size_t)(k0 + br) * ld + c0 + bc] : 0.0;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < n8; k0 += BIG_KS) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = tid + e * 256;
            sA[(i / BIG_KS) * (BIG_KS + 1) + (i % BIG_KS)] = pa[e];
            sB[(i / BIG_TILE) * (BIG_TILE + 1) + (i % BIG_TILE)] = pb[e];
        }
        __syncthreads();
        if (k0 + BIG_KS < n8) fetch(k0 + BIG_KS);                    // in flight during the multiplication below
#ifdef CHS_EMU
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 4; ++j)
                for (int e = 0; e < 2; ++e) {
                    const int r = wr + 8 * i + lane / 4, c = wc + 8 * j + 2 * (lane % 4) + e;
                    double acc = 0;
                    for (int k = 0; k < BIG_KS; ++k) acc += sA[r * (BIG_KS + 1) + k] * sB[k * (BIG_TILE + 1) + c];
                    d[i][j][e] += acc;
                }
#else
#pragma unroll
        for (int kk = 0; kk < BIG_KS; kk += 4) {
            double a[2], bb[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = sA[(wr + 8 * i + lane / 4) * (BIG_KS + 1) + kk + (lane % 4)];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = sB[(kk + (lane % 4)) * (BIG_TILE + 1) + wc + 8 * j + lane / 4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(d[i][j][0], d[i][j][1], a[i], bb[j]);
        }
#endif
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = r0 + wr + 8 * i + lane / 4, c = c0 + wc + 8 * j + 2 * (lane % 4);
            if (r < n8 && c < n8) { D[(size_t)r * ld + c] = d[i][j][0]; D[(size_t)r * ld + c + 1] = d[i][j][1]; }
        }
}

// H = (H + Seig*Mh)/CHeig, natural order (solver.py:201-206; multipliers from the 1-D table as everywhere)
CHS_KERNEL void k_big_update(double* H, const double* Mh, const double* lam, const Sim* S, int N, int ld) {
    if (S->halted) return;
    const double delx2 = S->p.delx * S->p.delx;
    const double lam1 = S->delt_coef / delx2, lam2 = S->p.kappa_tilde * lam1 / delx2;
    const size_t total = (size_t)N * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ky = (int)(i / N), kx = (int)(i % N);
        const double leig = __ldg(lam + ky) + __ldg(lam + kx);
        const double Se = __dmul_rn(lam1, leig);
        const double CH = __dadd_rn(1.0, __dmul_rn(__dmul_rn(lam2, leig), leig));
        const size_t o = (size_t)ky * ld + kx;
        H[o] = __ddiv_rn(__dadd_rn(H[o], __dmul_rn(Se, Mh[o])), CH);
    }
}

// copies between the N x N field (pitch N) and the padded GEMM operand (pitch ld, zero padding kept)
CHS_KERNEL void k_big_copy(const double* src, int sld, double* dst, int dld, int N, const Sim* S, int respect_halt) {
    if (respect_halt && S->halted) return;
    const size_t total = (size_t)N * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        dst[(i / N) * dld + (i % N)] = src[(i / N) * sld + (i % N)];
}

// per-element physics on the new field Up (padded, pitch ld): optional jitter, stored to U (pitch N); free energy,
// |U - mean|, SA count, Ra of row N/2+1, ||mu||^2 -> per-block partials part[v][gridDim.x]; mu -> A (padded)
CHS_KERNEL void k_big_phys(const double* Up, double* U, double* A, int N, int ld, Sim* S, const double2* logtab,
                           double mean_u, const double* noise, const double* noise_mean, int diag, int from_U, double* part) {
    CHS_SMEM_DECL
    double* red = reinterpret_cast<double*>(CHS_SMEM_PTR);          // 5 * blockDim.x
    if (diag && S->halted) return;
    const chs_params p = S->p;
    ThermoK k;
    k.RT = p.RT; k.mBRT = -p.BRT; k.A0 = p.A0; k.A1 = p.A1; k.m2A1 = -2.0 * p.A1; k.B = p.B;
    const double jv = noise ? p.jitter : 0.0;
    const double meanU = mean_u + (noise ? p.jitter * (2.0 * noise_mean[0] - 1.0) : 0.0);
    const int ra_row = N / 2 + 1;
    double ra_mean = 0;
    if (diag) {                                  // every block recomputes the (jittered) Ra row mean: cheap, deterministic
        double s = 0;
        for (int x = 0; x < N; ++x) {
            const size_t i = (size_t)ra_row * N + x;
            const double u = from_U ? U[i] : Up[(size_t)ra_row * ld + x];
            s += u + (noise ? jv * (2.0 * noise[i] - 1.0) : 0.0);
        }
        ra_mean = s / (double)N;
    }
    double fa = 0, fb = 0, fp = 0, ab = 0, mu2 = 0, ra = 0, cnt = 0;
    const size_t total = (size_t)N * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / N), x = (int)(i % N);
        double u = from_U ? U[i] : Up[(size_t)y * ld + x];
        if (noise) u += jv * (2.0 * noise[i] - 1.0);
        if (!from_U || noise) U[i] = u;
        const double mu = thermo_acc<1, true>(u, k, logtab, fa, fb, fp);
        A[(size_t)y * ld + x] = mu;
        mu2 = chs_fma(mu, mu, mu2);
        if (diag) {
            ab += fabs(u - meanU);
            cnt += (u < p.threshold) ? 1.0 : 0.0;
            if (y == ra_row) ra += fabs(u - ra_mean);
        }
    }
    const double v[5] = {chs_fma(k.RT, fa + fb, fp), ab, mu2, cnt, ra};
    for (int q = 0; q < 5; ++q) red[q * blockDim.x + threadIdx.x] = v[q];
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0;
        for (unsigned j = 0; j < blockDim.x; ++j) s += red[threadIdx.x * blockDim.x + j];
        const int slot[5] = {R_F, R_ABS, R_MU2, R_CNT, R_RA};
        part[slot[threadIdx.x] * gridDim.x + blockIdx.x] = (threadIdx.x == 4) ? s / (double)N : s;
    }
    if (threadIdx.x == 5) { part[R_GE * gridDim.x + blockIdx.x] = 0; part[R_EDGE * gridDim.x + blockIdx.x] = 0; }
}

}  // namespace CHS_NS
