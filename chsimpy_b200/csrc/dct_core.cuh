// Shared-memory radix-8/4/2 FFT building blocks for the orthonormal DCT-II / DCT-III
// (scipy.fftpack.dctn/idctn(norm='ortho'), called at reference chsimpy/solver.py:159,201,208).
//
// Algorithm (one "line" = one row or one column of N reals, M = N/2):
//   DCT-II :  Makhoul even/odd reorder v -> pack z[n] = v[2n] + i v[2n+1] -> M-point complex
//             FFT (in-place decimation-in-frequency, digit-reversed output) -> one combined
//             "post" pass (real-FFT untangle + quarter-wave twiddle) -> C[k].
//   DCT-III:  the exact transpose: "pre" pass -> in-place decimation-in-time inverse FFT
//             (digit-reversed input, natural output) -> un-reorder.
// Because the forward FFT leaves its output digit-reversed and the inverse FFT consumes
// digit-reversed input, no reordering pass is ever executed.
//
// Tile layout in shared memory: LINES lines are transformed together and the line index
// is the fastest-varying (lane) dimension:  element p of line l lives at  sm[p*LP + l].
// Every FFT access of a warp therefore touches LINES consecutive doubles -> bank-conflict
// free for any stride, and all lanes of a half-warp share their twiddle factors.
// tools/dct_model.py is the numpy model this file was derived from.
#pragma once
#include "chs_rt.h"

namespace chs {

CHS_CX constexpr int ilog2c(int x) { return x <= 1 ? 0 : 1 + ilog2c(x >> 1); }

// radix plan of the M-point complex FFT: an optional leading radix-2/4 stage, then radix-8
template <int M>
struct Rad {
    static constexpr int lg = ilog2c(M);
    static constexpr int rem = lg % 3;
    static constexpr int nst = lg / 3 + (rem ? 1 : 0);
    CHS_CX static constexpr int radix(int s) { return (rem != 0 && s == 0) ? (1 << rem) : 8; }
    CHS_CX static constexpr int blocklen(int s) {
        int Lb = M;
        for (int i = 0; i < s; ++i) Lb /= radix(i);
        return Lb;
    }
};

template <int N_>
struct Geo {
    static constexpr int N = N_;
    static constexpr int M = N / 2;
    static constexpr int LINES = 16;
    static constexpr int LP = LINES + 1;                       // line pitch (odd: conflict-free transposing I/O)
    static constexpr int TPL = (M / 8 < 32) ? (M / 8) : 32;    // threads per line
    static constexpr int NT = LINES * TPL;                     // threads per CTA
    static constexpr int NTILES = N / LINES;
    static constexpr int MINB = (N <= 512) ? 2 : 1;             // CTAs per SM the register budget is sized for
    static constexpr int TILE_DOUBLES = N * LP;
    static constexpr int SCRATCH_DOUBLES = 2 * TPL + (NT / 32 + 1) * 8 + 8 + 2 * 128;   // + fast_log table
    static constexpr int SMEM_BYTES = (TILE_DOUBLES + SCRATCH_DOUBLES) * 8;
    static_assert(N >= 32 && (N & (N - 1)) == 0, "FFT path needs a power of two >= 32");
};

// Makhoul reorder: physical index n -> position in v
template <int N>
CHS_DEV int mk_pos(int n) { return (n & 1) ? (N - 1 - (n >> 1)) : (n >> 1); }

// position of frequency k after the in-place DIF (mixed-radix digit reversal)
template <int M>
CHS_DEV int freq_pos(int k) {
    int pos = 0;
#pragma unroll
    for (int s = 0; s < Rad<M>::nst; ++s) {
        const int r = Rad<M>::radix(s), Lb = Rad<M>::blocklen(s);
        pos += (k % r) * (Lb / r);
        k /= r;
    }
    return pos;
}

// ------------------------------------------------------------------ small DFTs (natural order out)
template <bool INV>
CHS_DEV void dft4(double* r, double* i) {
    const double c0r = r[0] + r[2], c0i = i[0] + i[2];
    const double c1r = r[0] - r[2], c1i = i[0] - i[2];
    const double c2r = r[1] + r[3], c2i = i[1] + i[3];
    const double dr = r[1] - r[3], di = i[1] - i[3];
    const double c3r = INV ? -di : di, c3i = INV ? dr : -dr;   // d * (-/+ i)
    r[0] = c0r + c2r; i[0] = c0i + c2i;
    r[2] = c0r - c2r; i[2] = c0i - c2i;
    r[1] = c1r + c3r; i[1] = c1i + c3i;
    r[3] = c1r - c3r; i[3] = c1i - c3i;
}

template <int R, bool INV>
CHS_DEV void dft(double (&xr)[R], double (&xi)[R]) {
    if constexpr (R == 2) {
        const double tr = xr[0] - xr[1], ti = xi[0] - xi[1];
        xr[0] += xr[1]; xi[0] += xi[1];
        xr[1] = tr; xi[1] = ti;
    } else if constexpr (R == 4) {
        dft4<INV>(xr, xi);
    } else {
        static_assert(R == 8, "radix");
        constexpr double h = 0.70710678118654752440;
        double ar[4], ai[4], br[4], bi[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            ar[q] = xr[q] + xr[q + 4]; ai[q] = xi[q] + xi[q + 4];
            br[q] = xr[q] - xr[q + 4]; bi[q] = xi[q] - xi[q + 4];
        }
        // b_q *= W8^q,  W8 = exp(-/+ i pi/4)
        double x, y;
        x = br[1]; y = bi[1];
        if (INV) { br[1] = (x - y) * h; bi[1] = (x + y) * h; } else { br[1] = (x + y) * h; bi[1] = (y - x) * h; }
        x = br[2]; y = bi[2];
        if (INV) { br[2] = -y; bi[2] = x; } else { br[2] = y; bi[2] = -x; }
        x = br[3]; y = bi[3];
        if (INV) { br[3] = (-x - y) * h; bi[3] = (x - y) * h; } else { br[3] = (y - x) * h; bi[3] = (-x - y) * h; }
        dft4<INV>(ar, ai);
        dft4<INV>(br, bi);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            xr[2 * m] = ar[m]; xi[2 * m] = ai[m];
            xr[2 * m + 1] = br[m]; xi[2 * m + 1] = bi[m];
        }
    }
}

// ------------------------------------------------------------------ FFT stages on one line
// sl = smem + l (line base), t = thread index within the line, tw[m] = exp(-2 pi i m / M)
template <int N, int S>
CHS_DEV void fft_fwd_stage(double* sl, int t, const double2* __restrict__ tw) {
    using G = Geo<N>;
    constexpr int M = G::M, LP = G::LP, TPL = G::TPL;
    constexpr int r = Rad<M>::radix(S), Lb = Rad<M>::blocklen(S), st = Lb / r;
#pragma unroll
    for (int u0 = 0; u0 < M / r; u0 += TPL) {
        const int u = u0 + t;
        if (M / r < TPL && u >= M / r) break;
        const int j = u % st, base = (u / st) * Lb + j;
        double xr[r], xi[r];
#pragma unroll
        for (int q = 0; q < r; ++q) {
            const int c = base + q * st;
            xr[q] = sl[(2 * c) * LP];
            xi[q] = sl[(2 * c + 1) * LP];
        }
        dft<r, false>(xr, xi);
        if (st > 1) {
#pragma unroll
            for (int p = 1; p < r; ++p) {
                const double2 w = __ldg(tw + j * p * (M / Lb));
                const double a = xr[p], b = xi[p];
                xr[p] = a * w.x - b * w.y;
                xi[p] = a * w.y + b * w.x;
            }
        }
#pragma unroll
        for (int q = 0; q < r; ++q) {
            const int c = base + q * st;
            sl[(2 * c) * LP] = xr[q];
            sl[(2 * c + 1) * LP] = xi[q];
        }
    }
}

template <int N, int S>
CHS_DEV void fft_inv_stage(double* sl, int t, const double2* __restrict__ tw) {
    using G = Geo<N>;
    constexpr int M = G::M, LP = G::LP, TPL = G::TPL;
    constexpr int r = Rad<M>::radix(S), Lb = Rad<M>::blocklen(S), st = Lb / r;
#pragma unroll
    for (int u0 = 0; u0 < M / r; u0 += TPL) {
        const int u = u0 + t;
        if (M / r < TPL && u >= M / r) break;
        const int j = u % st, base = (u / st) * Lb + j;
        double xr[r], xi[r];
#pragma unroll
        for (int q = 0; q < r; ++q) {
            const int c = base + q * st;
            xr[q] = sl[(2 * c) * LP];
            xi[q] = sl[(2 * c + 1) * LP];
        }
        if (st > 1) {
#pragma unroll
            for (int p = 1; p < r; ++p) {
                const double2 w = __ldg(tw + j * p * (M / Lb));     // conj(w) applied
                const double a = xr[p], b = xi[p];
                xr[p] = a * w.x + b * w.y;
                xi[p] = b * w.x - a * w.y;
            }
        }
        dft<r, true>(xr, xi);
#pragma unroll
        for (int q = 0; q < r; ++q) {
            const int c = base + q * st;
            sl[(2 * c) * LP] = xr[q];
            sl[(2 * c + 1) * LP] = xi[q];
        }
    }
}

// all stages; every stage is followed by a block barrier
template <int N, int S = 0>
CHS_DEV void fft_fwd(double* sl, int t, const double2* __restrict__ tw) {
    if constexpr (S < Rad<N / 2>::nst) {
        fft_fwd_stage<N, S>(sl, t, tw);
        __syncthreads();
        fft_fwd<N, S + 1>(sl, t, tw);
    }
}
template <int N, int S = Rad<N / 2>::nst - 1>
CHS_DEV void fft_inv(double* sl, int t, const double2* __restrict__ tw) {
    if constexpr (S >= 0) {
        fft_inv_stage<N, S>(sl, t, tw);
        __syncthreads();
        fft_inv<N, S - 1>(sl, t, tw);
    }
}

// ------------------------------------------------------------------ post / pre passes
// Work item k in [0, M/2): k >= 1 couples Z[k], Z[M-k] <-> C[k], C[N-k], C[M-k], C[M+k];
// item 0 couples Z[0], Z[M/2] <-> C[0], C[M], C[M/2], C[3M/2].   om[m] = exp(-i pi m / (2N)).
template <int N>
CHS_DEV void item_index(int k, int (&idx)[4]) {
    constexpr int M = N / 2;
    if (k == 0) { idx[0] = 0; idx[1] = M; idx[2] = M / 2; idx[3] = M + M / 2; }
    else { idx[0] = k; idx[1] = N - k; idx[2] = M - k; idx[3] = M + k; }
}

template <int N>
CHS_DEV void post_item(const double* sl, int k, const double2* __restrict__ om, double (&c)[4]) {
    constexpr int M = N / 2, LP = Geo<N>::LP;
    if (k == 0) {
        const int p0 = freq_pos<M>(0), ph = freq_pos<M>(M / 2);
        const double ar = sl[(2 * p0) * LP], ai = sl[(2 * p0 + 1) * LP];
        const double hr = sl[(2 * ph) * LP], hi = sl[(2 * ph + 1) * LP];
        const double rn = sqrt(1.0 / N), s = sqrt(2.0 / N);
        const double2 w = __ldg(om + M / 2);
        c[0] = rn * (ar + ai);
        c[1] = rn * (ar - ai);
        c[2] = s * (w.x * hr + w.y * hi);          // Re(w * conj(Zh))
        c[3] = -s * (w.y * hr - w.x * hi);         // -Im(w * conj(Zh))
    } else {
        const int pa = freq_pos<M>(k), pb = freq_pos<M>(M - k);
        const double ar = sl[(2 * pa) * LP], ai = sl[(2 * pa + 1) * LP];
        const double br = sl[(2 * pb) * LP], bi = sl[(2 * pb + 1) * LP];
        const double2 t = __ldg(om + 4 * k), wk = __ldg(om + k), wm = __ldg(om + (M - k));
        const double sc = 0.5 * sqrt(2.0 / N);
        const double er = ar + br, ei = ai - bi;                  // 2E
        const double o_r = ai + bi, o_i = br - ar;                // 2O = -i (a - conj b)
        const double tr = t.x * o_r - t.y * o_i, ti = t.x * o_i + t.y * o_r;
        const double pr = (er + tr) * sc, pi = (ei + ti) * sc;
        const double qr = (er - tr) * sc, qi = (ei - ti) * sc;
        c[0] = wk.x * pr - wk.y * pi;
        c[1] = -(wk.x * pi + wk.y * pr);
        c[2] = wm.x * qr + wm.y * qi;                             // Re(wm * conj(Q))
        c[3] = wm.x * qi - wm.y * qr;                             // -Im(wm * conj(Q))
    }
}

template <int N>
CHS_DEV void pre_item(double* sl, int k, const double2* __restrict__ om, const double (&c)[4]) {
    constexpr int M = N / 2, LP = Geo<N>::LP;
    if (k == 0) {
        const int p0 = freq_pos<M>(0), ph = freq_pos<M>(M / 2);
        const double rn = sqrt(1.0 / N), is = 1.0 / (sqrt(2.0 / N) * M);
        const double2 w = __ldg(om + M / 2);
        sl[(2 * p0) * LP] = rn * (c[0] + c[1]);
        sl[(2 * p0 + 1) * LP] = rn * (c[0] - c[1]);
        const double u = c[2] * is, v = -c[3] * is;               // A/(sM) = u + i v
        const double vr = w.x * u + w.y * v, vi = w.x * v - w.y * u;   // conj(w)*A, w=(x,y): (x - i y)(u + i v)
        sl[(2 * ph) * LP] = vr;
        sl[(2 * ph + 1) * LP] = -vi;
    } else {
        const int pa = freq_pos<M>(k), pb = freq_pos<M>(M - k);
        const double2 t = __ldg(om + 4 * k), wk = __ldg(om + k), wm = __ldg(om + (M - k));
        const double sc = 1.0 / (sqrt(2.0 / N) * N);              // 1/(2 s M)
        // V = conj(wk) * (c0 - i c1),  V2 = conj(wm) * (c2 - i c3)
        const double vr = wk.x * c[0] - wk.y * c[1], vi = -wk.x * c[1] - wk.y * c[0];
        const double v2r = wm.x * c[2] - wm.y * c[3], v2i = -wm.x * c[3] - wm.y * c[2];
        const double er = (vr + v2r) * sc, ei = (vi - v2i) * sc;  // E
        const double dr = (vr - v2r) * sc, di = (vi + v2i) * sc;  // V - conj(V2)
        const double o_r = dr * t.x + di * t.y, o_i = di * t.x - dr * t.y;   // O = D * conj(t)
        sl[(2 * pa) * LP] = er - o_i;
        sl[(2 * pa + 1) * LP] = ei + o_r;
        sl[(2 * pb) * LP] = er + o_i;
        sl[(2 * pb + 1) * LP] = o_r - ei;
    }
}

}  // namespace chs
