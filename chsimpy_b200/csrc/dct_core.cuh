// Shared-memory radix-8/4/2 FFT building blocks for the orthonormal DCT-II / DCT-III
// (scipy.fftpack.dctn/idctn(norm='ortho'), called at reference chsimpy/solver.py:159,201,208).
//
// Algorithm (one "line" = one row or one column of N reals, M = N/2):
//   DCT-II :  Makhoul even/odd reorder v -> pack z[n] = v[2n] + i v[2n+1] -> M-point complex
//             FFT (in place, decimation in frequency, digit-reversed output) -> "post" step
//             (real-FFT untangle + quarter-wave twiddle) -> C[k].
//   DCT-III:  the exact transpose: "pre" step -> in-place decimation-in-time inverse FFT
//             (digit-reversed input, natural output) -> un-reorder.
// The forward FFT leaves its output digit-reversed and the inverse consumes digit-reversed
// input, so no reordering pass is ever executed.  The post/pre step couples frequency k with
// M-k; both live in the two 8-point blocks of the LAST radix-8 stage that one thread owns
// (fused_* in chs_kernels.cuh), so post/pre never touch shared memory on their own.
//
// Tile layout in shared memory: 16 lines are transformed together and the line index is
// the lane dimension: complex element c of line l is the double2 at sc[c*LPC + l].  A
// half-warp therefore reads 16 consecutive double2 (LDS.128, conflict free for any
// butterfly stride) and shares its twiddle factors.
// tools/dct_model.py is the numpy model of the transform algebra.
#pragma once
#include "chs_rt.h"

namespace CHS_NS {

CHS_CX constexpr int ilog2c(int x) { return x <= 1 ? 0 : 1 + ilog2c(x >> 1); }

// Twiddle tables tw (M entries) and om (N + N/4 entries) of the batched sizes N = 32 .. 1024 can live in
// CONSTANT memory (-DCHS_CONST_TABLES=1): one 56 KB image for all sizes (N at double2 offset
// 1.75 (N - 32): tw, then om), read with indexed LDC -- off the L1/LSU data pipe that the tile traffic
// saturates.  Accessors tab_tw<N>/tab_om<N> fall back to the global tables (line-major sizes, host build).
#ifndef CHS_CONST_TABLES
#define CHS_CONST_TABLES 0
#endif
#define CHS_CTAB_ENTRIES 3584
#if CHS_CONST_TABLES && !defined(CHS_EMU)
__constant__ double2 c_tab[CHS_CTAB_ENTRIES];
#endif
CHS_CX constexpr int ctab_off(int N) { return 7 * (N - 32) / 4; }
template <int N>
CHS_DEV double2 tab_tw(const double2* __restrict__ tw, int i) {
#if CHS_CONST_TABLES && !defined(CHS_EMU)
    if constexpr (N <= 1024) return c_tab[ctab_off(N) + i];
#endif
    return __ldg(tw + i);
}
template <int N>
CHS_DEV double2 tab_om(const double2* __restrict__ om, int i) {
#if CHS_CONST_TABLES && !defined(CHS_EMU)
    if constexpr (N <= 1024) return c_tab[ctab_off(N) + N / 2 + i];
#endif
    return __ldg(om + i);
}

#ifdef CHS_EMU
static inline double chs_fmad(double a, double b, double c) { return std::fma(a, b, c); }
#else
CHS_DEV double chs_fmad(double a, double b, double c) { return __fma_rn(a, b, c); }
#endif

// Radix plan of the M-point complex FFT.  Default: an optional leading radix-2/4 stage, then radix-8.
// For M = 4 * 8^j of the point-major (batched) geometry -- N = 64 and N = 512 -- the radix-4 stage comes
// LAST instead (8 .. 8 4): the fused post/pre passes then hold two 4-point blocks (32 registers of data)
// instead of two 8-point blocks per pairing unit, which is what lets the step kernels run at 80 registers.
// Host mirror: plan_radices() in chs_api.cu.
#ifndef CHS_LAST4
#define CHS_LAST4 1
#endif
template <int M>
struct Rad {
    static constexpr int lg = ilog2c(M);
    static constexpr int rem = lg % 3;
    static constexpr int nst = lg / 3 + (rem ? 1 : 0);
    static constexpr bool LAST4 = CHS_LAST4 && rem == 2 && M <= 512;
    CHS_CX static constexpr int radix(int s) {
        return LAST4 ? (s == nst - 1 ? 4 : 8) : ((rem != 0 && s == 0) ? (1 << rem) : 8);
    }
    static constexpr int RL = radix(nst - 1);               // radix of the last stage (4 or 8)
    CHS_CX static constexpr int blocklen(int s) {
        int Lb = M;
        for (int i = 0; i < s; ++i) Lb /= radix(i);
        return Lb;
    }
    // Per-stage twiddle tables of the line-major geometry: stage s, power p, butterfly j at
    // tws[tws_off(s) + (p-1)*st + j] = exp(-2 pi i j p / Lb) -- the lanes of a warp are consecutive j
    // there, so a warp reads contiguous entries (the natural table tw[j*p*M/Lb] is a strided gather)
    CHS_CX static constexpr int tws_off(int s) {
        int o = 0;
        for (int i = 0; i < s; ++i) o += (radix(i) - 1) * (blocklen(i) / radix(i));
        return o;
    }
    static constexpr int tws_len = tws_off(nst - 1);      // the last stage has no twiddles
};

#ifndef CHS_LINES
#define CHS_LINES 8
#endif
CHS_CX constexpr int geo_lines(int N) { return N <= 1024 ? CHS_LINES : 1; }   // lines (rows / columns) per tile
#define CHS_LOG_N 256       // entries of the fast_log table (fastlog.cuh): 2^CHS_LOG_BITS
#define CHS_LOG_BITS 8
#define CHS_SIMK 16         // doubles of the per-simulation constants image staged in shared memory

template <int N_>
struct Geo {
    static constexpr int N = N_;
    static constexpr int M = N / 2;
    // 8 lines per tile up to N = 1024 (the batched kernels); the long rows of the slab path
    // (N >= 2048, row kernels only) use ONE line per CTA of M/16 threads, so that several CTAs
    // are resident per SM and overlap each other's load / transform / store phases
    static constexpr int LINES = geo_lines(N);
    // Two tile layouts (complex point c of line l, in double2 units):
    //   point-major (N <= 1024): c*LINES + l, the lines of a point adjacent (they are the lanes of a
    //     warp: a quarter-warp's 16-byte accesses are one contiguous 128-byte wavefront for any
    //     butterfly stride); the tile is exactly 16*M*LINES bytes, which is also the layout of one
    //     column tile of T in HBM (see "pair-major T" in chs_kernels.cuh): the column kernel moves its
    //     tile with ONE bulk copy each way;
    //   line-major (N >= 2048): l*LOFF + c + pad(c), every warp works on ONE line so that its
    //     16-byte accesses are contiguous; pad() skews the 8-point blocks of the fused last stage
    //     (the blocks of 8 consecutive residues lie M/8, M/16, ... apart: the top digits of the
    //     position) over the banks.  pad is additive over the strides the stages use, see step().
    // (A line-major geometry for N = 256 .. 1024 with the threads of a line inside one warp was measured
    // in round 1: k_row +3 %, k_col -33 % because every table load stops being a warp broadcast.)
    static constexpr bool WARP_LINES = false;
    static constexpr bool LINE_MAJOR = (N >= 2048);
#ifndef CHS_STAGED_TABLES
#define CHS_STAGED_TABLES 0
#endif
    // twiddles read from per-stage tables (Rad<M>::tws_off) and the contiguous copy of om[4k]: the
    // entries adjacent lanes need are adjacent in memory.  Always on for the line-major geometry; for the
    // point-major one (4 distinct entries per warp load) measured 216.5 k vs 222 k sim-steps/s at N=512,
    // batch 1024 (-DCHS_STAGED_TABLES=1), so the natural tables stay the default there
    static constexpr bool STAGED_TABLES = LINE_MAJOR || CHS_STAGED_TABLES;
    static constexpr int LPC = LINE_MAJOR ? 1 : LINES;         // point pitch
    // first radix 8: the top 3 position bits; 4: top 2 bits + the low bit of the next digit (x4);
    // 2: the low 2 bits of the second digit + the top bit (x4)
    static constexpr int LG = ilog2c(M), REM = LG % 3;
    static constexpr int SH1 = (REM == 0) ? LG - 3 : ((REM == 1) ? LG - 4 : LG - 2);
    static constexpr int SH2 = (REM == 0) ? 30 : ((REM == 1) ? LG - 1 : LG - 5);
    static constexpr int W2 = (REM == 0) ? 0 : 4;
    CHS_CX static constexpr int pad(int c) { return LINE_MAJOR ? (c >> SH1) + W2 * (c >> SH2) : 0; }
    CHS_CX static constexpr int idx(int c) { return LINE_MAJOR ? c + pad(c) : c * LPC; }
    // idx(base + q*st) = idx(base) + q*step(st) for the points of one butterfly
    CHS_CX static constexpr int step(int st) { return LINE_MAJOR ? st + pad(st) : st * LPC; }
    static constexpr int LOFF_MIN = M + (M >> SH1) + W2 * (M >> SH2);
    static constexpr int LOFF = LINE_MAJOR ? LOFF_MIN : 1;    // line offset
    // every stage either strides by a multiple of a pad term's period or stays inside one period
    CHS_CX static constexpr bool pad_ok() {
        for (int s = 0; LINE_MAJOR && s < Rad<M>::nst; ++s) {
            const int Lb = Rad<M>::blocklen(s), st = Lb / Rad<M>::radix(s);
            if (!(st >= (1 << SH1) || Lb <= (1 << SH1))) return false;
            if (W2 && !(st >= (1 << SH2) || Lb <= (1 << SH2))) return false;
        }
        return true;
    }
    // complex points per thread and stage: 16 (throughput geometry), or CHS_PPT = 8 in the low-latency build
    // (chs_ll.cu: twice the threads per line, half the serial work per thread)
    static constexpr int PPT = LINE_MAJOR ? 16 : CHS_PPT;
    static constexpr int TPL = M / PPT;                        // threads per line
    CHS_CX static constexpr int line_of(int tid) { return LINE_MAJOR ? tid / TPL : tid % LINES; }
    CHS_CX static constexpr int t_of(int tid) { return LINE_MAJOR ? tid % TPL : tid / LINES; }
    static constexpr int NT = LINES * TPL;                     // threads per CTA
    static constexpr int NTILES = N / LINES;
#ifndef CHS_SEQ_STAGES
#define CHS_SEQ_STAGES 0
#endif
    // Register-light stages (-DCHS_SEQ_STAGES=1 -DCHS_REGS=80|96): the butterflies of a thread are processed
    // one after the other instead of all loads first, which fits 5-6 CTAs per SM.  Measured on B200
    // (profiles/r2_variants.md): 6 CTAs at 80 registers are SLOWER than 4 at 128 (the L1 data pipe, not
    // latency, is the limiter: more resident warps only raise mio_throttle / L1 misses), so the default keeps
    // the batched stages at 128 registers.
    static constexpr bool SEQ = CHS_SEQ_STAGES && !LINE_MAJOR;
#ifndef CHS_REGS
#define CHS_REGS 128
#endif
    static constexpr int REGS = LINE_MAJOR ? 128 : CHS_REGS;
    static constexpr int MINB = (65536 / REGS) / NT > 16 ? 16 : ((65536 / REGS) / NT > 0 ? (65536 / REGS) / NT : 1);   // CTAs per SM
#ifndef CHS_MINB_ROW
#define CHS_MINB_ROW MINB
#endif
#ifndef CHS_MINB_COL
#define CHS_MINB_COL MINB
#endif
    static constexpr int MINB_ROW = (NT == 128) ? CHS_MINB_ROW : MINB;   // N = 512: tuned on B200 (profiles/)
    static constexpr int MINB_COL = (NT == 128) ? CHS_MINB_COL : MINB;
    static constexpr int TILE_DOUBLES = LINE_MAJOR ? 2 * LINES * LOFF : 2 * M * LPC;
    static constexpr int LOG_STRIDE = 1;                       // fast_log table entry pitch (double2)
    // scratch after the tile (doubles): flag, mbarrier | x/y edge values | Ra | Sim image | fast_log table | reduction
    static constexpr int OFF_FLAG = TILE_DOUBLES;              // int flag; OFF_FLAG + 1: the bulk-copy mbarrier
    static constexpr int OFF_EDGE = OFF_FLAG + 2;              // [LINES][4]
    static constexpr int OFF_RA = OFF_EDGE + 4 * LINES;        // mean, spare, then TPL partials
    static constexpr int OFF_SIM = OFF_RA + 2 + TPL + (TPL & 1);         // staged per-simulation constants (CHS_SIMK doubles)
    static constexpr int OFF_LOGTAB = OFF_SIM + CHS_SIMK;      // CHS_LOG_N double2 (keeps double2 alignment)
    static constexpr int OFF_RED = OFF_LOGTAB + 2 * CHS_LOG_N; // (NT/32)*4 + 4 doubles on the GPU; NT*8 in the host emulation
    static constexpr int SMEM_BYTES = (OFF_RED + (NT / 32) * 4 + 4) * 8;
    CHS_CX static constexpr int log_off() { return OFF_LOGTAB; }          // doubles from the tile base
    static_assert(N >= 32 && (N & (N - 1)) == 0, "FFT path needs a power of two >= 32");
    static_assert((Rad<M>::RL == 8 || Rad<M>::RL == 4) && Rad<M>::nst >= 2, "plan must end with a radix-8 or radix-4 stage");
    static_assert((OFF_LOGTAB % 2) == 0 && (OFF_SIM % 2) == 0, "double2 alignment");
    static_assert(pad_ok(), "bank skew is not additive for this radix plan");
    static_assert(!LINE_MAJOR || (SH1 >= 3 && (W2 == 0 || SH2 >= 3)), "skew must be constant inside an 8-point block");
};

// Makhoul reorder: physical index n -> position in v
template <int N>
CHS_DEV int mk_pos(int n) { return (n & 1) ? (N - 1 - (n >> 1)) : (n >> 1); }

// offset (in doubles) of real element p (= v index: complex p>>1, part p&1) of line 0
template <int N>
CHS_DEV int real_off(int p) { return 2 * Geo<N>::idx(p >> 1) + (p & 1); }

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
        // even half: a_q = x_q + x_{q+4};  odd half: b_q = (x_q - x_{q+4}) * W8^q, W8 = exp(-/+ i pi/4).
        // The factor h = 1/sqrt2 of W8 and W8^3 is not applied to b_1, b_3 but folded into the last
        // layer of the odd 4-point DFT (8 FMAs instead of 4 multiplications + 8 additions).
        double ar[4], ai[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { ar[q] = xr[q] + xr[q + 4]; ai[q] = xi[q] + xi[q + 4]; }
        const double b0r = xr[0] - xr[4], b0i = xi[0] - xi[4];
        const double d1r = xr[1] - xr[5], d1i = xi[1] - xi[5];
        const double d2r = xr[2] - xr[6], d2i = xi[2] - xi[6];
        const double d3r = xr[3] - xr[7], d3i = xi[3] - xi[7];
        // u1 = b_1/h, u3 = b_3/h, b2 = b_2
        const double u1r = INV ? d1r - d1i : d1r + d1i, u1i = INV ? d1r + d1i : d1i - d1r;
        const double u3r = INV ? -(d3r + d3i) : d3i - d3r, u3i = INV ? d3r - d3i : -(d3r + d3i);
        const double b2r = INV ? -d2i : d2i, b2i = INV ? d2r : -d2r;
        const double c0r = b0r + b2r, c0i = b0i + b2i, c1r = b0r - b2r, c1i = b0i - b2i;
        const double sr = u1r + u3r, si = u1i + u3i;              // (b_1 + b_3)/h
        const double er = u1r - u3r, ei = u1i - u3i;              // (b_1 - b_3)/h
        dft4<INV>(ar, ai);
#pragma unroll
        for (int m = 0; m < 4; ++m) { xr[2 * m] = ar[m]; xi[2 * m] = ai[m]; }
        xr[1] = chs_fmad(h, sr, c0r);  xi[1] = chs_fmad(h, si, c0i);
        xr[5] = chs_fmad(-h, sr, c0r); xi[5] = chs_fmad(-h, si, c0i);
        if (INV) {                                                // c3 = i (b_1 - b_3)
            xr[3] = chs_fmad(-h, ei, c1r); xi[3] = chs_fmad(h, er, c1i);
            xr[7] = chs_fmad(h, ei, c1r);  xi[7] = chs_fmad(-h, er, c1i);
        } else {                                                  // c3 = -i (b_1 - b_3)
            xr[3] = chs_fmad(h, ei, c1r);  xi[3] = chs_fmad(-h, er, c1i);
            xr[7] = chs_fmad(-h, ei, c1r); xi[7] = chs_fmad(h, er, c1i);
        }
    }
}

// ------------------------------------------------------------------ regular FFT stages
// scl = (double2*)tile + l*LOFF (line base), t = thread index within the line,
// tw[m] = exp(-2 pi i m / M) -- or, in the line-major geometry, the per-stage tables Rad<M>::tws_off
// describes.  Every thread owns 16/r butterflies of radix r.
template <int N, int S, bool INV>
CHS_DEV void fft_stage(double2* scl, int t, const double2* __restrict__ tw) {
    using G = Geo<N>;
    constexpr int M = G::M, TPL = G::TPL;
    constexpr int r = Rad<M>::radix(S), Lb = Rad<M>::blocklen(S), st = Lb / r;
    constexpr int NB = G::PPT / r;
    static_assert(NB >= 1, "a thread owns at least one butterfly per stage");
    if constexpr (G::SEQ) {
        // register-light form: one butterfly at a time (r points live)
#pragma unroll 1
        for (int i = 0; i < NB; ++i) {
            const int u = t + i * TPL;
            const int j = u % st;
            double2* p = scl + G::idx((u / st) * Lb + j);
            double xr[r], xi[r];
#pragma unroll
            for (int q = 0; q < r; ++q) {
                const double2 v = p[q * G::step(st)];
                xr[q] = v.x; xi[q] = v.y;
            }
            if (!INV) dft<r, false>(xr, xi);
            if (st > 1) {
#pragma unroll
                for (int q = 1; q < r; ++q) {
                    const double2 w = G::STAGED_TABLES ? __ldg(tw + Rad<M>::tws_off(S) + (q - 1) * st + j) : tab_tw<N>(tw, j * q * (M / Lb));
                    const double a = xr[q], b = xi[q];
                    if (!INV) { xr[q] = a * w.x - b * w.y; xi[q] = a * w.y + b * w.x; }
                    else      { xr[q] = a * w.x + b * w.y; xi[q] = b * w.x - a * w.y; }     // conj(w)
                }
            }
            if (INV) dft<r, true>(xr, xi);
#pragma unroll
            for (int q = 0; q < r; ++q) p[q * G::step(st)] = make_double2(xr[q], xi[q]);
        }
    } else {
    // all 16 points of the thread are loaded before the first butterfly and stored after the
    // last one: the compiler cannot prove that the in-place stores of one butterfly do not
    // alias the loads of the next, so this is what exposes the memory-level parallelism
    double xr[NB][r], xi[NB][r];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int u = t + i * TPL;
        const int base = (u / st) * Lb + (u % st);
#pragma unroll
        for (int q = 0; q < r; ++q) {
            const double2 v = scl[G::idx(base) + q * G::step(st)];
            xr[i][q] = v.x; xi[i][q] = v.y;
        }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int j = (t + i * TPL) % st;
        if (!INV) dft<r, false>(xr[i], xi[i]);
        if (st > 1) {
#pragma unroll
            for (int p = 1; p < r; ++p) {
                const double2 w = G::STAGED_TABLES ? __ldg(tw + Rad<M>::tws_off(S) + (p - 1) * st + j) : tab_tw<N>(tw, j * p * (M / Lb));
                const double a = xr[i][p], b = xi[i][p];
                if (!INV) { xr[i][p] = a * w.x - b * w.y; xi[i][p] = a * w.y + b * w.x; }
                else      { xr[i][p] = a * w.x + b * w.y; xi[i][p] = b * w.x - a * w.y; }     // conj(w)
            }
        }
        if (INV) dft<r, true>(xr[i], xi[i]);
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int u = t + i * TPL;
        const int base = (u / st) * Lb + (u % st);
#pragma unroll
        for (int q = 0; q < r; ++q) scl[G::idx(base) + q * G::step(st)] = make_double2(xr[i][q], xi[i][q]);
    }
    }
}

// barrier between two passes over ONE line: the threads of a line share a warp in the WARP_LINES
// geometry (WARP = true; callers whose next access crosses lines use the block barrier)
template <int N, bool WARP>
CHS_DEV void line_barrier() {
    if constexpr (WARP && Geo<N>::WARP_LINES) __syncwarp();
    else __syncthreads();
}
// forward stages [S0, S1): each followed by a barrier (block, or line with WARP)
template <int N, int S0, int S1, bool WARP = false>
CHS_DEV void fft_fwd_range(double2* scl, int t, const double2* __restrict__ tw) {
    if constexpr (S0 < S1) {
        fft_stage<N, S0, false>(scl, t, tw);
        line_barrier<N, WARP>();
        fft_fwd_range<N, S0 + 1, S1, WARP>(scl, t, tw);
    }
}
// inverse stages S1-1 down to S0: each followed by a barrier
template <int N, int S0, int S1, bool WARP = false>
CHS_DEV void fft_inv_range(double2* scl, int t, const double2* __restrict__ tw) {
    if constexpr (S0 < S1) {
        fft_stage<N, S1 - 1, true>(scl, t, tw);
        line_barrier<N, WARP>();
        fft_inv_range<N, S0, S1 - 1, WARP>(scl, t, tw);
    }
}

// ------------------------------------------------------------------ post / pre algebra (registers)
// Tables (host: fill_om in chs_api.cu):  om[m] = sc * exp(-i pi m / (2N)), m < N, with the common
// scale sc = sqrt(2/N)/2 of the forward untangle and of the inverse (1/(2 s M) = sc) folded in;
// om[N + k] = exp(-2 pi i k / N) (= the unscaled om[4k]), k < N/4.  For 1 <= k < M/2:
//   post: Z[k]=(ar,ai), Z[M-k]=(br,bi)  ->  c = {C[k], C[N-k], C[M-k], C[M+k]}
//   pre : the inverse map, including the 1/M of the unnormalised inverse FFT.
// special: Z[0], Z[M/2] <-> {C[0], C[M], C[M/2], C[3M/2]}.
template <int N>
CHS_DEV void post_pair(int k, const double2* __restrict__ om, double ar, double ai, double br, double bi,
                       double (&c)[4]) {
    constexpr int M = N / 2;
    const double2 t = tab_om<N>(om, N + k), wk = tab_om<N>(om, k), wm = tab_om<N>(om, M - k);
    const double er = ar + br, ei = ai - bi;                  // 2E
    const double o_r = ai + bi, o_i = br - ar;                // 2O = -i (a - conj b)
    const double tr = t.x * o_r - t.y * o_i, ti = t.x * o_i + t.y * o_r;
    const double pr = er + tr, pi = ei + ti;
    const double qr = er - tr, qi = ei - ti;
    c[0] = wk.x * pr - wk.y * pi;
    c[1] = -(wk.x * pi + wk.y * pr);
    c[2] = wm.x * qr + wm.y * qi;                             // Re(wm * conj(Q))
    c[3] = wm.x * qi - wm.y * qr;                             // -Im(wm * conj(Q))
}

template <int N>
CHS_DEV void post_special(const double2* __restrict__ om, double ar, double ai, double hr, double hi,
                          double (&c)[4]) {
    constexpr int M = N / 2;
    const double rn = sqrt(1.0 / N);
    const double2 w = tab_om<N>(om, M / 2);       // sc * exp(-i pi/8), sc = s/2
    c[0] = rn * (ar + ai);
    c[1] = rn * (ar - ai);
    c[2] = 2.0 * (w.x * hr + w.y * hi);        // s Re(w * conj(Zh))
    c[3] = -2.0 * (w.y * hr - w.x * hi);       // -s Im(w * conj(Zh))
}

template <int N>
CHS_DEV void pre_pair(int k, const double2* __restrict__ om, const double (&c)[4], double& ar, double& ai,
                      double& br, double& bi) {
    constexpr int M = N / 2;
    const double2 t = tab_om<N>(om, N + k), wk = tab_om<N>(om, k), wm = tab_om<N>(om, M - k);
    const double vr = wk.x * c[0] - wk.y * c[1], vi = -wk.x * c[1] - wk.y * c[0];       // sc conj(wk)(c0 - i c1)
    const double v2r = wm.x * c[2] - wm.y * c[3], v2i = -wm.x * c[3] - wm.y * c[2];
    const double er = vr + v2r, ei = vi - v2i;                // E
    const double dr = vr - v2r, di = vi + v2i;                // V - conj(V2)
    const double o_r = dr * t.x + di * t.y, o_i = di * t.x - dr * t.y;   // O = D * conj(t)
    ar = er - o_i; ai = ei + o_r;
    br = er + o_i; bi = o_r - ei;
}

template <int N>
CHS_DEV void pre_special(const double2* __restrict__ om, const double (&c)[4], double& ar, double& ai,
                         double& hr, double& hi) {
    constexpr int M = N / 2;
    const double rn = sqrt(1.0 / N);
    const double2 w = tab_om<N>(om, M / 2);       // sc * exp(-i pi/8); 1/(s M) = 2 sc
    ar = rn * (c[0] + c[1]);
    ai = rn * (c[0] - c[1]);
    const double u = 2.0 * c[2], v = -2.0 * c[3];             // A/(sM) = (u + i v) * sc
    const double vr = w.x * u + w.y * v, vi = w.x * v - w.y * u;   // conj(w) * A
    hr = vr; hi = -vi;
}

// Pairing units of the fused last stage.  The last stage has radix RL (8, or 4 with Rad<M>::LAST4); its
// block of residue rho (frequencies rho + Q c, Q = M/RL, natural order after the RL-point DFT) starts at
// complex position freq_pos(rho).  The post/pre step couples frequency k with M - k, i.e. the blocks of
// residues rho and Q - rho: unit u (1 <= u < Q/2) = blocks {u, Q - u}, unit 0 = the two self-paired blocks
// {0, Q/2}.  A thread owns 16 points per stage = NU = 16/(2 RL) units: u = t + i*TPL.
template <int N>
struct Pairing {
    static constexpr int M = N / 2, RL = Rad<M>::RL, Q = M / RL, H = RL / 2;
    static constexpr int NU = Geo<N>::PPT / (2 * RL), TPL = Geo<N>::TPL;
    static_assert(NU >= 1 && NU * TPL == Q / 2, "units must cover all residues");
};

// Column slot s (PERM order used by T and hat_U along the x-spectral axis) holds frequency:
//   s = 2*pos(k) + 0 -> k ;  s = 2*pos(k) + 1 -> N-k  (k = 0: M).   Host table `kof`.

}  // namespace CHS_NS
