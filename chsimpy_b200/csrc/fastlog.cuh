// Table-driven natural logarithm for the chemical potential / free energy (reference
// chsimpy/solver.py:173,220 call np.log three times per element; it is the largest single
// consumer of FP64 issue slots in the step).
//
//   x = 2^k * z,  z in [0.6875, 1.375);  i = top LOG_BITS = 8 mantissa bits of (bits(x) - bits(0.6875))
//   table[i] = { invc, logc }  with invc = double(1/c_i), c_i := 1/invc (c_i ~ centre of
//   sub-interval i), logc = log(c_i) rounded to double
//   r = fma(z, invc, -1)              exact to one rounding of a 2^-8-sized number
//   log x = k ln2 + logc + log1p(r),  log1p(r) = r - r^2/2 + r^3/3 - ... - r^8/8  (|r| < 2^-7)
// The summation keeps a hi/lo split (k*Ln2hi is exact: Ln2hi has 11 trailing zero bits), so
// the result is within 1.2 ulp for |log x| >= 2^-7 (0.6 ulp except where k ln2 and log c cancel,
// x just below 0.6875; measured against 120-bit mpmath in tests/test_emu_kernels.py) and within
// 1e-18 absolutely below that.
// This is the evaluation scheme of the ARM optimized-routines / glibc 2.28+ log() (Szabolcs
// Nagy, 2018), restated; the table is generated on the host at library load (chs_api.cu).
// Non-finite, zero, negative and subnormal arguments take the libm slow path.
#pragma once
#include "chs_rt.h"

namespace CHS_NS {

constexpr int LOG_TABLE_N = CHS_LOG_N;
constexpr int LOG_BITS = CHS_LOG_BITS;               // sub-interval index = top LOG_BITS mantissa bits
constexpr int LOG_SHIFT = 52 - LOG_BITS;             // ... of the 64-bit pattern; 20 - LOG_BITS of the high word
static_assert((1 << LOG_BITS) == LOG_TABLE_N, "table size");
constexpr unsigned long long LOG_OFF = 0x3fe6000000000000ULL;

#ifdef CHS_EMU
static inline int chs_hiword(double x) { long long v; std::memcpy(&v, &x, 8); return (int)(v >> 32); }
static inline double chs_sethiword(double x, int hi) {
    unsigned long long v; std::memcpy(&v, &x, 8);
    v = (v & 0xffffffffULL) | ((unsigned long long)(unsigned)hi << 32);
    double r; std::memcpy(&r, &v, 8); return r;
}
static inline long long chs_d2ll(double x) { long long v; std::memcpy(&v, &x, 8); return v; }
static inline double chs_ll2d(long long v) { double x; std::memcpy(&x, &v, 8); return x; }
static inline double chs_fma(double a, double b, double c) { return std::fma(a, b, c); }
#else
CHS_DEV int chs_hiword(double x) { return __double2hiint(x); }
CHS_DEV double chs_sethiword(double x, int hi) { return __hiloint2double(hi, __double2loint(x)); }
CHS_DEV long long chs_d2ll(double x) { return __double_as_longlong(x); }
CHS_DEV double chs_ll2d(long long v) { return __longlong_as_double(v); }
CHS_DEV double chs_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
#endif

// out-of-line slow path (keeps the 64 inlined copies of fast_log small: the instruction
// cache matters for the fully unrolled row kernel)
#ifdef CHS_EMU
static inline double slow_log(double x) { return std::log(x); }
static inline double div_ge1(double a, double b) { return a / b; }
#else
__device__ __noinline__ double slow_log(double x) { return log(x); }
// a / b for finite b >= 1: reciprocal seed + 2 Newton steps + one residual correction
// (no denormal/overflow slow path needed, unlike the generic IEEE division sequence)
CHS_DEV double div_ge1(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = __fma_rn(-b, r, 1.0);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-b, r, 1.0);
    r = __fma_rn(r, e, r);
    const double q = a * r;
    return __fma_rn(__fma_rn(-b, q, a), r, q);
}
#endif

// tab: LOG_TABLE_N entries {invc, logc} (shared or global memory)
// true for arguments the table scheme does not cover: <= 0, subnormal, inf, nan
CHS_DEV bool log_needs_slow_path(double x) { return (unsigned)(chs_hiword(x) - 0x00100000) >= 0x7fe00000u; }

// true for u in [2^-1022, 1): then both u and 1 - u (>= 2^-53) are positive normal numbers, i.e. ONE
// unsigned compare on the high word covers both logarithms of the chemical potential
CHS_DEV bool in_open_unit_interval(double u) { return (unsigned)(chs_hiword(u) - 0x00100000) < (0x3ff00000u - 0x00100000u); }

// log x with a small ABSOLUTE error (<= ~1.5 ulp of max(|log x|, 2^-8); no hi/lo compensation, degree-7
// log1p series): what the free energy and the chemical potential need -- log u and log(1-u) enter them
// additively next to terms of size O(1), so the relative accuracy of fast_log near x = 1 buys nothing.
// 8 FP64 instructions (+ the int -> double conversion of k) instead of 16.  tab entry i at tab[i * STRIDE].
template <int STRIDE = 1>
CHS_DEV double log_abs_unchecked(double x, const double2* __restrict__ tab) {
    const int hx = chs_hiword(x);
    const int tmp = hx - 0x3fe60000;                                  // LOG_OFF >> 32
    const int i = (tmp >> (20 - LOG_BITS)) & (LOG_TABLE_N - 1);
    const int k = tmp >> 20;                                          // arithmetic shift: floor
    const double z = chs_sethiword(x, hx - (int)((unsigned)tmp & 0xfff00000u));
    const double2 e = tab[i * STRIDE];
    const double r = chs_fma(z, e.x, -1.0);
    constexpr double Ln2 = 0x1.62e42fefa39efp-1;
    const double w = chs_fma((double)k, Ln2, e.y);
    const double r2 = r * r;
    // log1p(r) = r + r2*(-1/2 + r/3 - r^2/4 + r^3/5),  |r| < 2^-9 with the 256-entry table: the first dropped term
    // r^6/6 is < 2^-56.5 = 1e-17 (with 128 entries, |r| < 2^-8, two more terms were needed: 4 FP64 instructions per
    // value of the field).  Horner in r: every FMA but the first has ONE literal operand (an FP64 instruction takes one
    // constant / immediate; a second literal is a register pair the compiler re-materialises with two moves per use
    // -- an Estrin form has one such FMA per coefficient pair).  The two logarithms of a value and the values of a
    // butterfly give the scheduler its parallelism.
    double h = chs_fma(r, 1.0 / 5, -0.25);
    h = chs_fma(h, r, 1.0 / 3);
    h = chs_fma(h, r, -0.5);
    return w + chs_fma(r2, h, r);
}

// fast path only: the caller checks log_needs_slow_path() (garbage, but no trap, for such arguments)
template <int STRIDE = 1>
CHS_DEV double fast_log_unchecked(double x, const double2* __restrict__ tab) {
    // all bit manipulation on the high 32-bit word (sign, exponent, 20 mantissa bits)
    const int hx = chs_hiword(x);
    const int tmp = hx - 0x3fe60000;                                  // LOG_OFF >> 32
    const int i = (tmp >> (20 - LOG_BITS)) & (LOG_TABLE_N - 1);
    const int k = tmp >> 20;                                          // arithmetic shift: floor
    const double z = chs_sethiword(x, hx - (int)((unsigned)tmp & 0xfff00000u));
    const double2 e = tab[i * STRIDE];
    const double r = chs_fma(z, e.x, -1.0);
    const double kd = (double)k;
    constexpr double Ln2hi = 0x1.62e42fefa3800p-1, Ln2lo = 0x1.ef35793c76730p-45;
    const double w = chs_fma(kd, Ln2hi, e.y);
    const double hi = w + r;
    const double lo = chs_fma(kd, Ln2lo, (w - hi) + r);
    const double r2 = r * r;
    // log1p(r) - r = r2*(-1/2 + r/3) + r2*r2*(-1/4 + r/5 + r2*(-1/6 + r/7 - r2/8))
    const double p = chs_fma(r2, chs_fma(r2, -1.0 / 8, chs_fma(r, 1.0 / 7, -1.0 / 6)), chs_fma(r, 1.0 / 5, -1.0 / 4));
    const double q = chs_fma(r, 1.0 / 3, -0.5);
    const double y = chs_fma(r2, chs_fma(r2, p, q), lo);
    return y + hi;
}

CHS_DEV double fast_log(double x, const double2* __restrict__ tab) {
    if (log_needs_slow_path(x)) return slow_log(x);
    return fast_log_unchecked<1>(x, tab);
}
CHS_DEV double log_abs(double x, const double2* __restrict__ tab) {
    if (log_needs_slow_path(x)) return slow_log(x);
    return log_abs_unchecked<1>(x, tab);
}

}  // namespace CHS_NS
