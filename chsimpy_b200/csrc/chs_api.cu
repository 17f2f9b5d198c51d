// C ABI of libchs_b200.so (see include/chs_b200.h).  Host-side orchestration only: every
// number is produced by the kernels in chs_kernels.cuh.  Compiled by nvcc for sm_100a, or
// by g++ with -DCHS_EMU as the host test harness (chs_rt.h).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "chs_kernels.cuh"
#include "chs_gemm.cuh"

using namespace chs;

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return -1; }

#define CHS_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_));             \
    } while (0)

struct chs_solver {
    int device, N, batch;
    double *U, *hatU, *T, *rows;
    long long rows_cap;
    cudaStream_t stream;
    // carved from the caller's workspace
    Sim* sims;
    double* part;
    double* colpart;
    double2* tw;
    double2* om;
    double* lam;
    double* gsin;
    double2* lamg;
    int* kof;
    double2* logtab;
    int* index;
    double* mean;
    // host mirrors
    std::vector<Sim> hsims;
    std::vector<int> hindex;
    std::vector<char> hstale;       // U of sim i is older than hat_U (steps ran without per-step U storage): chs_end materialises it
    int n_running;
    long long launches;
    int num_sms;
    bool gemm;           // DCT-as-GEMM path (chs_gemm.cuh): small / non-power-of-two N
    int N8, LD;
    double *Cm, *Ct;
    int cap_col[3], cap_row[4];     // resident CTAs per kernel mode (persistent grids)
    int cap_mix;
    unsigned long long* trace;      // -DCHS_TRACE=1 builds only (chs_debug_trace)
    int one_mode;                   // COL/ROW_STEP_LL instantiations: -1 = when a launch is at most one tile per SM, 0 / 1 = forced (CHS_ONE_PER_SM)
    int ll_max;                     // low-latency kernels when at most this many simulations run (0 = never)
    int sub_sims;                   // L2 blocking: simulations per sub-batch of chs_steps (0 = off)
    int mix_mode;                   // -1: mixed launches when enough simulations run (default), 0: never, 1: whenever possible
    // optional per-kernel timing (bench.py)
    bool timing;
    std::vector<cudaEvent_t> events;       // 4 per iteration: before col, after col, after row, after diag
    size_t ev_used;
    bool ev_diag;
    double t_ms[3];
    long long t_iters;
    // mixed launches (k_mix): events e0 L0 e1 L1 ... of every chs_steps call, and which launches were solo halves
    std::vector<cudaEvent_t> mix_events;
    size_t mix_used;
    std::vector<std::pair<size_t, size_t>> mix_runs;      // (first event, launches) per call
    double mix_ms[2];                                      // [0] mixed launches, [1] the solo half launches at both ends
    long long mix_n[2];
    long long mix_iters;
};

static cudaEvent_t next_event(chs_solver* s) {
    if (s->ev_used == s->events.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        s->events.push_back(e);
    }
    return s->events[s->ev_used++];
}

static cudaEvent_t next_mix_event(chs_solver* s) {
    if (s->mix_used == s->mix_events.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        s->mix_events.push_back(e);
    }
    return s->mix_events[s->mix_used++];
}

static int drain_events(chs_solver* s) {
    if (s->ev_used == 0 && s->mix_used == 0) return 0;
    CHS_CUDA(cudaStreamSynchronize(s->stream));
    for (const auto& run : s->mix_runs) {
        for (size_t j = 0; j < run.second; ++j) {
            float a = 0;
            cudaEventElapsedTime(&a, s->mix_events[run.first + j], s->mix_events[run.first + j + 1]);
            const int solo = (j == 0 || j + 1 == run.second) ? 1 : 0;
            s->mix_ms[solo] += a;
            s->mix_n[solo] += 1;
        }
    }
    s->mix_runs.clear();
    s->mix_used = 0;
    for (size_t i = 0; i + 4 <= s->ev_used; i += 4) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, s->events[i], s->events[i + 1]);
        cudaEventElapsedTime(&b, s->events[i + 1], s->events[i + 2]);
        cudaEventElapsedTime(&c, s->events[i + 2], s->events[i + 3]);
        s->t_ms[0] += a; s->t_ms[1] += b; s->t_ms[2] += c;
    }
    s->ev_used = 0;
    return 0;
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Layout {
    size_t sims, part, colpart, tw, om, lam, gsin, lamg, kof, logtab, index, mean, cm, ct, total;
};
static bool fft_supports(int N) { return N == 32 || N == 64 || N == 128 || N == 256 || N == 512 || N == 1024; }
static bool gemm_supports(int N) { return N >= GEMM_MIN_N && N <= GEMM_MAX_N; }
// Variant selection per N (and batch) by measurement on B200 (profiles/r1b_gemm_vs_fft.md):
// N = 32 up to two resident waves of simulations is faster as tensor-core GEMMs inside one CTA
// (15 vs 27 us per step for one simulation), from N = 64 on the FFT path wins everywhere.
static bool use_gemm_path(int N, int batch) {
    static const int force = [] { const char* e = getenv("CHS_FORCE_GEMM"); return e ? atoi(e) : -1; }();
    if (!gemm_supports(N)) return false;
    if (!fft_supports(N)) return true;
    if (force >= 0) return force == 1;
    return N == 32 && batch <= 296;
}
static int gemm_n8(int N) { return (N + 7) / 8 * 8; }
static int gemm_ld(int N) { return gemm_n8(N) + 4; }

static Layout layout(int N, int batch) {
    Layout L;
    size_t o = 0;
    const int ntiles = (N + CHS_LINES - 1) / CHS_LINES;
    L.sims = o; o = align_up(o + sizeof(Sim) * (size_t)batch);
    L.part = o; o = align_up(o + sizeof(double) * (size_t)batch * P_NSLOT * ntiles);
    L.colpart = o; o = align_up(o + sizeof(double) * (size_t)batch * ntiles * N);
    L.tw = o; o = align_up(o + sizeof(double2) * (size_t)(N / 2));
    L.om = o; o = align_up(o + sizeof(double2) * (size_t)(N + N / 4));      // om[N] + the contiguous copy of om[4k], k < N/4
    L.lam = o; o = align_up(o + sizeof(double) * (size_t)N);
    L.gsin = o; o = align_up(o + sizeof(double) * (size_t)N);
    L.lamg = o; o = align_up(o + sizeof(double2) * (size_t)(3 * (N / 4) + 3));
    L.kof = o; o = align_up(o + sizeof(int) * (size_t)N);
    L.logtab = o; o = align_up(o + sizeof(double2) * (size_t)LOG_TABLE_N);
    L.index = o; o = align_up(o + sizeof(int) * (size_t)batch);
    L.mean = o; o = align_up(o + sizeof(double) * (size_t)batch);
    const size_t n8 = (size_t)gemm_n8(N);
    L.cm = o; o = align_up(o + sizeof(double) * n8 * n8);
    L.ct = o; o = align_up(o + sizeof(double) * n8 * n8);
    L.total = o;
    return L;
}

extern "C" int32_t chs_abi_version(void) { return CHS_ABI_VERSION; }
extern "C" const char* chs_last_error(void) { return g_err.c_str(); }

extern "C" int32_t chs_supports_n(int32_t N) { return (fft_supports(N) || gemm_supports(N)) ? 1 : 0; }
extern "C" int32_t chs_uses_gemm(int32_t N, int32_t batch) { return use_gemm_path(N, batch) ? 1 : 0; }

extern "C" int64_t chs_workspace_bytes(int32_t N, int32_t batch) {
    if (!chs_supports_n(N) || batch < 1) return -1;
    return (int64_t)layout(N, batch).total;
}

// ---- dispatch on N ---------------------------------------------------------------------
#define CHS_FOR_N(N_, CALL)                  \
    switch (N_) {                            \
        case 32: { CALL(32); } break;        \
        case 64: { CALL(64); } break;        \
        case 128: { CALL(128); } break;      \
        case 256: { CALL(256); } break;      \
        case 512: { CALL(512); } break;      \
        case 1024: { CALL(1024); } break;    \
        default: return fail("unsupported N"); \
    }


// radix plan of the M-point FFT: the host mirror of Rad<M> (dct_core.cuh)
static std::vector<int> plan_radices(int M) {
    std::vector<int> rad;
    int lg = 0;
    while ((1 << lg) < M) ++lg;
    const int rem = lg % 3, nst = lg / 3 + (rem ? 1 : 0);
    const bool last4 = CHS_LAST4 && rem == 2 && M <= 512;
    for (int s = 0; s < nst; ++s) rad.push_back(last4 ? (s == nst - 1 ? 4 : 8) : ((rem != 0 && s == 0) ? (1 << rem) : 8));
    return rad;
}

// post/pre tables (dct_core.cuh): om[m] = sc exp(-i pi m/(2N)), m < N, sc = sqrt(2/N)/2; om[N + k] = exp(-2 pi i k/N), k < N/4
static void fill_om(int N, std::vector<double2>& om) {
    const long double pi = 3.14159265358979323846264338327950288L;
    const long double sc = sqrtl(2.0L / N) * 0.5L;
    om.assign((size_t)N + N / 4, make_double2(0.0, 0.0));
    for (int m = 0; m < N; ++m) {
        const long double a = -pi * m / (2.0L * N);
        om[m] = make_double2((double)(sc * cosl(a)), (double)(sc * sinl(a)));
    }
    for (int k = 0; k < N / 4; ++k) {
        const long double a = -2.0L * pi * k / N;
        om[N + k] = make_double2((double)cosl(a), (double)sinl(a));
    }
}

// packed per-item table of the spectral update (ColMid): for k < M/2 three double2
//   {lam[k], lam[N-k]}, {lam[M-k], lam[M+k]}, {g[k], g[M-k]}       (k = 0: rows 0, M, M/2, 3M/2; g = {0, 1/2})
static void fill_lamg(int N, const double* lam, const std::vector<double>& gs, std::vector<double2>& t) {
    const int M = N / 2;
    t.assign((size_t)3 * (M / 2), make_double2(0.0, 0.0));
    for (int k = 0; k < M / 2; ++k) {
        const int r0 = k, r1 = (k == 0) ? M : N - k, r2 = (k == 0) ? M / 2 : M - k, r3 = (k == 0) ? M + M / 2 : M + k;
        t[3 * k] = make_double2(lam[r0], lam[r1]);
        t[3 * k + 1] = make_double2(lam[r2], lam[r3]);
        t[3 * k + 2] = make_double2(gs[r0], gs[r2]);
    }
}

// slot -> frequency map of the plan: slot 2 pos(k) holds k, slot 2 pos(k) + 1 holds N - k (k = 0: M)
static void fill_kof(int N, std::vector<int>& kof) {
    const int M = N / 2;
    const std::vector<int> rad = plan_radices(M);
    kof.assign(N, 0);
    for (int k = 0; k < M; ++k) {
        int pos = 0, Lb = M, kk = k;
        for (int r : rad) { pos += (kk % r) * (Lb / r); kk /= r; Lb /= r; }
        kof[2 * pos] = k;
        kof[2 * pos + 1] = (k == 0) ? M : N - k;
    }
}

// FFT twiddles for the kernels of size N: the natural table tw[m] = exp(-2 pi i m / M) for the point-major
// tile geometry, the per-stage tables of Rad<M>::tws_off (same values, warp-contiguous order) for the
// line-major one.  Always fewer than M entries.
static void fill_twiddles(int N, std::vector<double2>& tw) {
    const int M = N / 2;
    const long double pi = 3.14159265358979323846264338327950288L;
    std::vector<double2> nat(M);
    for (int m = 0; m < M; ++m) { const long double a = -2.0L * pi * m / M; nat[m] = make_double2((double)cosl(a), (double)sinl(a)); }
    const bool line_major = (N >= 2048) || CHS_STAGED_TABLES;   // = Geo<N>::STAGED_TABLES
    tw.assign(M, make_double2(0.0, 0.0));
    if (!line_major) { tw = nat; return; }
    const std::vector<int> rad = plan_radices(M);
    size_t o = 0; int Lb = M;
    for (size_t s = 0; s + 1 < rad.size(); ++s) {
        const int r = rad[s], st = Lb / r;
        for (int p = 1; p < r; ++p)
            for (int j = 0; j < st; ++j) tw[o + (size_t)(p - 1) * st + j] = nat[(size_t)j * p * (M / Lb)];
        o += (size_t)(r - 1) * st;
        Lb /= r;
    }
}

// constant-memory image of the twiddle tables (dct_core.cuh, CHS_CONST_TABLES): region of size N
static int upload_const_tables(int N, const std::vector<double2>& tw, const std::vector<double2>& om, cudaStream_t stream) {
#if CHS_CONST_TABLES && !defined(CHS_EMU)
    if (N > 1024 || N < 32) return 0;
    const size_t off = sizeof(double2) * (size_t)ctab_off(N);
    CHS_CUDA(cudaMemcpyToSymbolAsync(c_tab, tw.data(), sizeof(double2) * (N / 2), off, cudaMemcpyHostToDevice, stream));
    CHS_CUDA(cudaMemcpyToSymbolAsync(c_tab, om.data(), sizeof(double2) * (N + N / 4), off + sizeof(double2) * (N / 2), cudaMemcpyHostToDevice, stream));
#else
    (void)N; (void)tw; (void)om; (void)stream;
#endif
    return 0;
}

template <class K>
static int resident_ctas(K kern, int threads, int smem, int num_sms) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, (size_t)smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * num_sms;
}

// Dynamic shared memory of the *_LL launches: more than half of an SM's, so that no two of these CTAs -- of one
// kernel, or a kernel and its programmatically launched dependents -- ever share an SM (both kernels ask for the same
// amount: no carve-out change at the kernel boundaries).  With the tile's own 36 KB two column CTAs fitted one SM while
// the dependents were resident, and a single simulation's step waited 4 us for the two tiles that shared an SM.
static int ll_smem_bytes(int tile_bytes) { return tile_bytes > 116 * 1024 ? tile_bytes : 116 * 1024; }

template <int N>
static int set_attrs(chs_solver* s) {
    const int b = Geo<N>::SMEM_BYTES;
    CHS_CUDA(cudaFuncSetAttribute(k_col<N, COL_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_col<N, COL_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_col<N, COL_INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_row<N, ROW_FWD_U>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_row<N, ROW_FWD_MU>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_row<N, ROW_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_row<N, ROW_INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_mix<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_col<N, COL_STEP_LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, ll_smem_bytes(b)));
    CHS_CUDA(cudaFuncSetAttribute(k_row<N, ROW_STEP_LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, ll_smem_bytes(b)));
    CHS_CUDA(cudaFuncSetAttribute(k_diag<N, DIAG_PREPARE>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_diag<N, DIAG_JITTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
#ifndef CHS_EMU
    // tuning experiment: shared-memory carve-out of the two step kernels in percent (default: the driver's choice)
    if (const char* e = getenv("CHS_CARVEOUT")) {
        CHS_CUDA(cudaFuncSetAttribute(k_col<N, COL_STEP>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
        CHS_CUDA(cudaFuncSetAttribute(k_row<N, ROW_STEP>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
    }
#endif
    const int nt = Geo<N>::NT, ns = s->num_sms;
    s->cap_col[COL_FWD] = resident_ctas(k_col<N, COL_FWD>, nt, b, ns);
    s->cap_col[COL_STEP] = resident_ctas(k_col<N, COL_STEP>, nt, b, ns);
    s->cap_col[COL_INV] = resident_ctas(k_col<N, COL_INV>, nt, b, ns);
    s->cap_row[ROW_FWD_U] = resident_ctas(k_row<N, ROW_FWD_U>, nt, b, ns);
    s->cap_row[ROW_FWD_MU] = resident_ctas(k_row<N, ROW_FWD_MU>, nt, b, ns);
    s->cap_row[ROW_STEP] = resident_ctas(k_row<N, ROW_STEP>, nt, b, ns);
    s->cap_row[ROW_INV] = resident_ctas(k_row<N, ROW_INV>, nt, b, ns);
    s->cap_mix = resident_ctas(k_mix<N>, nt, b, ns);
    return 0;
}

// Grid of a tile kernel: one CTA per tile.  (-DCHS_PERSISTENT builds resident CTAs that loop
// over tiles instead -- measured slower on B200, kept as a compile-time experiment; the host
// emulation always loops so that a few OS-thread blocks cover all tiles.)
static dim3 pgrid(int cap, int num_sms, int ntiles, int nsims) {
    const long long total = (long long)ntiles * nsims;
    (void)num_sms;
#if defined(CHS_EMU) || defined(CHS_PERSISTENT)
    return dim3((unsigned)(total < cap ? total : cap));
#else
    (void)cap;
    return dim3((unsigned)total);
#endif
}

// fewest running simulations for which a step runs as mixed launches (k_mix): below it a half batch no longer
// fills the GPU for several waves and the two half-size launches at the ends of a call cost more than mixing gains
#ifndef CHS_MIX_MIN_SIMS
#define CHS_MIX_MIN_SIMS 32
#endif

// L2 blocking of chs_steps: simulations per sub-batch (0 = off), see do_steps
static int default_sub_sims(int N) {
    (void)N;
    return 0;
}

// low-latency build of the step kernels (chs_ll.cu: 8 points per thread, 256 threads per tile)
extern "C" int chs_ll_init(int N);
extern "C" int chs_ll_launch(int N, int which, const void* kargs, int nsims, int pdl, void* stream);
extern "C" int chs_ll_kargs_size(void);
// most running simulations for which chs_steps uses it (N = 512): measured on B200 (profiles/r2c_lowlat.md)
#ifndef CHS_LL_MAX_SIMS
#define CHS_LL_MAX_SIMS 0
#endif

static KArgs base_args(chs_solver* s) {
    KArgs a;
    std::memset(&a, 0, sizeof(a));
    a.sims = s->sims; a.sim_index = nullptr;
    a.U = s->U; a.hatU = s->hatU; a.T = s->T;
    a.rows = s->rows; a.rows_cap = s->rows_cap;
    a.part = s->part; a.colpart = s->colpart;
    a.tw = s->tw; a.om = s->om; a.lam = s->lam; a.logtab = s->logtab;
    a.gsin = s->gsin; a.kof = s->kof; a.lamg = s->lamg;
    a.mean_host = s->mean;
    a.trace = s->trace;
    return a;
}

extern "C" chs_solver* chs_create(int32_t device, int32_t N, int32_t batch, double* U, double* hat_U, double* T,
                                  double* rows, int64_t rows_cap, void* workspace, int64_t workspace_bytes,
                                  const double* lambda_host, void* stream) {
    if (!chs_supports_n(N)) { fail("chs_create: unsupported N (FFT path: powers of two 32..1024; GEMM path: 8..104)"); return nullptr; }
    if (batch < 1 || rows_cap < 1) { fail("chs_create: bad batch/rows_cap"); return nullptr; }
    const Layout L = layout(N, batch);
    if (workspace_bytes < (int64_t)L.total) { fail("chs_create: workspace too small"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail("chs_create: cudaSetDevice failed"); return nullptr; }
    chs_solver* s = new chs_solver();
    s->device = device; s->N = N; s->batch = batch;
    s->U = U; s->hatU = hat_U; s->T = T; s->rows = rows; s->rows_cap = rows_cap;
    s->stream = (cudaStream_t)stream;
    unsigned char* w = (unsigned char*)workspace;
    s->sims = (Sim*)(w + L.sims);
    s->part = (double*)(w + L.part);
    s->colpart = (double*)(w + L.colpart);
    s->tw = (double2*)(w + L.tw);
    s->om = (double2*)(w + L.om);
    s->lam = (double*)(w + L.lam);
    s->logtab = (double2*)(w + L.logtab);
    s->gsin = (double*)(w + L.gsin);
    s->lamg = (double2*)(w + L.lamg);
    s->kof = (int*)(w + L.kof);
    s->index = (int*)(w + L.index);
    s->mean = (double*)(w + L.mean);
    s->Cm = (double*)(w + L.cm); s->Ct = (double*)(w + L.ct);
    s->gemm = use_gemm_path(N, batch); s->N8 = gemm_n8(N); s->LD = gemm_ld(N);
    s->hsims.assign(batch, Sim());
    std::memset(s->hsims.data(), 0, sizeof(Sim) * batch);
    s->hindex.resize(batch);
    for (int i = 0; i < batch; ++i) s->hindex[i] = i;
    s->hstale.assign(batch, 0);
    s->n_running = batch;
    s->launches = 0;
    s->timing = false; s->ev_used = 0; s->ev_diag = false; s->t_iters = 0;
    s->t_ms[0] = s->t_ms[1] = s->t_ms[2] = 0;
    s->trace = nullptr;
    s->one_mode = [] { const char* e = getenv("CHS_ONE_PER_SM"); return e ? atoi(e) : -1; }();
    s->ll_max = [] { const char* e = getenv("CHS_LL_MAX"); return e ? atoi(e) : CHS_LL_MAX_SIMS; }();
    s->sub_sims = [N] { const char* e = getenv("CHS_SUB"); return e ? atoi(e) : default_sub_sims(N); }();
    s->mix_mode = [] { const char* e = getenv("CHS_MIX"); return e ? atoi(e) : 0; }();
    s->mix_used = 0; s->mix_ms[0] = s->mix_ms[1] = 0; s->mix_n[0] = s->mix_n[1] = 0; s->mix_iters = 0;
    // twiddle tables in extended precision, rounded once
    const int M = N / 2;
    std::vector<double2> tw(M), om(N + N / 4);
    std::vector<double> gs(N);
    std::vector<int> kof(N);
    std::vector<double2> lt(LOG_TABLE_N);
    std::vector<double> cm, ct;
    const long double pi = 3.14159265358979323846264338327950288L;
    if (!s->gemm) {
        fill_twiddles(N, tw);
        fill_om(N, om);
        // gradient-energy weights sin^2(pi k/N) and the slot -> frequency map of the FFT plan
        for (int k = 0; k < N; ++k) {
            const long double sn = sinl(pi * k / N);
            gs[k] = (double)(sn * sn);
        }
        fill_kof(N, kof);
    } else {
        // orthonormal DCT-II matrix C[k][n] = f_k cos(pi k (2n+1) / (2N)), zero padded to N8 x N8
        const int n8 = s->N8;
        cm.assign((size_t)n8 * n8, 0.0);
        ct.assign((size_t)n8 * n8, 0.0);
        for (int k = 0; k < N; ++k)
            for (int n = 0; n < N; ++n) {
                const long double f = (k == 0) ? sqrtl(1.0L / N) : sqrtl(2.0L / N);
                const double v = (double)(f * cosl(pi * k * (2 * n + 1) / (2.0L * N)));
                cm[(size_t)k * n8 + n] = v;
                ct[(size_t)n * n8 + k] = v;
            }
    }
    // fast_log table (fastlog.cuh): sub-interval centres of [0.6875, 1.375) in bit-pattern space
    for (int i = 0; i < LOG_TABLE_N; ++i) {
        const unsigned long long b0 = LOG_OFF + ((unsigned long long)i << LOG_SHIFT);
        const unsigned long long b1 = LOG_OFF + ((unsigned long long)(i + 1) << LOG_SHIFT);
        double z0, z1;
        std::memcpy(&z0, &b0, 8);
        std::memcpy(&z1, &b1, 8);
        const double invc = (double)(1.0L / ((long double)z0 * 0.5L + (long double)z1 * 0.5L));
        lt[i] = make_double2(invc, (double)(-logl((long double)invc)));
    }
    bool ok = true;
    ok &= cudaMemsetAsync(workspace, 0, L.total, s->stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(s->tw, tw.data(), sizeof(double2) * M, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(s->om, om.data(), sizeof(double2) * om.size(), cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(s->gsin, gs.data(), sizeof(double) * N, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(s->kof, kof.data(), sizeof(int) * N, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(s->logtab, lt.data(), sizeof(double2) * LOG_TABLE_N, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(s->lam, lambda_host, sizeof(double) * N, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    if (!s->gemm) ok &= upload_const_tables(N, tw, om, s->stream) == 0;
    std::vector<double2> lamg;
    if (!s->gemm) {
        fill_lamg(N, lambda_host, gs, lamg);
        ok &= cudaMemcpyAsync(s->lamg, lamg.data(), sizeof(double2) * lamg.size(), cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    }
    if (s->gemm) {
        ok &= cudaMemcpyAsync(s->Cm, cm.data(), sizeof(double) * cm.size(), cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
        ok &= cudaMemcpyAsync(s->Ct, ct.data(), sizeof(double) * ct.size(), cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    }
    ok &= cudaMemcpyAsync(s->index, s->hindex.data(), sizeof(int) * batch, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok &= cudaStreamSynchronize(s->stream) == cudaSuccess;
    int rc = 0;
    s->num_sms = 1;
    cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (s->gemm) {
        const int smem = (2 * s->N8 * s->LD + 8 * GEMM_NT + 2 * LOG_TABLE_N) * (int)sizeof(double);
        if (cudaFuncSetAttribute(k_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
            fail("chs_create: k_gemm shared memory attribute");
            rc = -1;
        }
    }
#define CALL(NN) rc = set_attrs<NN>(s);
    if (!s->gemm) switch (N) {
        case 32: CALL(32) break; case 64: CALL(64) break; case 128: CALL(128) break;
        case 256: CALL(256) break; case 512: CALL(512) break; case 1024: CALL(1024) break;
    }
#undef CALL
    if (!ok || rc != 0) {
        if (ok) { /* g_err set by set_attrs */ } else fail("chs_create: table upload failed");
        delete s;
        return nullptr;
    }
    return s;
}

extern "C" void chs_destroy(chs_solver* s) {
    if (!s) return;
    for (cudaEvent_t e : s->events) cudaEventDestroy(e);
    for (cudaEvent_t e : s->mix_events) cudaEventDestroy(e);
    delete s;
}

extern "C" int chs_set_timing(chs_solver* s, int32_t enable) {
    if (!s) return fail("chs_set_timing: null handle");
    if (drain_events(s)) return -1;
    s->timing = enable != 0;
    return 0;
}

extern "C" int chs_get_timing(chs_solver* s, double* ms3, int64_t* n_iters) {
    if (!s) return fail("chs_get_timing: null handle");
    if (drain_events(s)) return -1;
    if (ms3) { ms3[0] = s->t_ms[0]; ms3[1] = s->t_ms[1]; ms3[2] = s->t_ms[2]; }
    if (n_iters) *n_iters = s->t_iters;
    s->t_ms[0] = s->t_ms[1] = s->t_ms[2] = 0;
    s->t_iters = 0;
    return 0;
}

// latency analysis (tools/trace_single.py, library built with -DCHS_TRACE=1): device buffer of >= 32 uint64 that
// tile 0 of every step launch fills with %globaltimer phase stamps; not part of the public header
extern "C" int chs_debug_trace(chs_solver* s, unsigned long long* dev_buf) {
    if (!s) return fail("chs_debug_trace: null handle");
    s->trace = dev_buf;
    return 0;
}

extern "C" int chs_set_mix(chs_solver* s, int32_t mode) {
    if (!s || mode < -1 || mode > 1) return fail("chs_set_mix: bad argument");
    s->mix_mode = mode;
    return 0;
}

extern "C" int chs_get_timing_mix(chs_solver* s, double* ms2, int64_t* n2, int64_t* n_iters) {
    if (!s) return fail("chs_get_timing_mix: null handle");
    if (drain_events(s)) return -1;
    if (ms2) { ms2[0] = s->mix_ms[0]; ms2[1] = s->mix_ms[1]; }
    if (n2) { n2[0] = s->mix_n[0]; n2[1] = s->mix_n[1]; }
    if (n_iters) *n_iters = s->mix_iters;
    s->mix_ms[0] = s->mix_ms[1] = 0;
    s->mix_n[0] = s->mix_n[1] = 0;
    s->mix_iters = 0;
    return 0;
}

extern "C" int chs_set_params(chs_solver* s, int32_t sim, const chs_params* p) {
    if (!s || sim < 0 || sim >= s->batch || !p) return fail("chs_set_params: bad argument");
    Sim& h = s->hsims[sim];
    h.p = *p;
    h.delt = p->delt;
    h.delt_coef = p->delt;
    sim_derive(h);
    CHS_CUDA(cudaMemcpyAsync(s->sims + sim, &h, sizeof(Sim), cudaMemcpyHostToDevice, s->stream));
    CHS_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

static int pull(chs_solver* s) {
    CHS_CUDA(cudaMemcpyAsync(s->hsims.data(), s->sims, sizeof(Sim) * s->batch, cudaMemcpyDeviceToHost, s->stream));
    CHS_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int chs_set_state(chs_solver* s, int32_t sim, const chs_state* st) {
    if (!s || sim < 0 || sim >= s->batch || !st) return fail("chs_set_state: bad argument");
    if (pull(s)) return -1;
    Sim& h = s->hsims[sim];
    h.delt = st->delt; h.time_delta_sum = st->time_delta_sum; h.time_passed = st->time_passed;
    h.tau0 = st->tau0; h.t0 = st->t0; h.computed_steps = st->computed_steps;
    h.skip_check = st->skip_check; h.stop_reason = st->stop_reason;
    CHS_CUDA(cudaMemcpyAsync(s->sims + sim, &h, sizeof(Sim), cudaMemcpyHostToDevice, s->stream));
    CHS_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int chs_get_state(chs_solver* s, int32_t sim, chs_state* st) {
    if (!s || sim < 0 || sim >= s->batch || !st) return fail("chs_get_state: bad argument");
    if (pull(s)) return -1;
    const Sim& h = s->hsims[sim];
    st->delt = h.delt; st->time_delta_sum = h.time_delta_sum; st->time_passed = h.time_passed;
    st->tau0 = h.tau0; st->t0 = h.t0; st->computed_steps = h.computed_steps;
    st->skip_check = h.skip_check; st->stop_reason = h.stop_reason;
    return 0;
}

static int refresh_index(chs_solver* s) {
    s->hindex.clear();
    for (int i = 0; i < s->batch; ++i)
        if (!s->hsims[i].halted) s->hindex.push_back(i);
    s->n_running = (int)s->hindex.size();
    if (s->n_running > 0)
        CHS_CUDA(cudaMemcpyAsync(s->index, s->hindex.data(), sizeof(int) * s->n_running, cudaMemcpyHostToDevice, s->stream));
    return 0;
}

// ---- DCT-as-GEMM path (chs_gemm.cuh): one CTA per simulation --------------------------------
static int gemm_launch(chs_solver* s, int mode, long long n_iters, const double* noise, const double* src, double* dst,
                       bool running_only) {
    GemmArgs g;
    std::memset(&g, 0, sizeof(g));
    g.sims = s->sims; g.U = s->U; g.hatU = s->hatU; g.rows = s->rows; g.rows_cap = s->rows_cap;
    g.Cm = s->Cm; g.Ct = s->Ct; g.lam = s->lam; g.logtab = s->logtab; g.noise = noise; g.mean_host = s->mean;
    g.N = s->N; g.N8 = s->N8; g.LD = s->LD; g.mode = mode; g.n_iters = n_iters; g.src = src; g.dst = dst;
    int nsims = s->batch;
    if (running_only) {
        if (s->n_running <= 0) return 0;
        nsims = s->n_running;
        g.sim_index = (s->n_running == s->batch) ? nullptr : s->index;
    }
    const int smem = (2 * s->N8 * s->LD + 8 * GEMM_NT + 2 * LOG_TABLE_N) * (int)sizeof(double);
    CHS_LAUNCH(k_gemm, dim3(nsims), dim3(GEMM_NT), smem, s->stream, g);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

template <int N>
static int do_prepare(chs_solver* s) {
    using G = Geo<N>;
    KArgs a = base_args(s);
    CHS_LAUNCH((k_diag<N, DIAG_PREPARE>), dim3(G::NTILES, s->batch), dim3(G::NT), G::SMEM_BYTES, s->stream, a);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int chs_prepare(chs_solver* s, const double* mean_U_host) {
    if (!s || !mean_U_host) return fail("chs_prepare: bad argument");
    CHS_CUDA(cudaMemcpyAsync(s->mean, mean_U_host, sizeof(double) * s->batch, cudaMemcpyHostToDevice, s->stream));
    CHS_CUDA(cudaStreamSynchronize(s->stream));        // mean_U_host may be pageable and short-lived
    std::fill(s->hstale.begin(), s->hstale.end(), 0);  // the caller has just set U
    if (s->gemm) return gemm_launch(s, 0, 0, nullptr, nullptr, nullptr, false);
#define CALL(NN) if (do_prepare<NN>(s)) return -1;
    CHS_FOR_N(s->N, CALL)
#undef CALL
    return 0;
}

template <int N>
static int do_begin(chs_solver* s) {
    using G = Geo<N>;
    KArgs a = base_args(s);
    a.nsims = s->batch;
    const dim3 block(G::NT);
    CHS_LAUNCH(k_begin, dim3((s->batch + 127) / 128), dim3(128), 0, s->stream, s->sims, s->batch);
    CHS_LAUNCH((k_row<N, ROW_FWD_U>), pgrid(s->cap_row[ROW_FWD_U], s->num_sms, G::NTILES, a.nsims), block, G::SMEM_BYTES, s->stream, a);    // U -> T
    CHS_LAUNCH((k_col<N, COL_FWD>), pgrid(s->cap_col[COL_FWD], s->num_sms, G::NTILES, a.nsims), block, G::SMEM_BYTES, s->stream, a);      // T -> hat_U
    CHS_LAUNCH((k_row<N, ROW_FWD_MU>), pgrid(s->cap_row[ROW_FWD_MU], s->num_sms, G::NTILES, a.nsims), block, G::SMEM_BYTES, s->stream, a);  // mu(U) -> T, pre-part
    s->launches += 4;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int chs_begin(chs_solver* s) {
    if (!s) return fail("chs_begin: null handle");
    if (s->gemm) {
        if (gemm_launch(s, 1, 0, nullptr, nullptr, nullptr, false)) return -1;
    } else {
#define CALL(NN) if (do_begin<NN>(s)) return -1;
        CHS_FOR_N(s->N, CALL)
#undef CALL
    }
    // the prologue's control step may already halt a sim (time limit); the index list is
    // compacted at the first chs_poll
    s->hindex.resize(s->batch);
    for (int i = 0; i < s->batch; ++i) s->hindex[i] = i;
    s->n_running = s->batch;
    CHS_CUDA(cudaMemcpyAsync(s->index, s->hindex.data(), sizeof(int) * s->batch, cudaMemcpyHostToDevice, s->stream));
    return 0;
}

template <int N>
static int do_steps(chs_solver* s, long long n_iters, const double* noise, const double* noise_mean, int last) {
    using G = Geo<N>;
    if (s->n_running <= 0 || n_iters <= 0) return 0;
    KArgs a = base_args(s);
    a.sim_index = s->index;                       // identity list until the first compaction
    a.nsims = s->n_running;
    const dim3 block(G::NT);
    // programmatic dependent launch of the two step kernels (not while per-kernel events are recorded)
    static const bool pdl_env = [] { const char* e = getenv("CHS_PDL"); return !e || atoi(e) != 0; }();
    // Only when the launch (nearly) fills the CTA slots: the dependents of a programmatic launch become resident while
    // their predecessor still runs, and with a partly filled GPU they pile up on the SMs the predecessor left free --
    // measured on B200, N=512: 2..4 simulations 47 us/step with, 29..36 us without (profiles/r2c_experiments.md, section 4).
    // The _LL launches (at most one tile per SM, each CTA alone on its SM by its shared-memory footprint) cannot pile up.
    static const int pdl_min_pct = [] { const char* e = getenv("CHS_PDL_MIN_PCT"); return e ? atoi(e) : 80; }();
    auto pdl_for = [&](int nsims, bool exclusive) {
        return pdl_env && !s->timing && (exclusive || (long long)G::NTILES * nsims * 100 >= (long long)pdl_min_pct * s->cap_row[ROW_STEP] ||
                                         2LL * G::NTILES * nsims <= s->num_sms);
    };
    const bool pdl = pdl_for(s->n_running, false);                     // (mixed launches: whole batch)
    if (s->n_running == s->batch) a.sim_index = nullptr;
    // without noise the step kernels keep the field in spectral form only (U is materialised by chs_end);
    // with noise k_row stores the jittered U every step
    for (int i = 0; i < s->n_running; ++i) s->hstale[s->hindex[i]] = noise ? 0 : 1;
    // ---- mixed launches: the running simulations are split into halves A and B, B half a step behind A:
    //   col(A,0) | row(A,0)+col(B,0) | col(A,1)+row(B,0) | ... | row(A,K-1)+col(B,K-1) | row(B,K-1)
    // = 2K-1 k_mix launches with column and row CTAs resident together + one half-size launch at each end;
    // the call still ends with every simulation exactly K steps further.
    const bool mix = !noise && s->mix_mode != 0 && s->n_running >= (s->mix_mode > 0 ? 2 : CHS_MIX_MIN_SIMS) && n_iters >= 2;
    if (mix) {
        const int nA = (s->n_running + 1) / 2, nB = s->n_running - nA;
        KArgs A = a, B = a;
        A.sim_index = s->index; A.nsims = nA;                 // (the identity list is resident until the first compaction)
        B.sim_index = s->index + nA; B.nsims = nB;
        static const int period_env = [] { const char* e = getenv("CHS_MIX_PERIOD"); return e ? atoi(e) : 0; }();   // tuning experiment
        const int period = period_env > 0 ? period_env : ((s->num_sms > 1) ? (s->num_sms & ~1) : 2);
        const dim3 gA = pgrid(s->cap_col[COL_STEP], s->num_sms, G::NTILES, nA), gB = pgrid(s->cap_row[ROW_STEP], s->num_sms, G::NTILES, nB);
        const dim3 gmix = pgrid(s->cap_mix, s->num_sms, 2 * G::NTILES, nA);
        const size_t ev0 = s->mix_used;
        if (s->timing) cudaEventRecord(next_mix_event(s), s->stream);
        A.last = 0;
        if (pdl) CHS_LAUNCH_PDL((k_col<N, COL_STEP>), gA, block, G::SMEM_BYTES, s->stream, A);
        else CHS_LAUNCH((k_col<N, COL_STEP>), gA, block, G::SMEM_BYTES, s->stream, A);
        if (s->timing) cudaEventRecord(next_mix_event(s), s->stream);
        for (long long it = 0; it < n_iters; ++it) {
            const int lastf = (last && it == n_iters - 1) ? 1 : 0;
            // row(A, it) + col(B, it)
            A.last = lastf; B.last = 0;
            if (pdl) CHS_LAUNCH_PDL((k_mix<N>), gmix, block, G::SMEM_BYTES, s->stream, B, A, period);
            else CHS_LAUNCH((k_mix<N>), gmix, block, G::SMEM_BYTES, s->stream, B, A, period);
            if (s->timing) cudaEventRecord(next_mix_event(s), s->stream);
            if (it + 1 < n_iters) {                          // col(A, it+1) + row(B, it)
                A.last = 0; B.last = 0;
                if (pdl) CHS_LAUNCH_PDL((k_mix<N>), gmix, block, G::SMEM_BYTES, s->stream, A, B, period);
                else CHS_LAUNCH((k_mix<N>), gmix, block, G::SMEM_BYTES, s->stream, A, B, period);
                if (s->timing) cudaEventRecord(next_mix_event(s), s->stream);
            }
        }
        B.last = last ? 1 : 0;
        if (pdl) CHS_LAUNCH_PDL((k_row<N, ROW_STEP>), gB, block, G::SMEM_BYTES, s->stream, B);
        else CHS_LAUNCH((k_row<N, ROW_STEP>), gB, block, G::SMEM_BYTES, s->stream, B);
        if (s->timing) {
            cudaEventRecord(next_mix_event(s), s->stream);
            s->mix_runs.push_back({ev0, (size_t)(2 * n_iters + 1)});
            s->mix_iters += n_iters;
            if (s->mix_used > 60000 && drain_events(s)) return -1;
        }
        s->launches += 2 * n_iters + 1;
        CHS_CUDA(cudaGetLastError());
        return 0;
    }
    // ---- L2 blocking: the running simulations are stepped in sub-batches of `sub` simulations, each through ALL
    // n_iters iterations before the next one starts: hat_U + T of a sub-batch (2 * 8 N^2 bytes per simulation)
    // stay resident in the 126 MB L2 from one iteration to the next, so only the first and the last iteration of
    // a call move the state through HBM.  Simulations are independent, so the order does not change any result.
    // few simulations (fewer tiles than SMs): the low-latency build of the two kernels (chs_ll.cu)
    const bool ll = s->ll_max > 0 && s->n_running <= s->ll_max && N == 512 && sizeof(KArgs) == (size_t)chs_ll_kargs_size() &&
                    chs_ll_init(N) == 0;
    const int sub_cfg = s->sub_sims;
    const int sub = (sub_cfg > 0 && sub_cfg < s->n_running && n_iters >= 2) ? sub_cfg : s->n_running;
    for (int sb0 = 0; sb0 < s->n_running; sb0 += sub) {
    KArgs b = a;
    if (sub < s->n_running) {
        b.sim_index = s->index + sb0;             // (the identity list is resident until the first compaction)
        b.nsims = (s->n_running - sb0 < sub) ? s->n_running - sb0 : sub;
    }
    // at most one tile per SM: the unrolled / register-rich instantiations of the two kernels (COL_STEP_LL, ROW_STEP_LL)
    // ... and up to two tiles per SM still run them (two CTAs of 244 / 255 registers fill an SM's register file, so no
    // third CTA -- e.g. an early dependent -- can join), with the tile's own shared memory instead of the exclusive amount
    // (measured: N=512 x 3..4 members 33.6 -> 30.2 us/step, N=256 x 5..8 26.6 -> 23.9, N=128 x 10..16 26.7 -> 21.7, N=64 x 20..36
    // 25.6 -> 21.7; not N = 1024: its 256-thread tiles fit only one such CTA per SM, 2 members 48.9 -> 55.1 us)
    static const int ll_tiles_env = [] { const char* e = getenv("CHS_LL_TILES_PER_SM"); return e ? atoi(e) : 0; }();
    const int ll_tiles = ll_tiles_env > 0 ? ll_tiles_env : (G::NT <= 128 ? 2 : 1);
    const long long ctas_b = (long long)G::NTILES * b.nsims;
    const bool one_per_sm = s->one_mode >= 0 ? s->one_mode != 0 : (ctas_b <= (long long)ll_tiles * s->num_sms);
    // two per SM: 100 KB each, so that a third never fits (tiles of fewer than 128 threads would fit by registers)
    const int ll_smem = (ctas_b <= (long long)s->num_sms) ? ll_smem_bytes(G::SMEM_BYTES)
                                                          : (G::SMEM_BYTES > 100 * 1024 ? G::SMEM_BYTES : 100 * 1024);
    const bool pdl = pdl_for(b.nsims, one_per_sm && !ll);
    const dim3 gcol_b = pgrid(s->cap_col[COL_STEP], s->num_sms, G::NTILES, b.nsims), grow_b = pgrid(s->cap_row[ROW_STEP], s->num_sms, G::NTILES, b.nsims);
    const dim3 grid_b(G::NTILES, b.nsims);
    for (long long it = 0; it < n_iters; ++it) {
        b.last = (last && it == n_iters - 1) ? 1 : 0;
        if (noise) {
            b.noise = noise + (size_t)it * N * N;
            b.noise_mean = noise_mean + it;
        }
        if (s->timing) {
            if (s->ev_used + 4 > 65536 && drain_events(s)) return -1;
            cudaEventRecord(next_event(s), s->stream);
        }
        if (ll) { if (chs_ll_launch(N, 0, &b, b.nsims, pdl ? 1 : 0, (void*)s->stream)) return fail("chs_steps: low-latency launch failed"); }
        else if (one_per_sm) {
            if (pdl) CHS_LAUNCH_PDL((k_col<N, COL_STEP_LL>), gcol_b, block, ll_smem, s->stream, b);
            else CHS_LAUNCH((k_col<N, COL_STEP_LL>), gcol_b, block, ll_smem, s->stream, b);
        }
        else if (pdl) CHS_LAUNCH_PDL((k_col<N, COL_STEP>), gcol_b, block, G::SMEM_BYTES, s->stream, b);
        else CHS_LAUNCH((k_col<N, COL_STEP>), gcol_b, block, G::SMEM_BYTES, s->stream, b);
        if (s->timing) cudaEventRecord(next_event(s), s->stream);
        if (ll) { if (chs_ll_launch(N, 1, &b, b.nsims, pdl ? 1 : 0, (void*)s->stream)) return fail("chs_steps: low-latency launch failed"); }
        else if (one_per_sm) {
            if (pdl) CHS_LAUNCH_PDL((k_row<N, ROW_STEP_LL>), grow_b, block, ll_smem, s->stream, b);
            else CHS_LAUNCH((k_row<N, ROW_STEP_LL>), grow_b, block, ll_smem, s->stream, b);
        }
        else if (pdl) CHS_LAUNCH_PDL((k_row<N, ROW_STEP>), grow_b, block, G::SMEM_BYTES, s->stream, b);
        else CHS_LAUNCH((k_row<N, ROW_STEP>), grow_b, block, G::SMEM_BYTES, s->stream, b);
        if (s->timing) cudaEventRecord(next_event(s), s->stream);
        s->launches += 2;
        if (noise) {
            if (pdl) CHS_LAUNCH_PDL((k_diag<N, DIAG_JITTER>), grid_b, block, G::SMEM_BYTES, s->stream, b);
            else CHS_LAUNCH((k_diag<N, DIAG_JITTER>), grid_b, block, G::SMEM_BYTES, s->stream, b);
            s->launches += 1;
        }
        if (s->timing) cudaEventRecord(next_event(s), s->stream);
    }
    }
    if (s->timing) s->t_iters += n_iters;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int chs_steps(chs_solver* s, int64_t n_iters, const double* noise, const double* noise_mean, int32_t last) {
    if (!s || n_iters < 0) return fail("chs_steps: bad argument");
    if (noise && !noise_mean) return fail("chs_steps: noise without noise_mean");
    if (s->gemm) return n_iters > 0 ? gemm_launch(s, 2, n_iters, noise, nullptr, nullptr, true) : 0;
#define CALL(NN) if (do_steps<NN>(s, n_iters, noise, noise_mean, last)) return -1;
    CHS_FOR_N(s->N, CALL)
#undef CALL
    return 0;
}

extern "C" int chs_poll(chs_solver* s, int32_t* stop_reason, int64_t* computed_steps, int64_t* rows_written) {
    if (!s) return fail("chs_poll: null handle");
    if (pull(s)) return -1;
    for (int i = 0; i < s->batch; ++i) {
        if (stop_reason) stop_reason[i] = s->hsims[i].stop_reason;
        if (computed_steps) computed_steps[i] = s->hsims[i].computed_steps;
        if (rows_written) rows_written[i] = s->hsims[i].rows_written;
    }
    if (refresh_index(s)) return -1;
    return s->n_running;
}

extern "C" int chs_rewind_rows(chs_solver* s) {
    if (!s) return fail("chs_rewind_rows: null handle");
    CHS_LAUNCH(k_rewind, dim3((s->batch + 127) / 128), dim3(128), 0, s->stream, s->sims, s->batch);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

template <int N>
static int do_end(chs_solver* s) {
    using G = Geo<N>;
    // host-tracked: which simulations stepped since their U was last valid.  No device read-back and no
    // synchronisation -- the inverse transform is just queued on the handle's stream.
    std::vector<int> stale;
    for (int i = 0; i < s->batch; ++i)
        if (s->hstale[i]) stale.push_back(i);
    std::fill(s->hstale.begin(), s->hstale.end(), 0);
    if (!stale.empty()) {
        CHS_CUDA(cudaMemcpyAsync(s->index, stale.data(), sizeof(int) * stale.size(), cudaMemcpyHostToDevice, s->stream));
        KArgs a = base_args(s);
        a.sim_index = s->index;
        a.nsims = (int)stale.size();
        const dim3 block(G::NT);
        CHS_LAUNCH((k_col<N, COL_INV>), pgrid(s->cap_col[COL_INV], s->num_sms, G::NTILES, a.nsims), block, G::SMEM_BYTES, s->stream, a);    // hat_U -> T
        CHS_LAUNCH((k_row<N, ROW_INV>), pgrid(s->cap_row[ROW_INV], s->num_sms, G::NTILES, a.nsims), block, G::SMEM_BYTES, s->stream, a);    // T -> U
        s->launches += 2;
        CHS_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int chs_end(chs_solver* s) {
    if (!s) return fail("chs_end: null handle");
    if (s->gemm) { CHS_CUDA(cudaStreamSynchronize(s->stream)); return 0; }     // U is written back by every launch
#define CALL(NN) if (do_end<NN>(s)) return -1;
    CHS_FOR_N(s->N, CALL)
#undef CALL
    return 0;
}

template <int N>
static int do_dctn(chs_solver* s, const double* in, double* out, bool inverse) {
    using G = Geo<N>;
    KArgs a = base_args(s);
    a.nsims = s->batch;
    const dim3 block(G::NT);
    if (!inverse) {
        a.src = in; a.dst = nullptr;
        CHS_LAUNCH((k_row<N, ROW_FWD_U>), pgrid(s->cap_row[ROW_FWD_U], s->num_sms, G::NTILES, a.nsims), block, G::SMEM_BYTES, s->stream, a);  // in -> T
        a.src = nullptr; a.dst = out; a.natural = 1;
        CHS_LAUNCH((k_col<N, COL_FWD>), pgrid(s->cap_col[COL_FWD], s->num_sms, G::NTILES, a.nsims), block, G::SMEM_BYTES, s->stream, a);    // T -> out
    } else {
        a.src = in; a.dst = nullptr; a.natural = 1;
        CHS_LAUNCH((k_col<N, COL_INV>), pgrid(s->cap_col[COL_INV], s->num_sms, G::NTILES, a.nsims), block, G::SMEM_BYTES, s->stream, a);    // in -> T
        a.src = nullptr; a.dst = out;
        CHS_LAUNCH((k_row<N, ROW_INV>), pgrid(s->cap_row[ROW_INV], s->num_sms, G::NTILES, a.nsims), block, G::SMEM_BYTES, s->stream, a);    // T -> out
    }
    s->launches += 2;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int chs_dctn(chs_solver* s, const double* in, double* out) {
    if (!s || !in || !out) return fail("chs_dctn: bad argument");
    if (s->gemm) return gemm_launch(s, 3, 0, nullptr, in, out, false);
#define CALL(NN) if (do_dctn<NN>(s, in, out, false)) return -1;
    CHS_FOR_N(s->N, CALL)
#undef CALL
    return 0;
}
extern "C" int chs_idctn(chs_solver* s, const double* in, double* out) {
    if (!s || !in || !out) return fail("chs_idctn: bad argument");
    if (s->gemm) return gemm_launch(s, 4, 0, nullptr, in, out, false);
#define CALL(NN) if (do_dctn<NN>(s, in, out, true)) return -1;
    CHS_FOR_N(s->N, CALL)
#undef CALL
    return 0;
}

extern "C" int chs_debug_log(chs_solver* s, const double* x, double* y, int64_t n) {
    if (!s || !x || !y || n < 0) return fail("chs_debug_log: bad argument");
    if (n == 0) return 0;
    CHS_LAUNCH(k_debug_log, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s->stream, x, y, (long long)n, s->logtab);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int64_t chs_launch_count(const chs_solver* s) { return s ? s->launches : 0; }

// =======================================================================================
//  Slab path (large single domain; see chs_slab.cuh and chsimpy_b200/slab.py)
// =======================================================================================
#include "chs_slab.cuh"
#include "chs_big.cuh"

struct chs_slab {
    int device, N, rows, row_base, world, rank;
    double *U, *rowsbuf;
    long long rows_cap;
    cudaStream_t stream;
    Sim* sim;
    double* part;        // [R_NVAL][ntiles] (also reused by prepare: [4][PREP_BLOCKS])
    double* part_ge;     // [UPD_BLOCKS]
    double* yedge;       // [1]
    double* vec;         // [R_NVAL]
    double2 *tw, *om, *logtab;
    double *lam, *gsin;
    double2* lamg;
    int* kof;
    Sim hsim;
    long long launches;
    int upd_used;
};
static const int SLAB_UPD_BLOCKS = 1184, SLAB_PREP_BLOCKS = 1184;

// sizes the FFT kernels do not take run on the GEMM-based arbitrary-N path (chs_big.cuh): one rank, any row count
static bool fft_slab_supports(int N) {
    return N == 64 || N == 128 || N == 256 || N == 512 || N == 1024 || N == 2048 || N == 4096 || N == 8192 || N == 16384;
}
static bool big_mode(int N) { return !fft_slab_supports(N) && N >= 8 && N <= 2048; }
static int slab_lines(int N) { return big_mode(N) ? 1 : geo_lines(N); }

struct SlabLayout { size_t sim, part, part_ge, yedge, vec, tw, om, lam, gsin, lamg, kof, logtab, total; };
static SlabLayout slab_layout(int N, int rows) {
    SlabLayout L; size_t o = 0;
    const int ntiles = rows / slab_lines(N);
    const size_t npart = (size_t)R_NVAL * ntiles > (size_t)4 * SLAB_PREP_BLOCKS ? (size_t)R_NVAL * ntiles : (size_t)4 * SLAB_PREP_BLOCKS;
    L.sim = o; o = align_up(o + sizeof(Sim));
    L.part = o; o = align_up(o + sizeof(double) * npart);
    L.part_ge = o; o = align_up(o + sizeof(double) * (size_t)(ntiles > SLAB_UPD_BLOCKS ? ntiles : SLAB_UPD_BLOCKS));
    L.yedge = o; o = align_up(o + sizeof(double) * 2);
    L.vec = o; o = align_up(o + sizeof(double) * R_NVAL);
    L.tw = o; o = align_up(o + sizeof(double2) * (size_t)(N / 2));
    L.om = o; o = align_up(o + sizeof(double2) * (size_t)(N + N / 4));      // om[N] + the contiguous copy of om[4k], k < N/4
    L.lam = o; o = align_up(o + sizeof(double) * (size_t)N);
    L.gsin = o; o = align_up(o + sizeof(double) * (size_t)N);
    L.lamg = o; o = align_up(o + sizeof(double2) * (size_t)(3 * (N / 4) + 3));
    L.kof = o; o = align_up(o + sizeof(int) * (size_t)N);
    L.logtab = o; o = align_up(o + sizeof(double2) * (size_t)LOG_TABLE_N);
    L.total = o;
    return L;
}

extern "C" int32_t chs_slab_supports_n(int32_t N) { return fft_slab_supports(N) ? 1 : 0; }
extern "C" int32_t chs_big_supports_n(int32_t N) { return big_mode(N) ? 1 : 0; }
extern "C" int32_t chs_slab_row_granularity(int32_t N) { return chs_slab_supports_n(N) ? slab_lines(N) : -1; }
extern "C" int64_t chs_slab_workspace_bytes(int32_t N, int32_t rows) {
    if ((!chs_slab_supports_n(N) && !big_mode(N)) || rows < 1 || rows % slab_lines(N)) return -1;
    return (int64_t)slab_layout(N, rows).total;
}

#define CHS_FOR_SLAB_N(N_, CALL)                \
    switch (N_) {                               \
        case 64: { CALL(64); } break;           \
        case 128: { CALL(128); } break;         \
        case 256: { CALL(256); } break;         \
        case 512: { CALL(512); } break;         \
        case 1024: { CALL(1024); } break;       \
        case 2048: { CALL(2048); } break;       \
        case 4096: { CALL(4096); } break;       \
        case 8192: { CALL(8192); } break;       \
        case 16384: { CALL(16384); } break;     \
        default: return fail("unsupported N for the slab path"); \
    }

template <int N>
static int slab_set_attrs() {
    const int b = Geo<N>::SMEM_BYTES;
    CHS_CUDA(cudaFuncSetAttribute(k_slab_row<N, S_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_slab_row<N, S_MU>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_slab_row<N, S_INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_slab_row<N, S_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_slab_row<N, S_YFWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    CHS_CUDA(cudaFuncSetAttribute(k_slab_row<N, S_YSTEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    return 0;
}

extern "C" chs_slab* chs_slab_create(int32_t device, int32_t N, int32_t rows, int32_t row_base, int32_t world, int32_t rank,
                                     const chs_params* p, double* U, double* rowsbuf, int64_t rows_cap,
                                     void* workspace, int64_t workspace_bytes, const double* lambda_host, void* stream) {
    const bool big = big_mode(N);
    if ((!chs_slab_supports_n(N) && !big) || rows < 1 || rows % slab_lines(N) || !p) { fail("chs_slab_create: bad N/rows"); return nullptr; }
    if (big && (world != 1 || rows != N)) { fail("chs_slab_create: the arbitrary-N path runs on one rank"); return nullptr; }
    const SlabLayout L = slab_layout(N, rows);
    if (workspace_bytes < (int64_t)L.total) { fail("chs_slab_create: workspace too small"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail("chs_slab_create: cudaSetDevice failed"); return nullptr; }
    chs_slab* s = new chs_slab();
    s->device = device; s->N = N; s->rows = rows; s->row_base = row_base; s->world = world; s->rank = rank;
    s->U = U; s->rowsbuf = rowsbuf; s->rows_cap = rows_cap; s->stream = (cudaStream_t)stream;
    unsigned char* w = (unsigned char*)workspace;
    s->sim = (Sim*)(w + L.sim); s->part = (double*)(w + L.part); s->part_ge = (double*)(w + L.part_ge);
    s->yedge = (double*)(w + L.yedge); s->vec = (double*)(w + L.vec);
    s->tw = (double2*)(w + L.tw); s->om = (double2*)(w + L.om); s->lam = (double*)(w + L.lam);
    s->gsin = (double*)(w + L.gsin); s->kof = (int*)(w + L.kof); s->logtab = (double2*)(w + L.logtab);
    s->lamg = (double2*)(w + L.lamg);
    s->launches = 0; s->upd_used = 0;
    std::memset(&s->hsim, 0, sizeof(Sim));
    s->hsim.p = *p; s->hsim.delt = p->delt; s->hsim.delt_coef = p->delt;
    sim_derive(s->hsim);
    const int M = N / 2;
    std::vector<double2> tw(M), om(N + N / 4), lt(LOG_TABLE_N);
    std::vector<double> gs(N);
    std::vector<int> kof(N);
    const long double pi = 3.14159265358979323846264338327950288L;
    if (!big) {                                     // FFT tables (powers of two only)
        fill_twiddles(N, tw);
        fill_om(N, om);
        fill_kof(N, kof);
    }
    for (int k = 0; k < N; ++k) { const long double sn = sinl(pi * k / N); gs[k] = (double)(sn * sn); }
    for (int i = 0; i < LOG_TABLE_N; ++i) {
        const unsigned long long b0 = LOG_OFF + ((unsigned long long)i << LOG_SHIFT), b1 = LOG_OFF + ((unsigned long long)(i + 1) << LOG_SHIFT);
        double z0, z1; std::memcpy(&z0, &b0, 8); std::memcpy(&z1, &b1, 8);
        const double invc = (double)(1.0L / ((long double)z0 * 0.5L + (long double)z1 * 0.5L));
        lt[i] = make_double2(invc, (double)(-logl((long double)invc)));
    }
    bool ok = cudaMemsetAsync(workspace, 0, L.total, s->stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(s->tw, tw.data(), sizeof(double2) * M, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(s->om, om.data(), sizeof(double2) * om.size(), cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(s->gsin, gs.data(), sizeof(double) * N, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(s->kof, kof.data(), sizeof(int) * N, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(s->logtab, lt.data(), sizeof(double2) * LOG_TABLE_N, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok &= cudaMemcpyAsync(s->lam, lambda_host, sizeof(double) * N, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    std::vector<double2> lamg;
    if (!big) {
        ok &= upload_const_tables(N, tw, om, s->stream) == 0;
        fill_lamg(N, lambda_host, gs, lamg);
        ok &= cudaMemcpyAsync(s->lamg, lamg.data(), sizeof(double2) * lamg.size(), cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    }
    ok &= cudaMemcpyAsync(s->sim, &s->hsim, sizeof(Sim), cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
    ok &= cudaStreamSynchronize(s->stream) == cudaSuccess;
    int rc = 0;
#define CALL(NN) rc = slab_set_attrs<NN>();
    if (!big) switch (N) {
        case 64: CALL(64) break; case 128: CALL(128) break; case 256: CALL(256) break; case 512: CALL(512) break;
        case 1024: CALL(1024) break; case 2048: CALL(2048) break; case 4096: CALL(4096) break;
        case 8192: CALL(8192) break; case 16384: CALL(16384) break;
    }
#undef CALL
    if (!ok || rc) { if (ok) {} else fail("chs_slab_create: table upload failed"); delete s; return nullptr; }
    return s;
}
extern "C" void chs_slab_destroy(chs_slab* s) { delete s; }
extern "C" double* chs_slab_vec(chs_slab* s) { return s ? s->vec : nullptr; }
extern "C" int64_t chs_slab_launch_count(const chs_slab* s) { return s ? s->launches : 0; }
extern "C" int chs_slab_set_stream(chs_slab* s, void* stream) {
    if (!s) return fail("chs_slab_set_stream: null handle");
    s->stream = (cudaStream_t)stream;
    return 0;
}

template <int N>
static int slab_row(chs_slab* s, int mode, const double* src, double* dst, int rows, int row_base, int diag, double mean_u,
                    double* H = nullptr, const double* noise = nullptr, const double* noise_mean = nullptr) {
    using G = Geo<N>;
    SlabArgs a;
    // a launch may cover a sub-range of the rank's rows (pipelined exchange): row_base is global
    const int local0 = row_base - s->row_base;
    if (local0 < 0 || local0 % G::LINES || local0 + rows > s->rows) return fail("chs_slab_row: rows outside the slab");
    a.src = src; a.dst = dst; a.Uout = s->U + (size_t)local0 * N; a.rows = rows; a.row_base = row_base; a.diag = diag; a.mean_u = mean_u;
    a.tile0 = local0 / G::LINES; a.tiles_total = s->rows / G::LINES;
    a.part = s->part; a.S = s->sim; a.tw = s->tw; a.om = s->om; a.logtab = s->logtab;
    a.noise = noise ? noise + (size_t)local0 * N : nullptr; a.noise_mean = noise_mean;
    a.H = H; a.part_ge = s->part_ge; a.lam = s->lam; a.gsin = s->gsin; a.kof = s->kof; a.lamg = s->lamg;
    const int ntiles = rows / G::LINES;
#ifdef CHS_EMU
    const dim3 grid(ntiles < 3 ? ntiles : 3);
#else
    const dim3 grid(ntiles);
#endif
    const dim3 block(G::NT);
    switch (mode) {
        case S_FWD: CHS_LAUNCH((k_slab_row<N, S_FWD>), grid, block, G::SMEM_BYTES, s->stream, a); break;
        case S_MU: CHS_LAUNCH((k_slab_row<N, S_MU>), grid, block, G::SMEM_BYTES, s->stream, a); break;
        case S_INV: CHS_LAUNCH((k_slab_row<N, S_INV>), grid, block, G::SMEM_BYTES, s->stream, a); break;
        case S_STEP: CHS_LAUNCH_PDL((k_slab_row<N, S_STEP>), grid, block, G::SMEM_BYTES, s->stream, a); break;
        case S_YFWD: CHS_LAUNCH((k_slab_row<N, S_YFWD>), grid, block, G::SMEM_BYTES, s->stream, a); break;
        case S_YSTEP: CHS_LAUNCH_PDL((k_slab_row<N, S_YSTEP>), grid, block, G::SMEM_BYTES, s->stream, a); break;
        default: return fail("chs_slab_row: bad mode");
    }
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

// mode: 0 S_FWD (physical rows -> row DCT-II), 1 S_MU (U rows -> mu -> row DCT-II), 2 S_INV (row
// DCT-III -> physical rows), 3 S_STEP (row DCT-III -> U stored in the handle's U buffer ->
// diagnostics + mu -> row DCT-II).  `rows` rows starting at src/dst; row_base = global index.
extern "C" int chs_slab_row(chs_slab* s, int32_t mode, const double* src, double* dst, int32_t rows, int32_t row_base,
                            int32_t diag, double mean_u) {
    if (!s || !src || !dst || rows < 1 || rows % slab_lines(s->N)) return fail("chs_slab_row: bad argument");
    if (mode == S_YSTEP) return fail("chs_slab_row: mode 5 is chs_slab_update");
    // S_YFWD: dst is hat_U' (x-slot rows, natural ky columns)
#define CALL(NN) if (slab_row<NN>(s, mode, src, dst, rows, row_base, diag, mean_u, mode == S_YFWD ? dst : nullptr)) return -1;
    CHS_FOR_SLAB_N(s->N, CALL)
#undef CALL
    return 0;
}

// The x pass of a step (mode S_STEP) with the per-step jitter of solver.py:210-211: `noise` = this step's uniform
// draws for the rank's rows ([rows of the rank][N], device), `noise_mean` = device scalar, mean of the whole
// N x N draw (both NULL without jitter).
extern "C" int chs_slab_step_x(chs_slab* s, const double* src, double* dst, int32_t rows, int32_t row_base, double mean_u,
                               const double* noise, const double* noise_mean) {
    if (!s || !src || !dst || rows < 1 || rows % slab_lines(s->N)) return fail("chs_slab_step_x: bad argument");
    if ((noise == nullptr) != (noise_mean == nullptr)) return fail("chs_slab_step_x: noise and noise_mean go together");
#define CALL(NN) if (slab_row<NN>(s, S_STEP, src, dst, rows, row_base, 1, mean_u, nullptr, noise, noise_mean)) return -1;
    CHS_FOR_SLAB_N(s->N, CALL)
#undef CALL
    return 0;
}

// adaptive dt: colsum[x] = sum over this rank's rows of delt_max/sqrt(1 + 62.5 mu(U[y][x])^2) (solver.py:182-183);
// scratch: at least 16*N doubles.  The caller adds the ranks' vectors (all-reduce) and hands the result to
// chs_slab_control*_dyn.
extern "C" int chs_slab_colsum(chs_slab* s, double* colsum, double* scratch) {
    if (!s || !colsum || !scratch) return fail("chs_slab_colsum: bad argument");
    const int nch = s->rows >= 16 ? 16 : 1;
#ifdef CHS_EMU
    const int nt = 32;
#else
    const int nt = 128;
#endif
    CHS_LAUNCH(k_slab_colsum, dim3((s->N + nt - 1) / nt, nch), dim3(nt), 0, s->stream, (const double*)s->U, s->rows, s->N,
               (const Sim*)s->sim, (const double2*)s->logtab, scratch);
    CHS_LAUNCH(k_slab_colsum_final, dim3((s->N + nt - 1) / nt), dim3(nt), 0, s->stream, (const double*)scratch, nch, s->N, colsum);
    s->launches += 2;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

// jitter: stencil gradient energy of the stored field; top / bot = boundary rows of the neighbouring ranks
// (device, N doubles each; ignored at the domain edges but must be valid pointers).  The partial sums replace the
// spectral ones of the y pass in the following chs_slab_sums* call.
extern "C" int chs_slab_grad(chs_slab* s, const double* top, const double* bot) {
    if (!s || !top || !bot) return fail("chs_slab_grad: bad argument");
#ifdef CHS_EMU
    const int nb = 2, nt = 64;
#else
    const int nb = SLAB_UPD_BLOCKS, nt = 256;
#endif
    CHS_LAUNCH(k_slab_grad, dim3(nb), dim3(nt), nt * sizeof(double), s->stream, (const double*)s->U, top, bot, s->rows, s->row_base,
               s->N, s->part_ge);
    s->upd_used = nb;
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

// numpy PCG64 draws into `out` on the slab handle's stream (see chs_pcg64_fill), and row means
extern "C" int chs_slab_pcg64_fill(chs_slab* s, uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo,
                                   uint64_t offset, double* out, int64_t count) {
    if (!s || !out || count < 0) return fail("chs_slab_pcg64_fill: bad argument");
    if (count == 0) return 0;
    const long long threads = (count + PCG_RUN - 1) / PCG_RUN;
#ifdef CHS_EMU
    const int nt = 32;
#else
    const int nt = 128;
#endif
    CHS_LAUNCH(k_pcg64_fill, dim3((unsigned)((threads + nt - 1) / nt)), dim3(nt), 0, s->stream, out, (long long)count,
               (unsigned long long)state_hi, (unsigned long long)state_lo, (unsigned long long)inc_hi,
               (unsigned long long)inc_lo, (unsigned long long)offset);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int chs_slab_row_means(chs_slab* s, const double* in, int64_t rows, int64_t cols, double* out) {
    if (!s || !in || !out || rows < 1 || cols < 1) return fail("chs_slab_row_means: bad argument");
#ifdef CHS_EMU
    const int nt = 32;
#else
    const int nt = 1024;            // one block per row of the noise chunk: 64 blocks read 128 MiB -- the bytes in flight per block are what counts
#endif
    CHS_LAUNCH(k_row_means, dim3((unsigned)rows), dim3(nt), nt * sizeof(double), s->stream, in, (long long)cols, out);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

// 4096-element tiles + bulk stores (k_slab_transpose_bulk) when shape and alignment allow; CHS_SLAB_BULK=0 keeps the 32 x 32 kernels
static bool slab_bulk_transpose(int R, int C, const double* in, const double* out, int in_ld, int out_ld) {
    static const bool on = [] { const char* e = getenv("CHS_SLAB_BULK"); return !e || atoi(e) != 0; }();
    return on && R % SLAB_TR == 0 && C % SLAB_TC == 0 && in_ld % 2 == 0 && out_ld % 2 == 0 &&
           (uintptr_t)in % 16 == 0 && (uintptr_t)out % 16 == 0;
}

extern "C" int chs_slab_transpose(chs_slab* s, const double* in, double* out, int32_t R, int32_t C, int32_t in_ld, int32_t out_ld) {
    if (!s || !in || !out) return fail("chs_slab_transpose: bad argument");
    // (local transposes: measured 4 % of a step SLOWER with the bulk-store tiles on one GPU -- N=8192 1.93 -> 2.01 ms,
    // N=16384 9.34 -> 9.68 ms -- so the 32 x 33 tile kernel stays unless CHS_SLAB_BULK=2 asks for the experiment)
    static const bool bulk_local = [] { const char* e = getenv("CHS_SLAB_BULK"); return e && atoi(e) == 2; }();
    if (bulk_local && slab_bulk_transpose(R, C, in, out, in_ld, out_ld)) {
        PeerDst pd;
        for (int i = 0; i < 8; ++i) { pd.p[i] = nullptr; pd.ld[i] = 0; }
        pd.p[0] = out; pd.ld[0] = out_ld;
        CHS_LAUNCH_PDL(k_slab_transpose_bulk, dim3(C / SLAB_TC, R / SLAB_TR, 1), dim3(256), SLAB_TC * SLAB_TP * sizeof(double),
                       s->stream, pd, in, (int)R, (int)C, (int)in_ld, 0, 1);
    } else {
        CHS_LAUNCH_PDL(k_slab_transpose, dim3((C + 31) / 32, (R + 31) / 32), dim3(256), 32 * 33 * sizeof(double), s->stream,
                       in, out, (int)R, (int)C, (int)in_ld, (int)out_ld);
    }
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

// One launch for the blocks of all P ranks: dst[p] = peer-mapped destination base for rank p (already offset
// to this rank's column range); the block for rank p is in[:, p*C .. p*C + C).
extern "C" int chs_slab_transpose_peers(chs_slab* s, const double* in, const uint64_t* dst, int32_t R, int32_t C,
                                        int32_t in_ld, int32_t out_ld) {
    if (!s || !in || !dst || s->world < 1 || s->world > 8) return fail("chs_slab_transpose_peers: bad argument (at most 8 ranks)");
    PeerPtrs pp;
    for (int i = 0; i < 8; ++i) pp.p[i] = (i < s->world) ? (double*)(uintptr_t)dst[i] : nullptr;
    bool bulk = slab_bulk_transpose(R, C, in, pp.p[0], in_ld, out_ld);
    for (int i = 1; i < s->world; ++i) bulk = bulk && ((uintptr_t)pp.p[i] % 16 == 0);
    if (bulk) {
        PeerDst pd;
        for (int i = 0; i < 8; ++i) { pd.p[i] = pp.p[i]; pd.ld[i] = out_ld; }
        CHS_LAUNCH_PDL(k_slab_transpose_bulk, dim3(C / SLAB_TC, R / SLAB_TR, s->world), dim3(256), SLAB_TC * SLAB_TP * sizeof(double),
                       s->stream, pd, in, (int)R, (int)C, (int)in_ld, s->rank, s->world);
    } else
        CHS_LAUNCH_PDL(k_slab_transpose_peers, dim3((C + 31) / 32, (R + 31) / 32, s->world), dim3(256), 32 * 33 * sizeof(double),
                       s->stream, pp, in, (int)R, (int)C, (int)in_ld, (int)out_ld, s->rank, s->world);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

// Copy-engine exchange (slab.py, CHS_SLAB_CE): the transposed blocks of `R` local rows (a row chunk) go to a local
// staging buffer -- block p = [C][R] row-major at stage + p*C*R -- except this rank's own block, which goes straight to
// `own` (leading dimension own_ld); chs_slab_copy_blocks then moves block p into rank p's buffer as ONE pitched
// device-to-device copy per peer on `copy_stream`: the copy engines drive NVLink while the SMs transform the next chunk.
extern "C" int chs_slab_transpose_stage(chs_slab* s, const double* in, double* stage, double* own, int32_t own_ld,
                                        int32_t R, int32_t C, int32_t in_ld) {
    if (!s || !in || !stage || !own || s->world < 1 || s->world > 8) return fail("chs_slab_transpose_stage: bad argument");
    if (R % SLAB_TR || C % SLAB_TC || in_ld % 2 || own_ld % 2 || (uintptr_t)in % 16 || (uintptr_t)stage % 16 || (uintptr_t)own % 16)
        return fail("chs_slab_transpose_stage: shape / alignment");
    PeerDst pd;
    for (int i = 0; i < 8; ++i) { pd.p[i] = nullptr; pd.ld[i] = 0; }
    for (int i = 0; i < s->world; ++i) { pd.p[i] = stage + (size_t)i * C * R; pd.ld[i] = R; }
    pd.p[s->rank] = own; pd.ld[s->rank] = own_ld;
    CHS_LAUNCH_PDL(k_slab_transpose_bulk, dim3(C / SLAB_TC, R / SLAB_TR, s->world), dim3(256), SLAB_TC * SLAB_TP * sizeof(double),
                   s->stream, pd, in, (int)R, (int)C, (int)in_ld, s->rank, s->world);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int chs_slab_copy_blocks(chs_slab* s, const uint64_t* dst, int64_t dst_pitch_bytes, const double* stage,
                                    int32_t R, int32_t C, void* copy_stream) {
    if (!s || !dst || !stage || s->world < 1 || s->world > 8) return fail("chs_slab_copy_blocks: bad argument");
    for (int i = 1; i < s->world; ++i) {                  // nearest peer first, like the SM-driven exchange
        const int p = (s->rank + i) % s->world;
#ifdef CHS_EMU
        for (int c = 0; c < C; ++c)
            std::memcpy((char*)(uintptr_t)dst[p] + (size_t)c * dst_pitch_bytes, stage + (size_t)p * C * R + (size_t)c * R, (size_t)R * 8);
        (void)copy_stream;
#else
        CHS_CUDA(cudaMemcpy2DAsync((void*)(uintptr_t)dst[p], (size_t)dst_pitch_bytes, stage + (size_t)p * C * R, (size_t)R * 8,
                                   (size_t)R * 8, (size_t)C, cudaMemcpyDeviceToDevice, (cudaStream_t)copy_stream));
#endif
    }
    return 0;
}

// The whole y pass of a step on `rows` local x-slot rows (global slot index slot_base + r), B holding
// the transposed x-transformed mu (physical y along the row) on entry:
//   H = (H + Seig*rowDCT(B))/CHeig ;  B = rowIDCT(H)      (one kernel, one read of B and H, one write of each)
extern "C" int chs_slab_update(chs_slab* s, double* H, double* B, int32_t rows, int32_t slot_base) {
    if (!s || !H || !B || rows < 1 || rows % slab_lines(s->N)) return fail("chs_slab_update: bad argument");
#define CALL(NN) if (slab_row<NN>(s, S_YSTEP, B, B, rows, slot_base, 0, 0.0, H)) return -1;
    CHS_FOR_SLAB_N(s->N, CALL)
#undef CALL
    s->upd_used = s->rows / slab_lines(s->N);
    return 0;
}

// y-edge rows of the stored U (device pointers to row 0/1 or N-2/N-1 of the global field)
extern "C" int chs_slab_yedge(chs_slab* s, const double* r0, const double* r1, int32_t accumulate) {
    if (!s || !r0 || !r1) return fail("chs_slab_yedge: bad argument");
    CHS_LAUNCH(k_slab_yedge, dim3(1), dim3(256), 256 * sizeof(double), s->stream, r0, r1, s->N, s->yedge, (int)accumulate);
    s->launches += 1;
    return 0;
}
extern "C" int chs_slab_clear_yedge(chs_slab* s) {
    if (!s) return fail("chs_slab_clear_yedge: null handle");
    CHS_CUDA(cudaMemsetAsync(s->yedge, 0, sizeof(double) * 2, s->stream));
    return 0;
}

// local sums -> vec (to be all-reduced over ranks by the caller when world > 1)
extern "C" int chs_slab_reduce(chs_slab* s, int32_t rows, int32_t with_update) {
    if (!s) return fail("chs_slab_reduce: null handle");
    CHS_LAUNCH(k_slab_reduce, dim3(1), dim3(32 * R_NVAL), 32 * R_NVAL * sizeof(double), s->stream, (const double*)s->part, (int)(rows / slab_lines(s->N)),
               (const double*)s->part_ge, with_update ? s->upd_used : 0, (const double*)s->yedge, s->vec);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

// the sums of one step in a single launch: per-tile partials + spectral gradient energy + the y-edge
// terms of the stored field (top_edge: this rank holds rows 0/1 of the domain, bottom_edge: rows N-2/N-1)
static int slab_sums(chs_slab* s, int32_t top_edge, int32_t bottom_edge, const uint64_t* peer_slots) {
    const size_t n = (size_t)s->N;
    const double* t0 = top_edge ? s->U : nullptr;
    const double* b0 = bottom_edge ? s->U + (size_t)(s->rows - 2) * n : nullptr;
    PeerPtrs pp;
    for (int i = 0; i < 8; ++i) pp.p[i] = (peer_slots && i < s->world) ? (double*)(uintptr_t)peer_slots[i] : nullptr;
    CHS_LAUNCH_PDL(k_slab_sums, dim3(1), dim3(128 * (R_NVAL + 1)), 128 * (R_NVAL + 1) * sizeof(double), s->stream,
               (const double*)s->part, (int)(s->rows / slab_lines(s->N)), (const double*)s->part_ge, s->upd_used,
               t0, t0 ? t0 + n : nullptr, b0, b0 ? b0 + n : nullptr, s->N, s->vec, pp, peer_slots ? s->world : 0);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int chs_slab_sums(chs_slab* s, int32_t top_edge, int32_t bottom_edge) {
    if (!s) return fail("chs_slab_sums: null handle");
    return slab_sums(s, top_edge, bottom_edge, nullptr);
}
// the same, and the 7 sums are also stored into peer_slots[r] (8 doubles in rank r's gather buffer, peer-mapped)
extern "C" int chs_slab_sums_peers(chs_slab* s, int32_t top_edge, int32_t bottom_edge, const uint64_t* peer_slots) {
    if (!s || !peer_slots || s->world > 8) return fail("chs_slab_sums_peers: bad argument");
    return slab_sums(s, top_edge, bottom_edge, peer_slots);
}

extern "C" int chs_slab_prepare(chs_slab* s, const double* U_halo, double mean_u) {
    if (!s || !U_halo) return fail("chs_slab_prepare: bad argument");
#ifdef CHS_EMU
    const int nb = 2, nt = 64;
#else
    const int nb = SLAB_PREP_BLOCKS, nt = 256;
#endif
    CHS_LAUNCH(k_slab_prepare, dim3(nb), dim3(nt), 4 * nt * sizeof(double), s->stream, U_halo, s->rows, s->row_base, s->N,
               mean_u, (const Sim*)s->sim, (const double2*)s->logtab, s->part);
    CHS_LAUNCH(k_slab_reduce_prepare, dim3(1), dim3(32), 0, s->stream, (const double*)s->part, nb, s->N, s->vec);
    s->launches += 2;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

// post: 0 prologue (only ||mu||^2 + pre part), 1 end of an iteration, 2 prepare (row 0)
static int slab_control(chs_slab* s, int32_t last, int32_t post, const double* allvec, const double* colsum) {
    CHS_LAUNCH_PDL(k_slab_control, dim3(1), dim3(32), 0, s->stream, s->sim, s->vec, s->rowsbuf, s->rows_cap, s->N,
               (int)last, (int)post, allvec, s->world, colsum);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int chs_slab_control(chs_slab* s, int32_t last, int32_t post) {
    if (!s) return fail("chs_slab_control: null handle");
    return slab_control(s, last, post, nullptr, nullptr);
}
// the rank sums come from the gather buffer allvec[world][8] (filled by every rank's chs_slab_sums_peers)
extern "C" int chs_slab_control_gathered(chs_slab* s, int32_t last, int32_t post, const double* allvec) {
    if (!s || !allvec) return fail("chs_slab_control_gathered: bad argument");
    return slab_control(s, last, post, allvec, nullptr);
}
// adaptive dt: colsum = the all-rank column sums of chs_slab_colsum (N doubles, device) when this control step
// is followed by an iteration that updates delt (solver.py:177-181), else NULL; allvec may be NULL
extern "C" int chs_slab_control_dyn(chs_slab* s, int32_t last, int32_t post, const double* allvec, const double* colsum) {
    if (!s) return fail("chs_slab_control_dyn: null handle");
    return slab_control(s, last, post, allvec, colsum);
}

extern "C" int chs_slab_begin(chs_slab* s) {
    if (!s) return fail("chs_slab_begin: null handle");
    CHS_LAUNCH(k_begin, dim3(1), dim3(32), 0, s->stream, s->sim, 1);
    s->launches += 1;
    return 0;
}
extern "C" int chs_slab_rewind_rows(chs_slab* s) {
    if (!s) return fail("chs_slab_rewind_rows: null handle");
    CHS_LAUNCH(k_rewind, dim3(1), dim3(32), 0, s->stream, s->sim, 1);
    s->launches += 1;
    return 0;
}

extern "C" int chs_slab_get_state(chs_slab* s, chs_state* st, int64_t* rows_written, int32_t* halted) {
    if (!s || !st) return fail("chs_slab_get_state: bad argument");
    CHS_CUDA(cudaMemcpyAsync(&s->hsim, s->sim, sizeof(Sim), cudaMemcpyDeviceToHost, s->stream));
    CHS_CUDA(cudaStreamSynchronize(s->stream));
    const Sim& h = s->hsim;
    st->delt = h.delt; st->time_delta_sum = h.time_delta_sum; st->time_passed = h.time_passed;
    st->tau0 = h.tau0; st->t0 = h.t0; st->computed_steps = h.computed_steps;
    st->skip_check = h.skip_check; st->stop_reason = h.stop_reason;
    if (rows_written) *rows_written = h.rows_written;
    if (halted) *halted = h.halted;
    return 0;
}
extern "C" int chs_slab_set_state(chs_slab* s, const chs_state* st) {
    if (!s || !st) return fail("chs_slab_set_state: bad argument");
    CHS_CUDA(cudaMemcpyAsync(&s->hsim, s->sim, sizeof(Sim), cudaMemcpyDeviceToHost, s->stream));
    CHS_CUDA(cudaStreamSynchronize(s->stream));
    Sim& h = s->hsim;
    h.delt = st->delt; h.time_delta_sum = st->time_delta_sum; h.time_passed = st->time_passed;
    h.tau0 = st->tau0; h.t0 = st->t0; h.computed_steps = st->computed_steps;
    h.skip_check = st->skip_check; h.stop_reason = st->stop_reason;
    CHS_CUDA(cudaMemcpyAsync(s->sim, &h, sizeof(Sim), cudaMemcpyHostToDevice, s->stream));
    CHS_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

// ---- device-side numpy PCG64 stream (jitter noise) -------------------------------------------
// ---- arbitrary-N path (chs_big.cuh): stage-level calls on a slab handle created for a size the FFT kernels do
// not take; every matrix operand is n8 x n8 (N rounded up to a multiple of 8), row-major with pitch ld, zero padded
static int big_blocks(const chs_slab* s) {
#ifdef CHS_EMU
    return 2;
#else
    return s->N < 592 ? s->N : 592;
#endif
}
extern "C" int chs_big_gemm(chs_slab* s, const double* A, const double* B, double* D, int32_t n8, int32_t ld) {
    if (!s || !A || !B || !D || n8 < 8 || n8 % 8 || ld < n8) return fail("chs_big_gemm: bad argument");
    const int nb = (n8 + BIG_TILE - 1) / BIG_TILE;
    CHS_LAUNCH(k_big_gemm, dim3(nb, nb), dim3(256), (BIG_TILE * (BIG_KS + 1) + BIG_KS * (BIG_TILE + 1)) * sizeof(double), s->stream,
               A, B, D, (int)n8, (int)ld);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int chs_big_update(chs_slab* s, double* H, const double* Mh, int32_t ld) {
    if (!s || !H || !Mh) return fail("chs_big_update: bad argument");
    CHS_LAUNCH(k_big_update, dim3(big_blocks(s)), dim3(256), 0, s->stream, H, Mh, (const double*)s->lam, (const Sim*)s->sim, s->N, (int)ld);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int chs_big_copy(chs_slab* s, const double* src, int32_t sld, double* dst, int32_t dld, int32_t respect_halt) {
    if (!s || !src || !dst) return fail("chs_big_copy: bad argument");
    CHS_LAUNCH(k_big_copy, dim3(big_blocks(s)), dim3(256), 0, s->stream, src, (int)sld, dst, (int)dld, s->N, (const Sim*)s->sim, (int)respect_halt);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}
// physics of the new field (Up, pitch ld; or the stored U when from_U): [jitter ->] U, diagnostics partials, mu -> A
extern "C" int chs_big_phys(chs_slab* s, const double* Up, double* A, int32_t ld, double mean_u, const double* noise,
                            const double* noise_mean, int32_t diag, int32_t from_U) {
    if (!s || !A || (!Up && !from_U)) return fail("chs_big_phys: bad argument");
#ifdef CHS_EMU
    const int nt = 32;
#else
    const int nt = 256;
#endif
    CHS_LAUNCH(k_big_phys, dim3(big_blocks(s)), dim3(nt), 5 * nt * sizeof(double), s->stream, Up, s->U, A, s->N, (int)ld, s->sim,
               (const double2*)s->logtab, mean_u, noise, noise_mean, (int)diag, (int)from_U, s->part);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}
// the 7 sums of a step from the k_big_phys partials (+ the stencil gradient partials of chs_slab_grad)
extern "C" int chs_big_sums(chs_slab* s) {
    if (!s) return fail("chs_big_sums: null handle");
    PeerPtrs pp;
    for (int i = 0; i < 8; ++i) pp.p[i] = nullptr;
    CHS_LAUNCH(k_slab_sums, dim3(1), dim3(128 * (R_NVAL + 1)), 128 * (R_NVAL + 1) * sizeof(double), s->stream,
               (const double*)s->part, big_blocks(s), (const double*)s->part_ge, s->upd_used,
               (const double*)nullptr, (const double*)nullptr, (const double*)nullptr, (const double*)nullptr, s->N, s->vec, pp, 0);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int chs_pcg64_fill(chs_solver* s, uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo,
                              uint64_t offset, double* out, int64_t count) {
    if (!s || !out || count < 0) return fail("chs_pcg64_fill: bad argument");
    if (count == 0) return 0;
    const long long threads = (count + PCG_RUN - 1) / PCG_RUN;
#ifdef CHS_EMU
    const int nt = 32;
#else
    const int nt = 128;
#endif
    CHS_LAUNCH(k_pcg64_fill, dim3((unsigned)((threads + nt - 1) / nt)), dim3(nt), 0, s->stream, out, (long long)count,
               (unsigned long long)state_hi, (unsigned long long)state_lo, (unsigned long long)inc_hi,
               (unsigned long long)inc_lo, (unsigned long long)offset);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int chs_lcg_fill(double* out, int32_t n1, int32_t n2, double seed, void* stream) {
    if (!out || n1 < 1 || n2 < 1) return fail("chs_lcg_fill: bad argument");
    CHS_LAUNCH(k_lcg_fill, dim3(1), dim3(1), 0, (cudaStream_t)stream, out, (int)n1, (int)n2, seed);
    CHS_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int chs_row_means(chs_solver* s, const double* in, int64_t rows, int64_t cols, double* out) {
    if (!s || !in || !out || rows < 1 || cols < 1) return fail("chs_row_means: bad argument");
#ifdef CHS_EMU
    const int nt = 32;
#else
    const int nt = 1024;            // one block per row of the noise chunk: 64 blocks read 128 MiB -- the bytes in flight per block are what counts
#endif
    CHS_LAUNCH(k_row_means, dim3((unsigned)rows), dim3(nt), nt * sizeof(double), s->stream, in, (long long)cols, out);
    s->launches += 1;
    CHS_CUDA(cudaGetLastError());
    return 0;
}
