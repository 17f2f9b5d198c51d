/*
 * chs_b200.h -- C ABI of libchs_b200.so, the B200-native (sm_100a) replacement for the
 * hot path of uncertaintyhub/chsimpy: the semi-implicit spectral Cahn-Hilliard stepper
 * (reference chsimpy/solver.py:84-252 + chsimpy/timedata.py + chsimpy/utils.py:34-49)
 * for one simulation or a batch of independent simulations (the A0/A1 ensemble of
 * chsimpy/experiment.py:84-126).
 *
 * The reference is pure Python and has no FFI seam of its own; the
 * narrowest seam is the `Solver` class.  Each entry point below names the reference
 * lines it stands in for.  Plain pointers and sizes only: the caller (PyTorch, in
 * chsimpy_b200/solver.py) owns every device buffer and passes raw device addresses;
 * the library owns no persistent device memory.
 *
 * Threading: a handle is single-owner (not thread-safe); all work of a handle is
 * issued on the one CUDA stream given at creation.  Return value: 0 on success,
 * negative on error (see chs_last_error()).
 */
#ifndef CHS_B200_H
#define CHS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHS_ABI_VERSION 1

/* stop reasons -- chsimpy/solver.py:132,198,246 ('None' | 'time-limit' | 'energy');
 * CHS_STOP_NAN is the device image of the AssertionError of chsimpy/timedata.py:10. */
enum { CHS_STOP_NONE = 0, CHS_STOP_ENERGY = 1, CHS_STOP_TIME = 2, CHS_STOP_NAN = 3 };

/* TimeData column order, chsimpy/timedata.py:8-9 */
enum { CHS_COL_IT = 0, CHS_COL_E, CHS_COL_E2, CHS_COL_SA, CHS_COL_DOMTIME, CHS_COL_RA,
       CHS_COL_L2, CHS_COL_PS, CHS_COL_DELT, CHS_NCOLS };

/* Per-simulation constants: chsimpy/parameters.py:24-64 and the derived scalars of
 * chsimpy/solution.py:25-50 (computed by the host exactly as the reference does). */
typedef struct chs_params {
    double RT, BRT, B, A0, A1;       /* R*T, B*R*T, B, Redlich-Kister A0(T), A1(T) */
    double Amr, kappa_tilde, L;      /* 1/Am, gradient-energy parameter, domain length */
    double delx;                     /* L/(N-1)  (quirk Q1) */
    double delt, delt_max;           /* initial / maximum time step */
    double M_tilde, threshold;       /* mobility factor; SA threshold */
    double time_limit_s;             /* params.time_max*60, <= 0: no limit */
    double jitter;                   /* 0: off; else amplitude in (0, 0.1) */
    int32_t full_sim, adaptive_time;
} chs_params;

/* Mutable per-simulation solver state that the reference keeps in Solver/Solution
 * attributes (chsimpy/solver.py:50-54,128-135). */
typedef struct chs_state {
    double delt, time_delta_sum, time_passed, tau0, t0;
    int64_t computed_steps;
    int32_t skip_check, stop_reason;
} chs_state;

typedef struct chs_solver chs_solver;   /* opaque */

/* Size in bytes of the device workspace the caller must provide for `batch` sims. */
int64_t chs_workspace_bytes(int32_t N, int32_t batch);

/* Supported N of the batched stepper: FFT path (powers of two 32..1024) or DCT-as-GEMM path on
 * the FP64 tensor cores (any N in 8..104).  chs_uses_gemm tells which one chs_create picks. */
int32_t chs_supports_n(int32_t N);
int32_t chs_uses_gemm(int32_t N, int32_t batch);

/* Creates a solver for `batch` independent N x N simulations on CUDA device `device`.
 * Buffers (device pointers, row-major, contiguous, owned by the caller):
 *   U      [batch][N][N]  concentration field           (Solution.U)
 *   hat_U  [batch][N][N]  its 2-D orthonormal DCT-II    (local `hat_U`, solver.py:159)
 *   T      [batch][N][N]  row/column intermediate
 *   rows   [batch][rows_cap][9]  TimeData rows written by the device since chs_begin
 *   workspace  chs_workspace_bytes(N, batch) bytes
 * `lambda_host[N]` (host) = 2*cos(pi*k/(N-1)) - 2, the 1-D Laplacian spectrum exactly as
 * the caller's numpy evaluates chsimpy/utils.py:34-36 (so the multipliers are bit-identical).
 * `stream` is a cudaStream_t (0 = default stream).  Stands in for Solver.__init__
 * (solver.py:45-57) minus the host-side initial-condition generators. */
chs_solver* chs_create(int32_t device, int32_t N, int32_t batch,
                       double* U, double* hat_U, double* T,
                       double* rows, int64_t rows_cap,
                       void* workspace, int64_t workspace_bytes,
                       const double* lambda_host, void* stream);
void chs_destroy(chs_solver*);

int chs_set_params(chs_solver*, int32_t sim, const chs_params*);
int chs_set_state(chs_solver*, int32_t sim, const chs_state*);
int chs_get_state(chs_solver*, int32_t sim, chs_state*);       /* synchronises the stream */

/* Solver.prepare(), solver.py:84-135: row 0 of TimeData (E, E2, PS, Ra of the U
 * currently in the U buffer; SA = L2 = domtime = 0) is written to rows[sim][0], and
 * computed_steps=1, tau0=t0=0, stop_reason=None.  delt / time_delta_sum / skip_check
 * are left alone (quirk Q16).  `mean_U[batch]` (host) = np.mean of each field. */
int chs_prepare(chs_solver*, const double* mean_U_host);

/* Entry of Solver.solve_or_resume, solver.py:158-163: hat_U = dctn(U, 'ortho') for
 * every simulation, multipliers reset to the initial delt (solver.py:151-152), and the
 * chemical potential of U staged for the first iteration.  Resets the TimeData write
 * cursor of every sim to 0. */
int chs_begin(chs_solver*);

/* `n_iters` iterations of the loop body solver.py:165-249 for every simulation that
 * has not stopped; nothing synchronises with the host (device-side stop flags).
 *   noise       NULL, or device pointer to [n_iters][N][N] uniform [0,1) draws (the
 *               host-precomputed per-step jitter, solver.py:210-211), shared by all sims
 *   noise_mean  NULL, or device pointer to [n_iters] means of those draws
 *   last        non-zero if the reference's `for` ends after these iterations (the
 *               "pre" part of the following iteration -- adaptive dt, time accounting,
 *               time-limit test, solver.py:177-199 -- is then NOT run ahead) */
int chs_steps(chs_solver*, int64_t n_iters, const double* noise, const double* noise_mean, int32_t last);

/* Non-blocking-ish poll: copies per-sim stop reasons, computed_steps and the number of
 * TimeData rows written since chs_begin to host arrays of length batch (any may be
 * NULL); returns the number of simulations still running, or <0 on error.
 * Synchronises the handle's stream. */
int chs_poll(chs_solver*, int32_t* stop_reason, int64_t* computed_steps, int64_t* rows_written);

/* Sets the TimeData write cursor of every simulation back to 0 (the caller has copied the
 * rows out); lets an arbitrarily long run reuse a rows buffer of rows_cap rows. */
int chs_rewind_rows(chs_solver*);

/* Exit of solve_or_resume (solver.py:251): materialises U = idctn(hat_U) in the U
 * buffer for every sim whose U is stale (no-jitter runs never store U per step).  The inverse transform
 * is queued on the handle's stream; the call does not synchronise (which simulations stepped is tracked on
 * the host). */
int chs_end(chs_solver*);

/* Stand-alone transforms on [batch][N][N] device arrays (in -> out, T is scratch):
 * scipy.fftpack.dctn / idctn with norm='ortho' as called at solver.py:159,201,208. */
int chs_dctn(chs_solver*, const double* in, double* out);
int chs_idctn(chs_solver*, const double* in, double* out);

/* Device-side numpy PCG64: out[i] = the (offset+i)-th `Generator(PCG64).random()` draw after the
 * generator state {state, inc} (128-bit values as hi/lo words, from `bit_generator.state`), bit
 * for bit.  Replaces the host-side `rng.random((N, N))` of the per-step jitter (solver.py:78-79,
 * 210-211) so that no noise crosses PCIe.  chs_row_means: out[r] = mean(in[r, :]). */
int chs_pcg64_fill(chs_solver*, uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo,
                   uint64_t offset, double* out, int64_t count);
int chs_row_means(chs_solver*, const double* in, int64_t rows, int64_t cols, double* out);
/* The reference's float64 LCG initial-condition generator (chsimpy/mport.py:8-32, pinned by the known-answer
 * vector of reference tests/test.py:19-37): out[n1][n2] (row-major, device) = matlab_lcg_sample(n1, n2, seed).
 * The recurrence is serial in float64 (the rounding of a*x is part of the specification): one device thread. */
int chs_lcg_fill(double* out, int32_t n1, int32_t n2, double seed, void* stream);

/* Self-test hook: y[i] = the device's table-driven natural log of x[i] (csrc/fastlog.cuh),
 * the routine that stands in for np.log at solver.py:173,220.  Device pointers. */
int chs_debug_log(chs_solver*, const double* x, double* y, int64_t n);

/* Number of kernels this handle has launched since creation (bench.py gpu_launches). */
int64_t chs_launch_count(const chs_solver*);

/* Optional per-kernel timing for bench.py: when enabled, chs_steps brackets every kernel
 * with CUDA events on the handle's stream.  chs_get_timing synchronises and returns the
 * accumulated device milliseconds {column kernel, row kernel, jitter diagnostics kernel}
 * and the number of iterations they cover, then clears the accumulators. */
int chs_set_timing(chs_solver*, int32_t enable);
int chs_get_timing(chs_solver*, double* ms3, int64_t* n_iters);
/* The same for calls that ran as mixed launches (k_mix: column half-step of one half of the batch + row
 * half-step of the other half in ONE launch, see chs_steps): ms2/n2 = {accumulated ms, launches} of
 * {the mixed launches, the two half-size launches at the ends of every call}; n_iters = iterations covered. */
int chs_get_timing_mix(chs_solver*, double* ms2, int64_t* n2, int64_t* n_iters);
/* Scheduling of chs_steps (results are identical either way): -1 = mixed launches when at least 32 simulations
 * run and the call covers >= 2 iterations without noise (default; environment CHS_MIX overrides the default),
 * 0 = always one column + one row launch per iteration, 1 = mixed whenever >= 2 simulations run. */
int chs_set_mix(chs_solver*, int32_t mode);

/* ---------------------------------------------------------------------------------------
 * Slab path: ONE large N x N simulation, row-slab decomposed over `world` ranks (world = 1:
 * a single GPU), N in {64 .. 16384}.  Stage-level entry points; the host layer
 * (chsimpy_b200/slab.py) sequences them; for world > 1 it calls chs_slab_transpose once per peer
 * with `out` pointing INTO the peer's buffer (symmetric memory mapped over NVLink; NCCL all-to-all
 * when peer mapping is unavailable) and all-reduces the 7 diagnostic sums.  Reference: the same loop body
 * chsimpy/solver.py:165-249; the reference has no counterpart for the decomposition. */
typedef struct chs_slab chs_slab;
int32_t chs_slab_supports_n(int32_t N);
int32_t chs_slab_row_granularity(int32_t N);           /* local row counts must be multiples of this */
int64_t chs_slab_workspace_bytes(int32_t N, int32_t rows);
chs_slab* chs_slab_create(int32_t device, int32_t N, int32_t rows, int32_t row_base, int32_t world, int32_t rank,
                          const chs_params*, double* U /*[rows][N]*/, double* rows_buf, int64_t rows_cap,
                          void* workspace, int64_t workspace_bytes, const double* lambda_host, void* stream);
void chs_slab_destroy(chs_slab*);
/* mode 0: physical rows -> row DCT-II (slot order); 1: U rows -> mu -> row DCT-II; 2: row DCT-III ->
 * physical rows; 3: row DCT-III -> U (stored in the handle's U buffer) -> diagnostics + mu -> row DCT-II;
 * 4: physical rows -> row DCT-II in natural frequency order (dst = hat_U' rows, solver.py:159) */
int chs_slab_row(chs_slab*, int32_t mode, const double* src, double* dst, int32_t rows, int32_t row_base,
                 int32_t diag, double mean_u);
int chs_slab_transpose(chs_slab*, const double* in, double* out, int32_t R, int32_t C, int32_t in_ld, int32_t out_ld);
/* the exchange of one pass in ONE launch (world <= 8): block p of `in` (columns p*C .. p*C+C) is transposed
 * straight into dst[p], the peer-mapped destination in rank p's buffer */
int chs_slab_transpose_peers(chs_slab*, const double* in, const uint64_t* dst /*[world] device addresses*/, int32_t R, int32_t C,
                             int32_t in_ld, int32_t out_ld);
/* Copy-engine exchange: chs_slab_transpose_stage transposes the blocks of `R` local rows (leading dimension in_ld,
 * block for rank p = columns p*C .. p*C+C) into a local staging buffer (block p = [C][R] at stage + p*C*R), this
 * rank's own block directly into `own` (leading dimension own_ld); chs_slab_copy_blocks queues one pitched
 * device-to-device copy per peer on `copy_stream` (dst[p] = peer-mapped destination of block p, row pitch in bytes). */
int chs_slab_transpose_stage(chs_slab*, const double* in, double* stage, double* own, int32_t own_ld,
                             int32_t R, int32_t C, int32_t in_ld);
int chs_slab_copy_blocks(chs_slab*, const uint64_t* dst, int64_t dst_pitch_bytes, const double* stage,
                         int32_t R, int32_t C, void* copy_stream);
/* the y pass of one step in one kernel: H = (H + Seig*rowDCT(B))/CHeig; B = rowIDCT(H)  (solver.py:201-208) */
int chs_slab_update(chs_slab*, double* H, double* B, int32_t rows, int32_t slot_base);
int chs_slab_yedge(chs_slab*, const double* row_a, const double* row_b, int32_t accumulate);
int chs_slab_clear_yedge(chs_slab*);
int chs_slab_reduce(chs_slab*, int32_t rows, int32_t with_update);    /* local sums -> chs_slab_vec() */
/* reduce + both yedge calls of one step in a single launch (top_edge / bottom_edge: this rank holds rows
 * 0,1 / N-2,N-1 of the domain) */
int chs_slab_sums(chs_slab*, int32_t top_edge, int32_t bottom_edge);
/* ... and stores the 7 sums into peer_slots[r] (8 doubles of rank r's gather buffer) for every rank r: the
 * cross-rank reduction then needs no collective, chs_slab_control_gathered adds the slots in rank order */
int chs_slab_sums_peers(chs_slab*, int32_t top_edge, int32_t bottom_edge, const uint64_t* peer_slots /*[world]*/);
double* chs_slab_vec(chs_slab*);                                      /* device pointer, 7 doubles */
int chs_slab_prepare(chs_slab*, const double* U_with_halo /*[rows+2][N]*/, double mean_u);   /* solver.py:84-127 */
int chs_slab_control(chs_slab*, int32_t last, int32_t post);          /* solver.py:195-199, 230-249 */
int chs_slab_control_gathered(chs_slab*, int32_t last, int32_t post, const double* allvec /*[world][8]*/);
/* --adaptive-time and --jitter on the slab path (solver.py:177-193, 210-211):
 *   chs_slab_colsum       per-column sums of delt_max/sqrt(1+62.5 mu^2) over the rank's rows -> colsum[N] (scratch >= 16 N)
 *   chs_slab_control_dyn  control step that also applies the adaptive-dt update from the all-rank colsum (or NULL)
 *   chs_slab_step_x       x pass with this step's noise rows + the mean of the whole draw (device scalar)
 *   chs_slab_grad         np.gradient stencil energy of the stored jittered field (neighbour boundary rows given)
 *   chs_slab_pcg64_fill / chs_slab_row_means   numpy PCG64 draws / row means on the slab handle's stream */
int chs_slab_colsum(chs_slab*, double* colsum, double* scratch);
/* Arbitrary-N path (the reference accepts any N, cli_parser.py:27): for sizes neither a power of two nor <= 104,
 * chs_slab_create builds a one-rank handle whose transforms are FP64 tensor-core GEMMs C.X.C^T / C^T.Y.C
 * (csrc/chs_big.cuh); chsimpy_b200/slab.py (BigEngine) sequences the stage calls.  Operands are n8 x n8 matrices
 * (N rounded up to a multiple of 8, zero padded), row-major with pitch ld. */
int32_t chs_big_supports_n(int32_t N);
int chs_big_gemm(chs_slab*, const double* A, const double* B, double* D, int32_t n8, int32_t ld);
int chs_big_update(chs_slab*, double* H, const double* Mh, int32_t ld);                 /* solver.py:201-206 */
int chs_big_copy(chs_slab*, const double* src, int32_t sld, double* dst, int32_t dld, int32_t respect_halt);
int chs_big_phys(chs_slab*, const double* Up, double* A, int32_t ld, double mean_u, const double* noise,
                 const double* noise_mean, int32_t diag, int32_t from_U);               /* solver.py:166-175, 210-228 */
int chs_big_sums(chs_slab*);
int chs_slab_control_dyn(chs_slab*, int32_t last, int32_t post, const double* allvec, const double* colsum);
int chs_slab_step_x(chs_slab*, const double* src, double* dst, int32_t rows, int32_t row_base, double mean_u,
                    const double* noise, const double* noise_mean);
int chs_slab_grad(chs_slab*, const double* top, const double* bot);
int chs_slab_pcg64_fill(chs_slab*, uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo,
                        uint64_t offset, double* out, int64_t count);
int chs_slab_row_means(chs_slab*, const double* in, int64_t rows, int64_t cols, double* out);
int chs_slab_begin(chs_slab*);
int chs_slab_rewind_rows(chs_slab*);
int chs_slab_get_state(chs_slab*, chs_state*, int64_t* rows_written, int32_t* halted);
int chs_slab_set_state(chs_slab*, const chs_state*);
int64_t chs_slab_launch_count(const chs_slab*);
/* stream of the following launches (the pipelined exchange issues the transposes of a row chunk on a
 * second stream while the next chunk is transformed) */
int chs_slab_set_stream(chs_slab*, void* stream);

const char* chs_last_error(void);
int32_t chs_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
