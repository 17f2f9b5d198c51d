// Plain C++ client of the C ABI (include/chs_b200.h): no Python, no PyTorch.
// Runs ONE simulation the way chsimpy/solver.py does (prepare -> solve_or_resume(n)) with the
// reference's `-g lcg` initial field (chsimpy/mport.py:8-32, solver.py:66) and prints the
// TimeData rows.  tests/test_gpu_parity.py builds it with
//     g++ -std=c++17 -Iinclude examples/c_abi_demo.cpp -o c_abi_demo -I$CUDA/include -L$CUDA/lib64 -lcudart -ldl
// and compares its output with the frozen reference run tests/golden/n64_lcg_k100.npz.
//
// usage: c_abi_demo <libchs_b200.so> N steps seed RT BRT B A0 A1 Amr kappa_tilde
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime_api.h>
#include <dlfcn.h>

#include "chs_b200.h"

#define CK(x) do { if ((x) != cudaSuccess) { std::fprintf(stderr, "CUDA error at %s:%d\n", __FILE__, __LINE__); return 2; } } while (0)

template <class F> static F sym(void* h, const char* name) {
    void* p = dlsym(h, name);
    if (!p) { std::fprintf(stderr, "missing symbol %s\n", name); std::exit(3); }
    return reinterpret_cast<F>(p);
}

int main(int argc, char** argv) {
    if (argc != 12) { std::fprintf(stderr, "usage: %s lib N steps seed RT BRT B A0 A1 Amr kappa_tilde\n", argv[0]); return 1; }
    void* lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { std::fprintf(stderr, "%s\n", dlerror()); return 1; }
    const int N = std::atoi(argv[2]), steps = std::atoi(argv[3]);
    const double seed = std::atof(argv[4]);
    chs_params p = {};
    p.RT = std::atof(argv[5]); p.BRT = std::atof(argv[6]); p.B = std::atof(argv[7]);
    p.A0 = std::atof(argv[8]); p.A1 = std::atof(argv[9]); p.Amr = std::atof(argv[10]); p.kappa_tilde = std::atof(argv[11]);
    p.L = 2.0; p.delx = p.L / (N - 1);
    p.delt = 3e-8; p.delt_max = 9e-8; p.M_tilde = 1.71e-8; p.threshold = 0.875;
    p.time_limit_s = 0; p.jitter = 0; p.full_sim = 1; p.adaptive_time = 0;

    auto workspace_bytes = sym<int64_t (*)(int32_t, int32_t)>(lib, "chs_workspace_bytes");
    auto create = sym<chs_solver* (*)(int32_t, int32_t, int32_t, double*, double*, double*, double*, int64_t, void*, int64_t,
                                      const double*, void*)>(lib, "chs_create");
    auto set_params = sym<int (*)(chs_solver*, int32_t, const chs_params*)>(lib, "chs_set_params");
    auto prepare = sym<int (*)(chs_solver*, const double*)>(lib, "chs_prepare");
    auto begin = sym<int (*)(chs_solver*)>(lib, "chs_begin");
    auto do_steps = sym<int (*)(chs_solver*, int64_t, const double*, const double*, int32_t)>(lib, "chs_steps");
    auto poll = sym<int (*)(chs_solver*, int32_t*, int64_t*, int64_t*)>(lib, "chs_poll");
    auto end = sym<int (*)(chs_solver*)>(lib, "chs_end");
    auto destroy = sym<void (*)(chs_solver*)>(lib, "chs_destroy");
    auto last_error = sym<const char* (*)()>(lib, "chs_last_error");

    // U_init = c0 + 0.01*c0*lcg  (un-centred, solver.py:66); x <- (a*x + c) mod 2^31 in float64, column-major fill
    std::vector<double> U((size_t)N * N), lam(N);
    double x = seed, sum = 0;
    for (int c = 0; c < N; ++c)
        for (int r = 0; r < N; ++r) {
            x = std::fmod(1103515245.0 * x + 12345.0, 2147483648.0);
            U[(size_t)r * N + c] = 0.875 + 0.01 * 0.875 * (x / 2147483647.0);
        }
    for (double v : U) sum += v;
    const double mean = sum / ((double)N * N);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < N; ++k) lam[k] = 2.0 * std::cos(pi * k / (N - 1)) - 2.0;      // utils.py:34-36

    const size_t fb = sizeof(double) * N * N;
    const int64_t rows_cap = steps + 4, wb = workspace_bytes(N, 1);
    if (wb <= 0) { std::fprintf(stderr, "N=%d is not supported\n", N); return 1; }
    double *dU, *dH, *dT, *dR; void* dW;
    CK(cudaSetDevice(0));
    CK(cudaMalloc((void**)&dU, fb)); CK(cudaMalloc((void**)&dH, fb)); CK(cudaMalloc((void**)&dT, fb));
    CK(cudaMalloc((void**)&dR, sizeof(double) * rows_cap * CHS_NCOLS)); CK(cudaMalloc(&dW, wb));
    CK(cudaMemcpy(dU, U.data(), fb, cudaMemcpyHostToDevice));

    chs_solver* s = create(0, N, 1, dU, dH, dT, dR, rows_cap, dW, wb, lam.data(), nullptr);
    if (!s) { std::fprintf(stderr, "chs_create: %s\n", last_error()); return 1; }
    std::vector<double> rows((size_t)rows_cap * CHS_NCOLS);
    int rc = set_params(s, 0, &p);
    rc |= prepare(s, &mean);                                   // Solver.prepare(): row 0
    CK(cudaMemcpy(rows.data(), dR, sizeof(double) * CHS_NCOLS, cudaMemcpyDeviceToHost));
    for (int c = 0; c < CHS_NCOLS; ++c) std::printf("%.17g%c", rows[c], c + 1 < CHS_NCOLS ? ' ' : '\n');
    rc |= begin(s);                                            // solve_or_resume(steps): steps-1 iterations (quirk Q3)
    rc |= do_steps(s, steps - 1, nullptr, nullptr, 1);
    int32_t stop = 0; int64_t cs = 0, rw = 0;
    const int running = poll(s, &stop, &cs, &rw);
    rc |= end(s);
    if (rc || running < 0) { std::fprintf(stderr, "error: %s\n", last_error()); return 1; }
    CK(cudaMemcpy(rows.data(), dR, sizeof(double) * rw * CHS_NCOLS, cudaMemcpyDeviceToHost));
    for (int64_t r = 0; r < rw; ++r)
        for (int c = 0; c < CHS_NCOLS; ++c) std::printf("%.17g%c", rows[r * CHS_NCOLS + c], c + 1 < CHS_NCOLS ? ' ' : '\n');
    std::fprintf(stderr, "computed_steps=%lld stop_reason=%d rows=%lld\n", (long long)cs, (int)stop, (long long)(rw + 1));
    destroy(s);
    cudaFree(dU); cudaFree(dH); cudaFree(dT); cudaFree(dR); cudaFree(dW);
    return 0;
}
