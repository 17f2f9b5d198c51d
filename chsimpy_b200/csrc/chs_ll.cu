// Low-latency build of the two step kernels: the SAME sources as chs_api.cu's (chs_kernels.cuh), compiled a
// second time as namespace chs_ll with 8 instead of 16 complex points per thread and FFT stage -- 32 threads
// per line, 256 threads (8 warps) per tile.  The global data layouts (pair-major T, tile-major hat_U, partial
// sums, Sim) do not depend on that split, so chs_steps may use either build from one call to the next.
//
// Why: a single N=512 simulation (BASELINE configs[1]) is 64 tiles per kernel -- fewer than the 148 SMs -- and its
// step time is the serial latency of ONE tile (2 x ~15 us with 4 warps per tile, one per scheduler).  Twice the
// warps per tile halve the dependent work per thread and give every scheduler two warps to interleave.
// Supported: N = 512 (the FFT plan must end with the radix-4 stage: one pairing unit per thread).
#define CHS_NS chs_ll
#define CHS_PPT 8
#include "chs_kernels.cuh"

#include <cstring>

using namespace chs_ll;

#ifndef CHS_EMU
#define CHS_LL_HIDDEN __attribute__((visibility("hidden")))
#else
#define CHS_LL_HIDDEN
#endif

static bool g_ll_ready = false;

// 0 on success; the attribute calls need the device of the handle to be current
extern "C" CHS_LL_HIDDEN int chs_ll_init(int N) {
    if (N != 512) return -1;
    if (g_ll_ready) return 0;
    const int b = Geo<512>::SMEM_BYTES;
    if (cudaFuncSetAttribute(k_col<512, COL_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, b) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_row<512, ROW_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, b) != cudaSuccess) return -1;
    g_ll_ready = true;
    return 0;
}

// which: 0 = k_col<STEP>, 1 = k_row<STEP>; kargs = the caller's KArgs image (same layout in both namespaces)
extern "C" CHS_LL_HIDDEN int chs_ll_launch(int N, int which, const void* kargs, int nsims, int pdl, void* stream) {
    if (N != 512 || !g_ll_ready) return -1;
    using G = Geo<512>;
    KArgs a;
    std::memcpy(&a, kargs, sizeof(KArgs));
    const dim3 block(G::NT);
#if defined(CHS_EMU)
    const long long total = (long long)G::NTILES * nsims;
    const dim3 grid((unsigned)(total < 3 ? total : 3));          // the emulated tile loop covers the rest
#else
    const dim3 grid((unsigned)(G::NTILES * nsims));
#endif
    cudaStream_t st = (cudaStream_t)stream;
    if (which == 0) {
        if (pdl) CHS_LAUNCH_PDL((k_col<512, COL_STEP>), grid, block, G::SMEM_BYTES, st, a);
        else CHS_LAUNCH((k_col<512, COL_STEP>), grid, block, G::SMEM_BYTES, st, a);
    } else {
        if (pdl) CHS_LAUNCH_PDL((k_row<512, ROW_STEP>), grid, block, G::SMEM_BYTES, st, a);
        else CHS_LAUNCH((k_row<512, ROW_STEP>), grid, block, G::SMEM_BYTES, st, a);
    }
    return 0;
}

extern "C" CHS_LL_HIDDEN int chs_ll_kargs_size(void) { return (int)sizeof(KArgs); }
