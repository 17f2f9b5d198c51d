// Host emulation of the tiny CUDA subset chs_kernels.cuh uses (TEST HARNESS ONLY, see
// chs_rt.h).  One OS thread per CUDA thread of a block, blocks run one after another.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct double2 { double x, y; };
static inline double2 make_double2(double a, double b) { return double2{a, b}; }

namespace emu {
inline thread_local dim3 t_threadIdx, t_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline unsigned char* g_smem = nullptr;
inline std::barrier<>* g_bar = nullptr;
inline thread_local std::barrier<>* t_wbar = nullptr;     // the 32 threads of this thread's warp

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
    g_blockDim = block;
    g_gridDim = grid;
    const unsigned nthreads = block.x * block.y * block.z;
    // generous slack: the emulated block reductions use NT*NV doubles of scratch
    std::vector<unsigned char> smem(smem_bytes + (size_t)nthreads * 16 * sizeof(double) + 64);
    g_smem = smem.data();
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                std::barrier<> bar((std::ptrdiff_t)nthreads);
                g_bar = &bar;
                std::vector<std::unique_ptr<std::barrier<>>> wbar;
                for (unsigned w0 = 0; w0 < nthreads; w0 += 32)
                    wbar.emplace_back(new std::barrier<>((std::ptrdiff_t)(nthreads - w0 < 32 ? nthreads - w0 : 32)));
                std::vector<std::thread> th;
                th.reserve(nthreads);
                for (unsigned t = 0; t < nthreads; ++t)
                    th.emplace_back([&, t] {
                        t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        t_blockIdx = dim3(bx, by, bz);
                        t_wbar = wbar[t / 32].get();
                        body();
                    });
                for (auto& x : th) x.join();
            }
    g_smem = nullptr;
    g_bar = nullptr;
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define __syncthreads() emu::g_bar->arrive_and_wait()
#define __syncwarp() emu::t_wbar->arrive_and_wait()
#define __restrict__
#define __launch_bounds__(...)
#define CHS_DEV static inline
#define CHS_MEM inline
#define CHS_HD static inline
#define CHS_KERNEL static
#define CHS_CX
#define CHS_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch(grid, block, smem, [&] { kern(__VA_ARGS__); })
#define CHS_LAUNCH_PDL(kern, grid, block, smem, stream, ...) CHS_LAUNCH(kern, grid, block, smem, stream, __VA_ARGS__)
#define CHS_PDL_TRIGGER()
#define CHS_PDL_WAIT()
#define CHS_SMEM_DECL
#define CHS_SMEM_PTR (emu::g_smem)

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double __dmul_rn(double a, double b) { return a * b; }   // built with -ffp-contract=off
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicSub(int* p, int v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }

// ---- runtime API subset -------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return 0; }
typedef int cudaEvent_t;
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = 0; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }
static inline void chs_cp_async16(void* d, const void* s) { std::memcpy(d, s, 16); }
static inline void chs_cp_async8(void* d, const void* s) { std::memcpy(d, s, 8); }
static inline void chs_cp_async_wait_all() {}
// bulk copies: synchronous memcpy in the host build (the mbarrier is a no-op)
static inline void chs_mbar_init(void*, unsigned) {}
static inline void chs_mbar_expect_tx(void*, unsigned) {}
#define chs_mbar_wait(bar, parity) emu::g_bar->arrive_and_wait()      /* the copying thread arrives after its memcpy */
static inline void chs_bulk_g2s(void* d, const void* s, unsigned n, void*) { std::memcpy(d, s, n); }
static inline void chs_bulk_s2g(void* d, const void* s, unsigned n) { std::memcpy(d, s, n); }
static inline void chs_fence_async_smem() {}
static inline void chs_bulk_commit_wait() {}
#define CHS_SYNCWARP() emu::t_wbar->arrive_and_wait()
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount };
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 3; return 0; }   // 3 "SMs": exercises the persistent loops
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 1; return 0; }
