// Runtime shim: the kernels in chs_kernels.cuh are written once and compiled either by
// nvcc for sm_100a (the product) or -- with -DCHS_EMU -- by g++ for the host, where
// every CUDA thread of a block becomes an OS thread and __syncthreads() a barrier.
// The host build exists ONLY so that the index logic of the kernels can be unit-tested
// in the GPU-less build container (tests/, via tests/emu_lib.py); the Python package
// never loads it and has no CPU fallback.
#pragma once

// The kernel headers are compiled twice into the library: as namespace chs with 16 points per thread and FFT
// stage (the throughput geometry, chs_api.cu) and as namespace chs_ll with 8 (twice the threads per line: the
// low-latency geometry of a single simulation / a few simulations, chs_ll.cu).
#ifndef CHS_NS
#define CHS_NS chs
#endif
#ifndef CHS_PPT
#define CHS_PPT 16
#endif

#ifdef CHS_EMU
#include "emu.h"
#else
#include <cuda_runtime.h>
#define CHS_DEV __device__ __forceinline__
#define CHS_MEM __device__ __forceinline__
#define CHS_HD __host__ __device__ __forceinline__
#define CHS_KERNEL __global__
#define CHS_CX __host__ __device__
#define CHS_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is
// still running; it must execute CHS_PDL_WAIT() before it touches anything the predecessor (or, by
// transitivity, any earlier kernel) writes.  Hides the launch + scheduling latency between the two
// dependent kernels of a step (matters for small batches, where a kernel runs for 15-20 us).
#define CHS_LAUNCH_PDL(kern, grid_, block_, smem_, stream_, ...)                                      \
    do {                                                                                           \
        cudaLaunchConfig_t cfg_ = {};                                                              \
        cfg_.gridDim = grid_; cfg_.blockDim = block_; cfg_.dynamicSmemBytes = smem_; cfg_.stream = stream_; \
        cudaLaunchAttribute at_[1];                                                                \
        at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                            \
        at_[0].val.programmaticStreamSerializationAllowed = 1;                                     \
        cfg_.attrs = at_; cfg_.numAttrs = 1;                                                       \
        cudaLaunchKernelEx(&cfg_, kern, __VA_ARGS__);                                                     \
    } while (0)
#define CHS_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define CHS_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define CHS_SMEM_DECL extern __shared__ __align__(16) unsigned char chs_smem_raw[];
#define CHS_SMEM_PTR (chs_smem_raw)
// asynchronous global->shared copies (LDGSTS): no register staging, all of a tile in flight
__device__ __forceinline__ void chs_cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src));
}
__device__ __forceinline__ void chs_cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src));
}
__device__ __forceinline__ void chs_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// ---- bulk asynchronous copies (the TMA unit, SASS UBLKCP): one thread moves a contiguous block between
// global and shared memory without touching the LSU / L1 data pipe.  Load: arm the mbarrier with the byte
// count, issue, every consumer waits on the barrier's phase.  Store: make the generic-proxy writes to shared
// memory visible to the async proxy, issue, wait until the source has been read.
__device__ __forceinline__ void chs_mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void chs_mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void chs_mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void chs_bulk_g2s(void* dst_smem, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void chs_bulk_s2g(void* dst, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"((unsigned)__cvta_generic_to_shared(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void chs_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// (.read: the shared-memory source may be reused / the CTA may exit; the writes themselves complete before the
// grid does -- the epilogue convention of TMA stores)
__device__ __forceinline__ void chs_bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;\ncp.async.bulk.wait_group.read 0;" ::: "memory");
}
#define CHS_SYNCWARP() __syncwarp()
#endif
