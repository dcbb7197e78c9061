// Shared helpers for the stemseg_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/stemseg_b200.h"

namespace stemseg {

// ---- error reporting behind the C ABI (thread-local message, negative return codes) ------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what, const char* file, int line);

#define SS_CUDA_OK(call)                                                              \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess) return ::stemseg::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define SS_REQUIRE(cond, ...)                      \
    do {                                           \
        if (!(cond)) {                             \
            ::stemseg::set_error(__VA_ARGS__);     \
            return STEMSEG_ERR_INVALID_ARGUMENT;   \
        }                                          \
    } while (0)

int device_sm_count();          // cached cudaDevAttrMultiProcessorCount of the current device
int require_sm100();            // 0 if the current device is compute capability 10.x, else error code

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Grid-wide barrier for cooperatively launched kernels (all CTAs co-resident). `counter` is zeroed before the
// launch and only ever grows; `generation` is the 1-based index of this barrier.  Spins are bounded so a logic
// error traps instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int generation) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned int target = generation * gridDim.x;
        unsigned long long spins = 0;
        while (ld_acquire_u32(counter) < target) {
            if (++spins > (1ull << 31)) __trap();
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    return v;
}

}  // namespace stemseg
