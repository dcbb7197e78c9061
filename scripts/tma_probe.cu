// Developer probe: does a 2-D tiled TMA load accept an inner (contiguous-dimension) start coordinate that is not a
// multiple of 16 bytes / is negative, with and without 128-byte swizzle?   usage: tma_probe <swizzle 0|1> <c0>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, uint16_t* out) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32 * 128);
    const uint32_t bar_a = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
    const uint32_t dst_a = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(32 * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst_a), "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar_a), "r"(c0), "r"(0) : "memory");
    }
    uint32_t done = 0;
    for (uint32_t spins = 0; !done && spins < (1u << 22); ++spins)
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
    if (!done) { if (threadIdx.x == 0) out[0] = 0xDEAD; return; }
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main(int argc, char** argv) {
    const int swz = atoi(argv[1]), c0 = atoi(argv[2]);
    const int rows = 32, K = 512;
    std::vector<uint16_t> h(rows * K);
    for (int r = 0; r < rows; ++r) for (int k = 0; k < K; ++k) h[r * K + k] = static_cast<uint16_t>(1 + r * K + k);
    uint16_t *d, *o;
    cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 32 * 64 * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(o, 0, 32 * 64 * 2);
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    CUtensorMap map;
    cuuint64_t dims[2] = {K, rows}; cuuint64_t strides[1] = {K * 2}; cuuint32_t box[2] = {64, 32}; cuuint32_t es[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeFn>(fnp)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es,
        CU_TENSOR_MAP_INTERLEAVE_NONE, swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("swz=%d c0=%d encode failed %d\n", swz, c0, (int)r); return 1; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
    probe<<<1, 128, 8192>>>(map, c0, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("swz=%d c0=%d CUDA error: %s\n", swz, c0, cudaGetErrorString(e)); return 2; }
    std::vector<uint16_t> got(32 * 64);
    cudaMemcpy(got.data(), o, got.size() * 2, cudaMemcpyDeviceToHost);
    if (got[0] == 0xDEAD) { printf("swz=%d c0=%d TIMEOUT (bytes never arrived)\n", swz, c0); return 3; }
    int bad = 0;
    for (int rr = 0; rr < 32; ++rr) for (int k = 0; k < 64; ++k) {
        const int kk = c0 + k;
        const uint16_t want = (kk >= 0 && kk < K) ? h[rr * K + kk] : 0;
        int chunk = k / 8, within = k % 8;
        int pos = swz ? rr * 64 + ((chunk ^ (rr & 7)) * 8 + within) : rr * 64 + k;
        if (got[pos] != want) ++bad;
    }
    printf("swz=%d c0=%d ok, mismatches=%d\n", swz, c0, bad);
    return 0;
}
