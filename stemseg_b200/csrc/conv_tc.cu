// K1: 3x3x3 / 1x1x1 convolution over NDHWC bf16 planes as a tcgen05 implicit GEMM (sm_100a).
//
// Replaces nn.Conv3d(C, C', 3, padding=1) (stemseg/modeling/embedding_decoder.py:21-57 and the same lines of
// seediness_decoder.py / semseg_decoder.py) and the 1x1x1 merge convs (embedding_decoder.py:68,74,80).
//
//   D[voxel, cout] = sum_{tap, cin} A[voxel + tap, cin] * W[cout, tap, cin]          (fp32 accumulate in TMEM)
//
// * M tile = a (tt x th x tw) = 128-voxel box of one sample.  The im2col gather is done by TMA: one tiled 5-D
//   box load per (tap, 64/32-channel chunk) with the start coordinate shifted by the tap offset; out-of-bounds
//   elements (the zero padding, and ragged tile edges) are zero-filled by the TMA unit.  The box lands in shared
//   memory exactly as the canonical K-major SWIZZLE_128B/64B UMMA operand (one voxel = one 128/64-byte row).
// * B tile = [BLOCK_N couts][BLOCK_K] slab of the pre-packed weight matrix [Cout][tap][Cin] (K-major).
// * fp32 parity mode (PLANES == 2): activations and weights are split x = hi + lo into two bf16 planes and three
//   MMAs (hi*hi, hi*lo, lo*hi) accumulate into the same TMEM tile -- ~2^-17 relative operand error instead of
//   bf16's 2^-9 (measured 1.2e-5 norm-wise on the variance logits, budget 1e-4).  PLANES == 1 is the plain bf16 mode.
// * persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (single elected
//   thread) + TMEM allocator, warps 2-5 = epilogue (tcgen05.ld -> +bias -> fp32 NDHWC store).  Two TMEM accumulator
//   stages so the epilogue of tile i overlaps the main loop of tile i+1.
#include "common.cuh"
#include "trilinear.cuh"

#include <cuda.h>
#include <cstdlib>
#include <cuda_bf16.h>

namespace stemseg {
namespace {

constexpr int kBlockM = 128;
constexpr int kUmmaK = 16;                 // bf16
constexpr int kNumThreads = 192;           // 6 warps
constexpr int kNumEpilogueThreads = 128;
constexpr int kAccStages = 2;

struct ConvTcParams {
    int n, t, h, w;
    int cin, cout;
    int ntaps;                 // 27 (3x3x3, pad 1) or 1
    int tt, th, tw;            // voxel box of one M tile (tt*th*tw == 128)
    int tiles_t, tiles_h, tiles_w;
    int n_tiles_n;
    int k_slices;              // split-K over the taps: slice s handles taps [s*taps_per_slice, (s+1)*taps_per_slice)
    int taps_per_slice;
    size_t slice_stride;       // elements between the partial outputs of consecutive slices
    int num_tiles;             // n * tiles_t * tiles_h * tiles_w * n_tiles_n * k_slices
    float* out;                // [k_slices][n][t][h][w][cout] fp32 (partial sums when k_slices > 1)
    const float* bias;         // [cout] or nullptr
    // ---- wgrad mode (backward_ops.cu): plain K-major GEMM out[m][n] = sum_k A[m][k] B[n][k + b_k_offset[slice]] over the
    // zero-padded transposed planes; slice = filter tap, K range additionally split over k_splits CTAs ---------------
    int wgrad_mode;
    int b_k_offset[27];        // (dt*hp + dh)*pitch: a multiple of 8 elements (TMA needs a 16-byte aligned inner coordinate)
    int b_row_offset[27];      // (dw + 1)*cin: the dw shift selects one of three pre-shifted copies of the B rows
    int k_splits;              // >= 1
    int k_chunks_total;        // K chunks of the whole reduction (cin / BLOCK_K in normal mode)
    int tiles_per_cta;         // 0: persistent (tile = blockIdx.x + k*gridDim.x); >0: CTA b owns tiles [b*tpc, (b+1)*tpc)
                               // so that a long layer is a stream of short-lived CTAs and higher-priority kernels of
                               // other graph branches get SMs while it runs
    int num_stages;            // smem pipeline depth
    int operand_fp16;          // 1: the (single) operand planes hold fp16 instead of bf16 values
    int out_bf16;              // 1: epilogue mode 0 stores bf16 instead of fp32
    float* stat_partial;       // optional [n][cout][tiles_per_sample][2]: per-tile (sum, sum of squares) of every output
                               // channel, consumed by gn_finalize (GroupNorm statistics without re-reading the output)
    int tiles_per_sample;
    // ---- epilogue mode 1: fused output heads (the accumulator row never leaves the SM) ----------------------
    //   out[n][j][t][h][w] = act_j( W_out[j] . acc_row + up(p_low)[j] + b_j ) + coord_j
    int epi_mode;              // 0 = fp32 NDHWC store (+bias), 1 = output heads
    int head_j;                // number of outputs J
    const float* head_w;       // [J][cout]
    const float* head_b;       // [J] or nullptr
    const int* head_act;       // [J] 0 identity, 1 tanh(0.25 v), 2 sigmoid
    const int* head_coord;     // [J] 0 none, 1 t, 2 y, 3 x
    const float* p_low;        // [n][tl][hl][wl][J] = W_out . y_low at the low resolution
    int st, tl, hl, wl;        // trilinear geometry (t = tl*st, h = 2 hl, w = 2 wl)
    float x_abs, y_abs, t_abs;
    float* head_out;           // [n][J][t][h][w]
};

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && spins > (1u << 24)) __trap();
    }
}

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :
        : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
          "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// tcgen05.commit: the mbarrier is arrived on once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory, rows of ROW_BYTES (= swizzle span) bytes, 8-row groups contiguous.
// Descriptor fields (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30) (unused for a
// single swizzle atom along K), SBO>>4 [32,46) = 8 rows * ROW_BYTES, version=1 [46,48), layout type [61,64).
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    constexpr uint64_t layout = ROW_BYTES == 128 ? 2ull : (ROW_BYTES == 64 ? 4ull : 6ull);   // SW128 / SW64 / SW32
    constexpr uint64_t sbo = (8ull * ROW_BYTES) >> 4;
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// Instruction descriptor (InstrDescriptor): D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1, K-major A and B,
// N>>3 at [17,23), M>>4 at [24,29).
template <int N>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(kBlockM >> 4) << 24);
}
// same with A and B as F16 (format code 0 in both operand fields)
template <int N>
__device__ __forceinline__ constexpr uint32_t make_idesc_f16() {
    return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(kBlockM >> 4) << 24);
}

// Column sums over the 32 rows held by a warp: v[i] is this lane's (row's) value of column i.  Butterfly
// reduce-scatter (31 shuffles instead of 160): afterwards v[0] of lane l is the sum of column l over the 32 lanes.
__device__ __forceinline__ void warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
}

template <int BLOCK_N>
constexpr int tmem_columns() {
    return kAccStages * BLOCK_N <= 32 ? 32 : kAccStages * BLOCK_N <= 64 ? 64 : kAccStages * BLOCK_N <= 128 ? 128
           : kAccStages * BLOCK_N <= 256 ? 256 : 512;
}

template <int BLOCK_N, int BLOCK_K, int PLANES>
constexpr int stage_bytes() {
    return PLANES * (kBlockM * BLOCK_K * 2 + BLOCK_N * BLOCK_K * 2);
}

struct TileCoord {
    int n, t0, h0, w0, n_tile, slice, ksplit;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvTcParams& p, int tile) {
    TileCoord c;
    c.slice = tile % p.k_slices;
    tile /= p.k_slices;
    c.ksplit = tile % p.k_splits;
    tile /= p.k_splits;
    c.n_tile = tile % p.n_tiles_n;
    int m = tile / p.n_tiles_n;
    c.w0 = (m % p.tiles_w) * p.tw;
    m /= p.tiles_w;
    c.h0 = (m % p.tiles_h) * p.th;
    m /= p.tiles_h;
    c.t0 = (m % p.tiles_t) * p.tt;
    c.n = m / p.tiles_t;
    return c;
}

constexpr int kMaxStages = 8;
constexpr int kHeadJChunkTc = 8;

// MINB = 2: two CTAs per SM (register cap 168, half the shared-memory ring, 2 x 256 TMEM columns) -- used for the fused
// merge + output-heads launches, whose time is spent in the epilogue warps (J x BLOCK_N FMAs per voxel, tensor pipe
// 4-8 % active, profiles/r02_conv_ncu_summary.json): twice the epilogue warps per SM
template <int BLOCK_N, int BLOCK_K, int PLANES, int MINB = 1>
__global__ void __launch_bounds__(kNumThreads, MINB)
conv_tc_kernel(const __grid_constant__ CUtensorMap a_map0, const __grid_constant__ CUtensorMap a_map1,
               const __grid_constant__ CUtensorMap b_map0, const __grid_constant__ CUtensorMap b_map1,
               const ConvTcParams p) {
    constexpr int ROW_BYTES = BLOCK_K * 2;
    constexpr int A_BYTES = kBlockM * ROW_BYTES;
    constexpr int B_BYTES = BLOCK_N * ROW_BYTES;
    constexpr int STAGE_BYTES = stage_bytes<BLOCK_N, BLOCK_K, PLANES>();
    constexpr int TMEM_COLS = tmem_columns<BLOCK_N>();
    static_assert(BLOCK_K == 64 || BLOCK_K == 32, "BLOCK_K must equal one swizzle atom (64 or 32 bf16)");
    static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "invalid UMMA N");
    static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "operand tiles must keep 1024-byte alignment");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int STAGES = p.num_stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full_bar = empty_bar + kMaxStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + kAccStages;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + kAccStages);
    float* s_stat = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);      // [4 quarters][BLOCK_N][2]
    float* s_head_w = s_stat + 4 * BLOCK_N * 2;                                        // [J][BLOCK_N] then [J] bias

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int k_chunks = p.cin / BLOCK_K;
    // wgrad mode: K range of split s = chunks [s*k_chunks/k_splits, (s+1)*k_chunks/k_splits) (never empty: the host keeps
    // k_splits <= k_chunks)
    auto split_begin = [&](int ksplit) -> int {
        return static_cast<int>(static_cast<long long>(ksplit) * k_chunks / p.k_splits);
    };
    auto k_blocks_of = [&](const TileCoord& tc) -> int {
        if (p.k_splits == 1) return p.taps_per_slice * k_chunks;
        return split_begin(tc.ksplit + 1) - split_begin(tc.ksplit);
    };
    int tile_first = blockIdx.x, tile_last = p.num_tiles, tile_step = gridDim.x;
    if (p.tiles_per_cta > 0) {
        tile_first = blockIdx.x * p.tiles_per_cta;
        tile_last = tile_first + p.tiles_per_cta < p.num_tiles ? tile_first + p.tiles_per_cta : p.num_tiles;
        tile_step = 1;
    }

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&a_map0);
        prefetch_tmap(&b_map0);
        if (PLANES == 2) {
            prefetch_tmap(&a_map1);
            prefetch_tmap(&b_map1);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        for (int a = 0; a < kAccStages; ++a) {
            mbar_init(tmem_full_bar + a, 1);
            mbar_init(tmem_empty_bar + a, kNumEpilogueThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // whole warp: allocate the accumulator columns, publish the base address through smem
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile_first; tile < tile_last; tile += tile_step) {
                const TileCoord tc = decode_tile(p, tile);
                const int num_k_blocks = k_blocks_of(tc);
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    int tap = tc.slice * p.taps_per_slice + kb / k_chunks, c0 = (kb % k_chunks) * BLOCK_K;
                    if (p.wgrad_mode) {
                        tap = 0;
                        c0 = (split_begin(tc.ksplit) + kb) * BLOCK_K;
                    }
                    int dt = 0, dh = 0, dw = 0;
                    if (p.ntaps == 27) {
                        dt = tap / 9 - 1;
                        dh = (tap / 3) % 3 - 1;
                        dw = tap % 3 - 1;
                    }
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t* st = smem + stage * STAGE_BYTES;
                    mbar_expect_tx(full_bar + stage, STAGE_BYTES);
                    tma_load_5d(&a_map0, full_bar + stage, st, c0, tc.w0 + dw, tc.h0 + dh, tc.t0 + dt, tc.n);
                    if (PLANES == 2)
                        tma_load_5d(&a_map1, full_bar + stage, st + A_BYTES, c0, tc.w0 + dw, tc.h0 + dh, tc.t0 + dt, tc.n);
                    uint8_t* sb = st + PLANES * A_BYTES;
                    const int kcoord = p.wgrad_mode ? c0 + p.b_k_offset[tc.slice] : tap * p.cin + c0;
                    const int brow = tc.n_tile * BLOCK_N + (p.wgrad_mode ? p.b_row_offset[tc.slice] : 0);
                    tma_load_2d(&b_map0, full_bar + stage, sb, kcoord, brow);
                    if (PLANES == 2) tma_load_2d(&b_map1, full_bar + stage, sb + B_BYTES, kcoord, brow);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = p.operand_fp16 ? make_idesc_f16<BLOCK_N>() : make_idesc<BLOCK_N>();
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = tile_first; tile < tile_last; tile += tile_step) {
            mbar_wait(tmem_empty_bar + acc, acc_phase ^ 1);       // epilogue has drained this accumulator
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
            const int num_k_blocks = k_blocks_of(decode_tile(p, tile));
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(full_bar + stage, phase);                // TMA bytes have landed
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t b0 = a0 + PLANES * A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / kUmmaK; ++k) {
                        const uint32_t koff = k * kUmmaK * 2;      // bytes inside the swizzle atom
                        const uint64_t a_hi = make_kmajor_desc<ROW_BYTES>(a0 + koff);
                        const uint64_t b_hi = make_kmajor_desc<ROW_BYTES>(b0 + koff);
                        umma_bf16(tmem_d, a_hi, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                        if (PLANES == 2) {
                            const uint64_t a_lo = make_kmajor_desc<ROW_BYTES>(a0 + A_BYTES + koff);
                            const uint64_t b_lo = make_kmajor_desc<ROW_BYTES>(b0 + B_BYTES + koff);
                            umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
                            umma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
                        }
                    }
                    umma_commit(empty_bar + stage);                               // frees the smem stage
                    if (kb == num_k_blocks - 1) umma_commit(tmem_full_bar + acc); // accumulator complete
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================== epilogue warps (2..5) =====================
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;          // accumulator row == voxel index inside the tile box
        const int dw = row % p.tw, dh = (row / p.tw) % p.th, dt = row / (p.tw * p.th);
        if (p.epi_mode == 1) {     // stage the output-conv weights in shared memory (epilogue warps only)
            const int et = threadIdx.x - (kNumThreads - kNumEpilogueThreads);
            for (int i = et; i < p.head_j * BLOCK_N; i += kNumEpilogueThreads) s_head_w[i] = p.head_w[i];
            for (int i = et; i < p.head_j; i += kNumEpilogueThreads)
                s_head_w[p.head_j * BLOCK_N + i] = p.head_b ? p.head_b[i] : 0.f;
            asm volatile("bar.sync 1, %0;" ::"n"(kNumEpilogueThreads) : "memory");
        }
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = tile_first; tile < tile_last; tile += tile_step) {
            const TileCoord tc = decode_tile(p, tile);
            const int t = tc.t0 + dt, h = tc.h0 + dh, w = tc.w0 + dw;
            const bool valid = t < p.t && h < p.h && w < p.w;
            mbar_wait(tmem_full_bar + acc, acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(acc * BLOCK_N);
            if (p.epi_mode == 0) {
                float* out_row = p.out + (static_cast<size_t>(tc.ksplit) * p.k_slices + tc.slice) * p.slice_stride +
                                 ((((static_cast<size_t>(tc.n) * p.t + t) * p.h + h) * p.w + w) * p.cout +
                                  static_cast<size_t>(tc.n_tile) * BLOCK_N);
                const float* bias = (p.bias && tc.slice == 0) ? p.bias + tc.n_tile * BLOCK_N : nullptr;
                const bool stats = p.stat_partial != nullptr;
                if (stats) asm volatile("bar.sync 1, %0;" ::"n"(kNumEpilogueThreads) : "memory");  // s_stat free again
#pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + c, v);
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int q = 0; q < 32; q += 4) {
                        float4 o;
                        o.x = __uint_as_float(v[q + 0]);
                        o.y = __uint_as_float(v[q + 1]);
                        o.z = __uint_as_float(v[q + 2]);
                        o.w = __uint_as_float(v[q + 3]);
                        if (bias) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c + q));
                            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                        }
                        if (valid && !p.out_bf16) *reinterpret_cast<float4*>(out_row + c + q) = o;
                        f[q + 0] = valid ? o.x : 0.f;
                        f[q + 1] = valid ? o.y : 0.f;
                        f[q + 2] = valid ? o.z : 0.f;
                        f[q + 3] = valid ? o.w : 0.f;
                    }
                    if (p.out_bf16 && valid) {       // same element index, 2-byte elements: 32 channels = 4 x 16 bytes
                        __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(p.out) + (out_row - p.out) + c;
#pragma unroll
                        for (int q = 0; q < 32; q += 8) {
                            __nv_bfloat162 h[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(f[q + 2 * k], f[q + 2 * k + 1]);
                            *reinterpret_cast<uint4*>(o16 + q) = *reinterpret_cast<const uint4*>(h);
                        }
                    }
                    if (stats) {
                        float f2[32];
#pragma unroll
                        for (int q = 0; q < 32; ++q) f2[q] = f[q] * f[q];
                        warp_column_sums(f, lane);
                        warp_column_sums(f2, lane);
                        s_stat[(quarter * BLOCK_N + c + lane) * 2 + 0] = f[0];
                        s_stat[(quarter * BLOCK_N + c + lane) * 2 + 1] = f2[0];
                    }
                }
                if (stats) {
                    asm volatile("bar.sync 1, %0;" ::"n"(kNumEpilogueThreads) : "memory");
                    const int et = threadIdx.x - (kNumThreads - kNumEpilogueThreads);
                    const int m_in_sample = (tile / (p.k_slices * p.k_splits * p.n_tiles_n)) % p.tiles_per_sample;
                    for (int col = et; col < BLOCK_N; col += kNumEpilogueThreads) {
                        float a = 0.f, b = 0.f;
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {        // fixed order: deterministic
                            a += s_stat[(qq * BLOCK_N + col) * 2 + 0];
                            b += s_stat[(qq * BLOCK_N + col) * 2 + 1];
                        }
                        float* dstp = p.stat_partial +
                                      ((static_cast<size_t>(tc.n) * p.cout + tc.n_tile * BLOCK_N + col) * p.tiles_per_sample +
                                       m_in_sample) * 2;
                        dstp[0] = a;
                        dstp[1] = b;
                    }
                }
            } else {
                // fused output heads: every thread owns one voxel's accumulator row
                const int J = p.head_j;
                const float* s_head_b = s_head_w + J * BLOCK_N;
                Tri tr;
                if (valid) tr = make_tri(tc.n, t, h, w, p.st, p.tl, p.hl, p.wl, J);
                const size_t spatial = static_cast<size_t>(p.t) * p.h * p.w;
                const size_t vox = (static_cast<size_t>(t) * p.h + h) * p.w + w;
#pragma unroll 1
                for (int j0 = 0; j0 < J; j0 += kHeadJChunkTc) {
                    float a[kHeadJChunkTc];
#pragma unroll
                    for (int j = 0; j < kHeadJChunkTc; ++j) a[j] = 0.f;
#pragma unroll 1
                    for (int c = 0; c < BLOCK_N; c += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr + c, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < kHeadJChunkTc; ++j) {
                            if (j0 + j < J) {
                                const float4* wr = reinterpret_cast<const float4*>(s_head_w + (j0 + j) * BLOCK_N + c);
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    const float4 w4 = wr[q];
                                    a[j] = fmaf(__uint_as_float(v[4 * q + 0]), w4.x, a[j]);
                                    a[j] = fmaf(__uint_as_float(v[4 * q + 1]), w4.y, a[j]);
                                    a[j] = fmaf(__uint_as_float(v[4 * q + 2]), w4.z, a[j]);
                                    a[j] = fmaf(__uint_as_float(v[4 * q + 3]), w4.w, a[j]);
                                }
                            }
                        }
                    }
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < kHeadJChunkTc; ++j) {
                            const int jj = j0 + j;
                            if (jj < J) {
                                float val = a[j] + s_head_b[jj];
#pragma unroll
                                for (int k = 0; k < 8; ++k)
                                    if (tr.wgt[k] != 0.f) val = fmaf(tr.wgt[k], __ldg(p.p_low + tr.off[k] + jj), val);
                                const int ac = p.head_act[jj];
                                if (ac == 1) val = tanhf(0.25f * val);
                                else if (ac == 2) val = 1.0f / (1.0f + expf(-val));
                                const int cd = p.head_coord[jj];
                                if (cd == 1) val += linspace_value(p.t_abs, p.t, t);
                                else if (cd == 2) val += linspace_value(p.y_abs, p.h, h);
                                else if (cd == 3) val += linspace_value(p.x_abs, p.w, w);
                                p.head_out[(static_cast<size_t>(tc.n) * J + jj) * spatial + vox] = val;
                            }
                        }
                    }
                }
            }
            tcgen05_fence_before();
            mbar_arrive(tmem_empty_bar + acc);
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Direct weight gradient: dW[tap][co][ci] = sum_vox dy[vox][co] * x[vox + tap][ci] as a tcgen05 GEMM whose REDUCTION
// dimension is the voxel index.  Both operands stay in the NDHWC bf16 plane layout of the forward pass and are fed to
// the tensor core as MN-major tiles: one TMA box = (64 channels) x (32 voxels) lands as 32 rows of 128 bytes, which
// is exactly the canonical SWIZZLE_128B MN-major atom sequence (8 K-rows x 64 MN-elements per 1024-byte atom; SBO =
// 1024 bytes between 8-row groups, LBO = one box = 4096 bytes between 64-channel blocks).  The tap shift and the
// zero padding are, as in the forward pass, TMA start coordinates + out-of-bounds zero fill on the x operand -- no
// transposed copies, no padded volume, no im2col buffer.
//   M = 128 output channels (two 64-channel boxes of dy), N = BLOCK_N input channels (BLOCK_N/64 boxes of x),
//   K = 32 voxels per pipeline stage (two UMMA K=16 steps).
// tile = (tap, m_tile, n_tile, k_split); taps vary fastest so that CTAs running at the same time read the same voxel
// range (L2 reuse across the 27 taps).  Partial sums per k_split are reduced by stemseg_wgrad_reduce.
// ---------------------------------------------------------------------------------------------------------------
struct WgradParams {
    int t, h, w;
    int cin, cout;
    int ntaps;
    int tt, th, tw;            // 32-voxel box
    int tiles_t, tiles_h, tiles_w, k_tiles;
    int m_tiles, n_tiles, k_splits;
    int num_tiles;
    int num_stages;
    float* out;                // [k_splits][ntaps][cout][cin]
};

constexpr int kWgBoxVoxels = 32;
constexpr int kWgBoxBytes = kWgBoxVoxels * 128;     // 64 channels x 32 voxels, bf16

// MN-major SWIZZLE_128B operand: LBO = stride between 64-element blocks along M/N, SBO = stride between 8-row K groups
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr) {
    constexpr uint64_t lbo = static_cast<uint64_t>(kWgBoxBytes) >> 4;
    constexpr uint64_t sbo = 1024ull >> 4;
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}

template <int BLOCK_N, int PLANES>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap dy_map0, const __grid_constant__ CUtensorMap dy_map1,
                  const __grid_constant__ CUtensorMap x_map0, const __grid_constant__ CUtensorMap x_map1,
                  const WgradParams p) {
    constexpr int NB = BLOCK_N / 64;
    constexpr int A_BYTES = 2 * kWgBoxBytes;                // per plane: 128 output channels
    constexpr int B_BYTES = NB * kWgBoxBytes;               // per plane: BLOCK_N input channels
    constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
    constexpr int TMEM_COLS = tmem_columns<BLOCK_N>();
    static_assert(BLOCK_N % 64 == 0 && BLOCK_N >= 64 && BLOCK_N <= 256, "wgrad N tile is a multiple of 64 channels");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int STAGES = p.num_stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full_bar = empty_bar + kMaxStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + kAccStages;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + kAccStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    struct WTile { int tap, m_tile, n_tile, k_begin, k_end; };
    auto decode = [&](int tile) -> WTile {
        WTile c;
        c.tap = tile % p.ntaps;
        tile /= p.ntaps;
        c.m_tile = tile % p.m_tiles;
        tile /= p.m_tiles;
        c.n_tile = tile % p.n_tiles;
        const int ks = tile / p.n_tiles;
        c.k_begin = static_cast<int>(static_cast<long long>(ks) * p.k_tiles / p.k_splits);
        c.k_end = static_cast<int>(static_cast<long long>(ks + 1) * p.k_tiles / p.k_splits);
        return c;
    };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&dy_map0);
        prefetch_tmap(&x_map0);
        if (PLANES == 2) {
            prefetch_tmap(&dy_map1);
            prefetch_tmap(&x_map1);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        for (int a = 0; a < kAccStages; ++a) {
            mbar_init(tmem_full_bar + a, 1);
            mbar_init(tmem_empty_bar + a, kNumEpilogueThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const WTile tc = decode(tile);
                int dt = 0, dh = 0, dw = 0;
                if (p.ntaps == 27) {
                    dt = tc.tap / 9 - 1;
                    dh = (tc.tap / 3) % 3 - 1;
                    dw = tc.tap % 3 - 1;
                }
                for (int kt = tc.k_begin; kt < tc.k_end; ++kt) {
                    const int w0 = (kt % p.tiles_w) * p.tw;
                    const int h0 = ((kt / p.tiles_w) % p.tiles_h) * p.th;
                    const int t0 = (kt / (p.tiles_w * p.tiles_h)) * p.tt;
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t* st = smem + stage * STAGE_BYTES;
                    mbar_expect_tx(full_bar + stage, STAGE_BYTES);
#pragma unroll
                    for (int pl = 0; pl < PLANES; ++pl) {
                        const CUtensorMap* dm = pl == 0 ? &dy_map0 : &dy_map1;
                        const CUtensorMap* xm = pl == 0 ? &x_map0 : &x_map1;
#pragma unroll
                        for (int mb = 0; mb < 2; ++mb)
                            tma_load_5d(dm, full_bar + stage, st + (pl * 2 + mb) * kWgBoxBytes,
                                        tc.m_tile * kBlockM + mb * 64, w0, h0, t0, 0);
                        uint8_t* sb = st + PLANES * A_BYTES + pl * B_BYTES;
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb)
                            tma_load_5d(xm, full_bar + stage, sb + nb * kWgBoxBytes, tc.n_tile * BLOCK_N + nb * 64,
                                        w0 + dw, h0 + dh, t0 + dt, 0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc<BLOCK_N>() | (1u << 15) | (1u << 16);     // A and B are MN-major
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const WTile tc = decode(tile);
            mbar_wait(tmem_empty_bar + acc, acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
            const int num_k = tc.k_end - tc.k_begin;
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(full_bar + stage, phase);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t b0 = a0 + PLANES * A_BYTES;
#pragma unroll
                    for (int k = 0; k < kWgBoxVoxels / kUmmaK; ++k) {
                        const uint32_t koff = k * kUmmaK * 128;        // 16 voxel rows of 128 bytes
                        const uint64_t a_hi = make_mnmajor_desc(a0 + koff);
                        const uint64_t b_hi = make_mnmajor_desc(b0 + koff);
                        umma_bf16(tmem_d, a_hi, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                        if (PLANES == 2) {
                            const uint64_t a_lo = make_mnmajor_desc(a0 + A_BYTES + koff);
                            const uint64_t b_lo = make_mnmajor_desc(b0 + B_BYTES + koff);
                            umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
                            umma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
                        }
                    }
                    umma_commit(empty_bar + stage);
                    if (kb == num_k - 1) umma_commit(tmem_full_bar + acc);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================== epilogue warps (2..5): fp32 partial sums -> [k_split][tap][co][ci] =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const WTile tc = decode(tile);
            const int ks = tile / (p.ntaps * p.m_tiles * p.n_tiles);
            const int co = tc.m_tile * kBlockM + row;
            mbar_wait(tmem_full_bar + acc, acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(acc * BLOCK_N);
            float* out_row = p.out + ((static_cast<size_t>(ks) * p.ntaps + tc.tap) * p.cout + co) * p.cin +
                             static_cast<size_t>(tc.n_tile) * BLOCK_N;
            const int cols_left = p.cin - tc.n_tile * BLOCK_N;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + c, v);
                tmem_ld_wait();
                if (co < p.cout && c < cols_left) {          // cin is a multiple of 32: whole 32-column groups
#pragma unroll
                    for (int q = 0; q < 32; q += 4) {
                        float4 o;
                        o.x = __uint_as_float(v[q + 0]);
                        o.y = __uint_as_float(v[q + 1]);
                        o.z = __uint_as_float(v[q + 2]);
                        o.w = __uint_as_float(v[q + 3]);
                        *reinterpret_cast<float4*>(out_row + c + q) = o;
                    }
                }
            }
            tcgen05_fence_before();
            mbar_arrive(tmem_empty_bar + acc);
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || ptr == nullptr)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}

int encode_act_map(CUtensorMap* map, const void* base, const ConvTcParams& p, int block_k) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return STEMSEG_ERR_CUDA;
    }
    const cuuint64_t dims[5] = {static_cast<cuuint64_t>(p.cin), static_cast<cuuint64_t>(p.w),
                                static_cast<cuuint64_t>(p.h), static_cast<cuuint64_t>(p.t),
                                static_cast<cuuint64_t>(p.n)};
    const cuuint64_t c2 = static_cast<cuuint64_t>(p.cin) * 2;
    const cuuint64_t strides[4] = {c2, c2 * p.w, c2 * p.w * p.h, c2 * p.w * p.h * p.t};
    const cuuint32_t box[5] = {static_cast<cuuint32_t>(block_k), static_cast<cuuint32_t>(p.tw),
                               static_cast<cuuint32_t>(p.th), static_cast<cuuint32_t>(p.tt), 1u};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapSwizzle sw = block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(activation) failed with CUresult %d", static_cast<int>(r));
        return STEMSEG_ERR_CUDA;
    }
    return STEMSEG_OK;
}

int encode_weight_map(CUtensorMap* map, const void* base, int64_t k_total, int64_t cout_rows, int block_k,
                      int block_n) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return STEMSEG_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_total), static_cast<cuuint64_t>(cout_rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(k_total) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(block_k), static_cast<cuuint32_t>(block_n)};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weights) failed with CUresult %d", static_cast<int>(r));
        return STEMSEG_ERR_CUDA;
    }
    return STEMSEG_OK;
}

constexpr int kSmemBudgetDefault = 200 * 1024;
constexpr int kSmemLimit = 227 * 1024;

// Shared memory the operand ring may take per CTA.  STEMSEG_CONV_SMEM_KB lowers it (e.g. 150) so that another resident
// kernel's CTAs -- NCCL's all-reduce during the data-parallel backward pass -- fit next to a persistent conv CTA.
int smem_budget() {
    static int budget = -1;
    if (budget < 0) {
        budget = kSmemBudgetDefault;
        const char* e = getenv("STEMSEG_CONV_SMEM_KB");
        if (e != nullptr) {
            const int kb = atoi(e);
            if (kb >= 64 && kb <= 200) budget = kb * 1024;
        }
    }
    return budget;
}

template <int BLOCK_N, int BLOCK_K, int PLANES>
int launch_variant(const CUtensorMap* maps, ConvTcParams& p, int max_ctas, cudaStream_t stream) {
    constexpr int stage = stage_bytes<BLOCK_N, BLOCK_K, PLANES>();
    const int head_bytes = p.epi_mode == 1 ? static_cast<int>(align_up((p.head_j * BLOCK_N + p.head_j) * sizeof(float), 16)) : 0;
    const int fixed = 1024 /*align*/ + 256 /*barriers*/ + 4 * BLOCK_N * 2 * 4 /*stat staging*/ + head_bytes;
    int stages = (smem_budget() - head_bytes - 4 * BLOCK_N * 2 * 4) / stage;
    if (stages * stage + fixed > kSmemLimit) stages = (kSmemLimit - fixed) / stage;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2 || stages * stage + fixed > kSmemLimit) {
        set_error("conv_tc: tile %dx%dx%d does not fit shared memory with %d head outputs", BLOCK_N, BLOCK_K, PLANES,
                  p.head_j);
        return STEMSEG_ERR_UNSUPPORTED;
    }
    if constexpr (BLOCK_N <= 128) {
        // fused output heads: epilogue-bound -> two CTAs per SM when two pipeline stages fit in half the shared memory
        constexpr int kHalfSmem = 110 * 1024;
        const int stages2 = (kHalfSmem - fixed) / stage;
        if (p.epi_mode == 1 && stages2 >= 2 && p.tiles_per_cta == 0) {
            p.num_stages = stages2 > kMaxStages ? kMaxStages : stages2;
            auto kernel2 = conv_tc_kernel<BLOCK_N, BLOCK_K, PLANES, 2>;
            SS_CUDA_OK(cudaFuncSetAttribute(kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, kHalfSmem));
            int grid2 = 2 * device_sm_count();
            if (max_ctas > 0 && grid2 > max_ctas) grid2 = max_ctas;
            if (grid2 > p.num_tiles) grid2 = p.num_tiles;
            kernel2<<<grid2, kNumThreads, p.num_stages * stage + fixed, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
            SS_CUDA_OK(cudaGetLastError());
            return STEMSEG_OK;
        }
    }
    p.num_stages = stages;
    const int smem_bytes = stages * stage + fixed;
    auto kernel = conv_tc_kernel<BLOCK_N, BLOCK_K, PLANES>;
    SS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    int grid = device_sm_count();
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    if (grid > p.num_tiles) grid = p.num_tiles;
    if (p.tiles_per_cta > 0) grid = (p.num_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    kernel<<<grid, kNumThreads, smem_bytes, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

template <int BLOCK_K, int PLANES>
int launch_by_n(int block_n, const CUtensorMap* maps, ConvTcParams& p, int max_ctas, cudaStream_t stream) {
    switch (block_n) {
        case 256: return launch_variant<256, BLOCK_K, PLANES>(maps, p, max_ctas, stream);
        case 128: return launch_variant<128, BLOCK_K, PLANES>(maps, p, max_ctas, stream);
        case 64: return launch_variant<64, BLOCK_K, PLANES>(maps, p, max_ctas, stream);
        case 32: return launch_variant<32, BLOCK_K, PLANES>(maps, p, max_ctas, stream);
    }
    set_error("conv_tc: unsupported BLOCK_N %d", block_n);
    return STEMSEG_ERR_UNSUPPORTED;
}

// choose the 128-voxel box (tt, th, tw) with the least padded volume (ties: prefer wide w, then h)
void choose_box(int t, int h, int w, int* tt, int* th, int* tw) {
    long long best = -1;
    for (int a = 1; a <= 128; a *= 2)
        for (int b = 1; a * b <= 128; b *= 2) {
            const int c = 128 / (a * b);
            // a = tt, b = th, c = tw
            const long long vol = 1ll * ((t + a - 1) / a) * ((h + b - 1) / b) * ((w + c - 1) / c);
            const long long score = vol * 1024 - c * 8 - b;     // fewer tiles first, then wider rows
            if (best < 0 || score < best) {
                best = score;
                *tt = a; *th = b; *tw = c;
            }
        }
}

int block_n_for(int cout) { return cout % 256 == 0 ? 256 : cout % 128 == 0 ? 128 : cout % 64 == 0 ? 64 : 32; }

// split-K heuristic: layers with fewer tiles than SMs are latency-bound on one tile's serial K loop; split the
// taps over k_slices CTAs (<= 2 waves in total) -- the consumers add the partial outputs in a fixed order.
int auto_split(const StemsegConvShape* s) {
    if (s->kernel_size != 3) return 1;
    int tt, th, tw;
    choose_box(s->t, s->h, s->w, &tt, &th, &tw);
    const long long tiles = 1ll * s->n * ((s->t + tt - 1) / tt) * ((s->h + th - 1) / th) * ((s->w + tw - 1) / tw) *
                            (s->cout / block_n_for(s->cout));
    const int sms = device_sm_count();
    // cost model: rounds of CTAs x (fraction of the K loop per CTA + fixed per-tile overhead ~4 % of a full tile)
    int best = 1;
    double best_cost = 1e30;
    const int options[4] = {1, 3, 9, 27};
    for (int o : options) {
        const long long rounds = (tiles * o + sms - 1) / sms;
        const double cost = static_cast<double>(rounds) * (1.0 / o + 0.04) + (o > 1 ? 0.05 : 0.0);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = o;
        }
    }
    return best;
}

}  // namespace
}  // namespace stemseg

using namespace stemseg;

extern "C" int32_t stemseg_conv3d_auto_split(const StemsegConvShape* s) {
    if (s == nullptr || s->n < 1 || s->t < 1 || s->h < 1 || s->w < 1 || s->cout < 32) return 1;
    return auto_split(s);
}

static int32_t conv3d_impl(const void* act_planes, const void* weight_planes, const float* bias, float* out,
                           float* stat_partial, const StemsegConvShape* s, int32_t max_ctas, void* stream_,
                           const ConvTcParams* head) {
    SS_REQUIRE(s != nullptr && act_planes && weight_planes && (out || head), "conv3d: null pointer");
    SS_REQUIRE(s->planes == 1 || s->planes == 2 || s->planes == STEMSEG_PLANES_FP16,
               "conv3d: planes must be 1, 2 or STEMSEG_PLANES_FP16");
    const int plane_count = s->planes == 2 ? 2 : 1;
    SS_REQUIRE(s->kernel_size == 3 || s->kernel_size == 1, "conv3d: kernel_size must be 1 or 3");
    SS_REQUIRE(s->n >= 1 && s->t >= 1 && s->h >= 1 && s->w >= 1, "conv3d: empty volume");
    SS_REQUIRE(s->cin >= 32 && s->cin % 32 == 0, "conv3d: cin must be a positive multiple of 32 (got %d)", s->cin);
    SS_REQUIRE(s->cout >= 32 && s->cout % 32 == 0, "conv3d: cout must be a positive multiple of 32 (got %d)", s->cout);
    SS_REQUIRE((reinterpret_cast<uintptr_t>(act_planes) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight_planes) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
               "conv3d: pointers must be 16-byte aligned");
    int rc = require_sm100();
    if (rc != STEMSEG_OK) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);

    ConvTcParams p;
    p.n = s->n; p.t = s->t; p.h = s->h; p.w = s->w;
    p.cin = s->cin; p.cout = s->cout;
    p.ntaps = s->kernel_size == 3 ? 27 : 1;
    choose_box(p.t, p.h, p.w, &p.tt, &p.th, &p.tw);
    p.tiles_t = (p.t + p.tt - 1) / p.tt;
    p.tiles_h = (p.h + p.th - 1) / p.th;
    p.tiles_w = (p.w + p.tw - 1) / p.tw;
    const int block_n = block_n_for(s->cout);
    p.n_tiles_n = s->cout / block_n;
    p.k_slices = s->split_k >= 1 ? s->split_k : 1;
    SS_REQUIRE(p.ntaps % p.k_slices == 0, "conv3d: split_k %d does not divide the %d taps", p.k_slices, p.ntaps);
    p.taps_per_slice = p.ntaps / p.k_slices;
    p.slice_stride = static_cast<size_t>(p.n) * p.t * p.h * p.w * p.cout;
    const long long tiles = 1ll * p.n * p.tiles_t * p.tiles_h * p.tiles_w * p.n_tiles_n * p.k_slices;
    SS_REQUIRE(tiles < 0x7FFFFFFFll, "conv3d: too many tiles");
    p.num_tiles = static_cast<int>(tiles);
    p.out = out;
    p.bias = bias;
    p.num_stages = 0;
    p.operand_fp16 = s->planes == STEMSEG_PLANES_FP16 ? 1 : 0;
    p.out_bf16 = s->out_bf16 ? 1 : 0;
    SS_REQUIRE(!p.out_bf16 || (p.k_slices == 1 && head == nullptr), "conv3d: out_bf16 needs an unsplit plain-store launch");
    p.wgrad_mode = 0;
    p.k_splits = 1;
    p.k_chunks_total = 0;
    for (int i = 0; i < 27; ++i) p.b_k_offset[i] = 0;
    p.tiles_per_cta = s->tiles_per_cta > 0 ? s->tiles_per_cta : 0;
    p.tiles_per_sample = p.tiles_t * p.tiles_h * p.tiles_w;
    p.stat_partial = stat_partial;
    SS_REQUIRE(stat_partial == nullptr || (p.k_slices == 1 && head == nullptr),
               "conv3d: fused statistics need an unsplit plain-store launch");
    p.epi_mode = 0;
    p.head_j = 0;
    if (head != nullptr) {
        SS_REQUIRE(p.n_tiles_n == 1 && p.k_slices == 1, "conv3d+heads: cout must fit one N tile (<= 256, got %d)", s->cout);
        p.epi_mode = 1;
        p.head_j = head->head_j; p.head_w = head->head_w; p.head_b = head->head_b;
        p.head_act = head->head_act; p.head_coord = head->head_coord; p.p_low = head->p_low;
        p.st = head->st; p.tl = head->tl; p.hl = head->hl; p.wl = head->wl;
        p.x_abs = head->x_abs; p.y_abs = head->y_abs; p.t_abs = head->t_abs;
        p.head_out = head->head_out;
    }

    // BLOCK_K: 64 channels (SWIZZLE_128B) when the stage still leaves >= 3 pipeline stages, else 32 (SWIZZLE_64B)
    int block_k = (s->cin % 64 == 0) ? 64 : 32;
    if (block_k == 64 && plane_count == 2 && block_n == 256) block_k = 32;
    // fused heads (two CTAs per SM, launch_variant): two-plane stages of 64 channels would not fit twice
    if (block_k == 64 && plane_count == 2 && head != nullptr && block_n <= 128) block_k = 32;

    const size_t act_plane_bytes = static_cast<size_t>(p.n) * p.t * p.h * p.w * p.cin * 2;
    const int64_t k_total = static_cast<int64_t>(p.ntaps) * p.cin;
    const size_t w_plane_bytes = static_cast<size_t>(k_total) * p.cout * 2;
    CUtensorMap maps[4];
    const uint8_t* a = static_cast<const uint8_t*>(act_planes);
    const uint8_t* b = static_cast<const uint8_t*>(weight_planes);
    for (int pl = 0; pl < 2; ++pl) {
        const int src = pl < plane_count ? pl : 0;
        rc = encode_act_map(&maps[pl], a + src * act_plane_bytes, p, block_k);
        if (rc != STEMSEG_OK) return rc;
        rc = encode_weight_map(&maps[2 + pl], b + src * w_plane_bytes, k_total, p.cout, block_k, block_n);
        if (rc != STEMSEG_OK) return rc;
    }
    if (plane_count == 2)
        return block_k == 64 ? launch_by_n<64, 2>(block_n, maps, p, max_ctas, stream)
                             : launch_by_n<32, 2>(block_n, maps, p, max_ctas, stream);
    return block_k == 64 ? launch_by_n<64, 1>(block_n, maps, p, max_ctas, stream)
                         : launch_by_n<32, 1>(block_n, maps, p, max_ctas, stream);
}

extern "C" int32_t stemseg_conv3d_bf16_planes(const void* act_planes, const void* weight_planes, const float* bias,
                                              float* out, float* stat_partial, const StemsegConvShape* s,
                                              int32_t max_ctas, void* stream_) {
    return conv3d_impl(act_planes, weight_planes, bias, out, stat_partial, s, max_ctas, stream_, nullptr);
}

extern "C" int32_t stemseg_conv3d_tiles_per_sample(const StemsegConvShape* s) {
    if (s == nullptr || s->t < 1 || s->h < 1 || s->w < 1) return 0;
    int tt, th, tw;
    choose_box(s->t, s->h, s->w, &tt, &th, &tw);
    return ((s->t + tt - 1) / tt) * ((s->h + th - 1) / th) * ((s->w + tw - 1) / tw);
}

extern "C" int32_t stemseg_conv1x1_head_output(const void* act_planes, const void* weight_planes,
                                               const StemsegConvShape* s, const float* p_low, int32_t t_scale,
                                               const float* out_weight, const float* out_bias,
                                               const int32_t* activation, const int32_t* coordinate, int32_t n_out,
                                               float time_scale, float* out, int32_t max_ctas, void* stream_) {
    SS_REQUIRE(s != nullptr && p_low && out_weight && activation && coordinate && out, "conv1x1_head_output: null pointer");
    SS_REQUIRE(s->kernel_size == 1 && s->split_k <= 1, "conv1x1_head_output: the merge conv must be 1x1x1, unsplit");
    SS_REQUIRE(t_scale == 1 || t_scale == 2, "conv1x1_head_output: temporal scale must be 1 or 2");
    SS_REQUIRE(s->h % 2 == 0 && s->w % 2 == 0 && s->t % t_scale == 0, "conv1x1_head_output: bad upsampling geometry");
    SS_REQUIRE(n_out >= 1 && n_out <= 64, "conv1x1_head_output: n_out %d out of range [1,64]", n_out);
    SS_REQUIRE((reinterpret_cast<uintptr_t>(out_weight) & 15) == 0, "conv1x1_head_output: out_weight must be 16-byte aligned");
    ConvTcParams head;
    head.head_j = n_out; head.head_w = out_weight; head.head_b = out_bias;
    head.head_act = activation; head.head_coord = coordinate; head.p_low = p_low;
    head.st = t_scale; head.tl = s->t / t_scale; head.hl = s->h / 2; head.wl = s->w / 2;
    head.x_abs = fmaxf(1.0f, static_cast<float>(static_cast<double>(s->w) / static_cast<double>(s->h)));
    head.y_abs = fmaxf(1.0f, static_cast<float>(static_cast<double>(s->h) / static_cast<double>(s->w)));
    head.t_abs = time_scale;
    head.head_out = out;
    return conv3d_impl(act_planes, weight_planes, nullptr, nullptr, nullptr, s, max_ctas, stream_, &head);
}

namespace {
// 32-voxel box (tt, th, tw) with the fewest tiles (ties: wider rows)
void choose_box32(int t, int h, int w, int* tt, int* th, int* tw) {
    long long best = -1;
    for (int a = 1; a <= kWgBoxVoxels; a *= 2)
        for (int b = 1; a * b <= kWgBoxVoxels; b *= 2) {
            const int c = kWgBoxVoxels / (a * b);
            const long long vol = 1ll * ((t + a - 1) / a) * ((h + b - 1) / b) * ((w + c - 1) / c);
            const long long score = vol * 1024 - c * 8 - b;
            if (best < 0 || score < best) {
                best = score;
                *tt = a; *th = b; *tw = c;
            }
        }
}
int wgrad_direct_block_n(int cin) { return cin % 256 == 0 ? 256 : cin % 128 == 0 ? 128 : 64; }

int encode_box_map(CUtensorMap* map, const void* base, int channels, int t, int h, int w, int tt, int th, int tw) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return STEMSEG_ERR_CUDA;
    }
    const cuuint64_t dims[5] = {static_cast<cuuint64_t>(channels), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h),
                                static_cast<cuuint64_t>(t), 1};
    const cuuint64_t c2 = static_cast<cuuint64_t>(channels) * 2;
    const cuuint64_t strides[4] = {c2, c2 * w, c2 * w * h, c2 * w * h * t};
    const cuuint32_t box[5] = {64u, static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(th), static_cast<cuuint32_t>(tt), 1u};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(wgrad operand, %d channels) failed with CUresult %d", channels, static_cast<int>(r));
        return STEMSEG_ERR_CUDA;
    }
    return STEMSEG_OK;
}

template <int BLOCK_N, int PLANES>
int launch_wgrad(const CUtensorMap* maps, WgradParams& p, cudaStream_t stream) {
    constexpr int stage = PLANES * (2 + BLOCK_N / 64) * kWgBoxBytes;
    const int fixed = 1024 + 256;
    int stages = smem_budget() / stage;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages * stage + fixed > kSmemLimit) stages = (kSmemLimit - fixed) / stage;
    if (stages < 2) {
        set_error("conv_wgrad: stage of %d bytes does not fit shared memory", stage);
        return STEMSEG_ERR_UNSUPPORTED;
    }
    p.num_stages = stages;
    auto kernel = conv_wgrad_kernel<BLOCK_N, PLANES>;
    SS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    int grid = device_sm_count();
    if (grid > p.num_tiles) grid = p.num_tiles;
    kernel<<<grid, kNumThreads, stages * stage + fixed, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}
}  // namespace

extern "C" int32_t stemseg_wgrad_direct_k_splits(int32_t cout, int32_t cin, int32_t t, int32_t h, int32_t w,
                                                 int32_t kernel_size) {
    if (cout < 1 || cin < 32 || t < 1 || h < 1 || w < 1) return 1;
    int tt, th, tw;
    choose_box32(t, h, w, &tt, &th, &tw);
    const long long k_tiles = 1ll * ((t + tt - 1) / tt) * ((h + th - 1) / th) * ((w + tw - 1) / tw);
    const int bn = wgrad_direct_block_n(cin);
    const long long base = 1ll * (kernel_size == 3 ? 27 : 1) * ((cout + kBlockM - 1) / kBlockM) * ((cin + bn - 1) / bn);
    long long k = (2ll * device_sm_count() + base - 1) / base;            // ~2 tiles per SM
    if (k > k_tiles / 4) k = k_tiles / 4;                                 // at least 4 K blocks per tile
    if (k > 64) k = 64;
    if (k < 1) k = 1;
    return static_cast<int32_t>(k);
}

// slices[ks][tap][co][ci] = sum over the voxel range of split ks of dy[vox][co] * x[vox + tap][ci]
// dy_planes / x_planes: NDHWC bf16 planes [P][1][t][h][w][C] (stemseg_to_planes / the forward pass's activations)
extern "C" int32_t stemseg_conv3d_wgrad_direct(const void* dy_planes, const void* x_planes, int32_t cout, int32_t cin,
                                               int32_t t, int32_t h, int32_t w, int32_t kernel_size, int32_t planes,
                                               int32_t k_splits, float* slices, void* stream_) {
    SS_REQUIRE(dy_planes && x_planes && slices, "conv3d_wgrad_direct: null pointer");
    SS_REQUIRE(planes == 1 || planes == 2, "conv3d_wgrad_direct: planes must be 1 or 2");
    SS_REQUIRE(kernel_size == 3 || kernel_size == 1, "conv3d_wgrad_direct: kernel_size must be 1 or 3");
    SS_REQUIRE(cout >= 8 && cout % 8 == 0 && cin >= 32 && cin % 32 == 0,
               "conv3d_wgrad_direct: cout must be a multiple of 8 and cin of 32 (got %d, %d)", cout, cin);
    SS_REQUIRE(t >= 1 && h >= 1 && w >= 1, "conv3d_wgrad_direct: empty volume");
    SS_REQUIRE((reinterpret_cast<uintptr_t>(dy_planes) & 15) == 0 && (reinterpret_cast<uintptr_t>(x_planes) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(slices) & 15) == 0,
               "conv3d_wgrad_direct: pointers must be 16-byte aligned");
    int rc = require_sm100();
    if (rc != STEMSEG_OK) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WgradParams p;
    p.t = t; p.h = h; p.w = w;
    p.cin = cin; p.cout = cout;
    p.ntaps = kernel_size == 3 ? 27 : 1;
    choose_box32(t, h, w, &p.tt, &p.th, &p.tw);
    p.tiles_t = (t + p.tt - 1) / p.tt;
    p.tiles_h = (h + p.th - 1) / p.th;
    p.tiles_w = (w + p.tw - 1) / p.tw;
    p.k_tiles = p.tiles_t * p.tiles_h * p.tiles_w;
    const int bn = wgrad_direct_block_n(cin);
    p.m_tiles = (cout + kBlockM - 1) / kBlockM;
    p.n_tiles = (cin + bn - 1) / bn;
    SS_REQUIRE(k_splits >= 1 && k_splits <= 64 && k_splits <= p.k_tiles,
               "conv3d_wgrad_direct: k_splits %d out of range (1..min(64, %d))", k_splits, p.k_tiles);
    p.k_splits = k_splits;
    p.num_tiles = p.ntaps * p.m_tiles * p.n_tiles * p.k_splits;
    p.num_stages = 0;
    p.out = slices;
    const size_t vox = static_cast<size_t>(t) * h * w;
    CUtensorMap maps[4];
    for (int pl = 0; pl < 2; ++pl) {
        const int src = pl < planes ? pl : 0;
        rc = encode_box_map(&maps[pl], static_cast<const uint8_t*>(dy_planes) + src * vox * cout * 2, cout, t, h, w, p.tt,
                            p.th, p.tw);
        if (rc != STEMSEG_OK) return rc;
        rc = encode_box_map(&maps[2 + pl], static_cast<const uint8_t*>(x_planes) + src * vox * cin * 2, cin, t, h, w, p.tt,
                            p.th, p.tw);
        if (rc != STEMSEG_OK) return rc;
    }
    if (planes == 2) {
        if (bn == 256) return launch_wgrad<256, 2>(maps, p, stream);
        if (bn == 128) return launch_wgrad<128, 2>(maps, p, stream);
        return launch_wgrad<64, 2>(maps, p, stream);
    }
    if (bn == 256) return launch_wgrad<256, 1>(maps, p, stream);
    if (bn == 128) return launch_wgrad<128, 1>(maps, p, stream);
    return launch_wgrad<64, 1>(maps, p, stream);
}

namespace {
int wgrad_block_k(int cin, int planes) { return planes == 2 && block_n_for(cin) == 256 ? 32 : 64; }
long long wgrad_k_pad(int t, int h, int w, int kernel_size) {
    const int pad = kernel_size == 3 ? 1 : 0;
    const int pitch = pad ? (w + 2 + 7) / 8 * 8 : w;                      // backward_ops.cu: padded_row_pitch
    const long long k_true = 1ll * (t + 2 * pad) * (h + 2 * pad) * pitch;
    return (k_true + 63) / 64 * 64;
}
}  // namespace

extern "C" int32_t stemseg_wgrad_k_splits(int32_t cout, int32_t cin, int32_t t, int32_t h, int32_t w,
                                          int32_t kernel_size, int32_t planes) {
    const int taps = kernel_size == 3 ? 27 : 1;
    const long long tiles = 1ll * ((cout + kBlockM - 1) / kBlockM) * (cin / block_n_for(cin)) * taps;
    long long k = (3ll * device_sm_count() + tiles - 1) / tiles;          // aim at ~3 CTAs' worth of tiles per SM
    const long long k_chunks = wgrad_k_pad(t, h, w, kernel_size) / wgrad_block_k(cin, planes);
    if (k > k_chunks) k = k_chunks;                                       // every split owns at least one K chunk
    if (k < 1) k = 1;
    if (k > 32) k = 32;
    return static_cast<int32_t>(k);
}

// dW partial sums for a stride-1 / pad-(k-1)/2 convolution from the zero-padded transposed planes:
//   slices[ks][tap][co][ci] = sum_{p in K range ks} dyT[co][p] * xT[ci][p + delta(tap)]
extern "C" int32_t stemseg_conv3d_wgrad(const void* dyT_planes, const void* xT_planes, int32_t cout, int32_t cin,
                                        int32_t t, int32_t h, int32_t w, int32_t kernel_size, int32_t planes,
                                        int32_t k_splits, float* slices, void* stream_) {
    SS_REQUIRE(dyT_planes && xT_planes && slices, "conv3d_wgrad: null pointer");
    SS_REQUIRE(planes == 1 || planes == 2, "conv3d_wgrad: planes must be 1 or 2");
    SS_REQUIRE(kernel_size == 3 || kernel_size == 1, "conv3d_wgrad: kernel_size must be 1 or 3");
    SS_REQUIRE(cout >= 1 && cin >= 32 && cin % 32 == 0, "conv3d_wgrad: cin must be a multiple of 32 (got %d)", cin);
    SS_REQUIRE(k_splits >= 1 && k_splits <= 64, "conv3d_wgrad: k_splits out of range");
    int rc = require_sm100();
    if (rc != STEMSEG_OK) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int pad = kernel_size == 3 ? 1 : 0;
    const long long k_pad = wgrad_k_pad(t, h, w, kernel_size);
    SS_REQUIRE(k_pad < 0x7FFFFFFFll, "conv3d_wgrad: volume too large");
    SS_REQUIRE(k_splits <= k_pad / wgrad_block_k(cin, planes),
               "conv3d_wgrad: k_splits %d exceeds the %lld K chunks of this volume (use stemseg_wgrad_k_splits)", k_splits,
               k_pad / wgrad_block_k(cin, planes));

    ConvTcParams p;
    p.n = 1; p.t = 1; p.h = 1; p.w = cout;                 // GEMM rows = output channels of the convolution
    p.cin = static_cast<int>(k_pad);                       // GEMM K = padded voxels
    p.cout = cin;                                          // GEMM N = input channels of the convolution
    p.ntaps = 1;
    choose_box(1, 1, cout, &p.tt, &p.th, &p.tw);
    p.tiles_t = 1; p.tiles_h = 1;
    p.tiles_w = (cout + p.tw - 1) / p.tw;
    const int block_n = block_n_for(cin);
    p.n_tiles_n = cin / block_n;
    p.k_slices = kernel_size == 3 ? 27 : 1;
    p.taps_per_slice = 1;
    p.slice_stride = static_cast<size_t>(cout) * cin;
    p.k_splits = k_splits;
    p.wgrad_mode = 1;
    p.operand_fp16 = 0;
    p.out_bf16 = 0;
    const int hp = h + 2 * pad, pitch = pad ? (w + 2 + 7) / 8 * 8 : w;
    const int shifts = kernel_size == 3 ? 3 : 1;
    for (int tap = 0; tap < 27; ++tap) {
        const int dt = tap / 9 - 1, dh = (tap / 3) % 3 - 1, dw = tap % 3 - 1;
        p.b_k_offset[tap] = kernel_size == 3 ? (dt * hp + dh) * pitch : 0;
        p.b_row_offset[tap] = kernel_size == 3 ? (dw + 1) * cin : 0;
    }
    const long long tiles = 1ll * p.tiles_w * p.n_tiles_n * p.k_slices * p.k_splits;
    p.num_tiles = static_cast<int>(tiles);
    p.out = slices;
    p.bias = nullptr;
    p.num_stages = 0;
    p.tiles_per_cta = 0;
    p.tiles_per_sample = p.tiles_w;
    p.stat_partial = nullptr;
    p.epi_mode = 0;
    p.head_j = 0;
    p.k_chunks_total = 0;

    const int block_k = wgrad_block_k(cin, planes);
    const size_t a_plane_bytes = static_cast<size_t>(cout) * k_pad * 2;
    const size_t b_plane_bytes = static_cast<size_t>(shifts) * cin * k_pad * 2;
    CUtensorMap maps[4];
    const uint8_t* a = static_cast<const uint8_t*>(dyT_planes);
    const uint8_t* b = static_cast<const uint8_t*>(xT_planes);
    for (int pl = 0; pl < 2; ++pl) {
        const int src = pl < planes ? pl : 0;
        rc = encode_act_map(&maps[pl], a + src * a_plane_bytes, p, block_k);
        if (rc != STEMSEG_OK) return rc;
        rc = encode_weight_map(&maps[2 + pl], b + src * b_plane_bytes, k_pad, static_cast<int64_t>(shifts) * cin, block_k,
                               block_n);
        if (rc != STEMSEG_OK) return rc;
    }
    if (planes == 2)
        return block_k == 64 ? launch_by_n<64, 2>(block_n, maps, p, 0, stream) : launch_by_n<32, 2>(block_n, maps, p, 0, stream);
    return block_k == 64 ? launch_by_n<64, 1>(block_n, maps, p, 0, stream) : launch_by_n<32, 1>(block_n, maps, p, 0, stream);
}
