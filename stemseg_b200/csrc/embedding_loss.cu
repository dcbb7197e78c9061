// Embedding loss of one training sequence, forward AND gradient, on the device (sm_100a) -- SURVEY.md §8f rank 3.
//
// Replaces (paths relative to the reference root)
//   EmbeddingLoss.forward                              stemseg/modeling/losses/embedding_loss.py:35-157
//   EmbeddingLoss.compute_prob_map                     embedding_loss.py:159-178
//   EmbeddingLoss.compute_bandwidth_smoothness_loss    embedding_loss.py:180-185
//   lovasz_hinge_flat / lovasz_grad                    stemseg/modeling/losses/_lovasz.py:139-157, :18-31
// and the autograd graph torch builds behind them (the reference evaluates, per instance, a Gaussian probability
// map over ALL voxels, a full descending sort of the hinge errors, a cumulative Jaccard gradient, and -- through
// autograd -- the adjoints of the masked means that define the instance centre and bandwidth).
//
// Layout: the head output of one sequence is channels-first [C][M] (M = T*H*W voxels, the API layout of the heads):
// E embedding rows, V = E - n_free variance rows; seediness is a separate [M] row (last channel of the embedding
// head, or the seediness head's output).  masks [I][M] uint8, ignore [M] uint8.
//
// Pipeline (all on the caller's stream, no host synchronisation; reductions accumulate in double):
//   1 loss_stats      per ORIGINAL instance i: point count, sum of embeddings / variances / 10*exp(variance); background
//                     count and sum of (non-ignored) squared seediness
//   2 loss_prepare    one thread: drop empty instances (kept slot n <- original id), centre mu, bandwidth, target
//                     (the reference pairs kept slot n with masks[n] -- the ORIGINAL index n, quirk preserved)
//   3 loss_prob       per (slot, voxel): p = exp(-0.5 sum (x-mu)^2 bw); hinge error; 64-bit sort key
//                     (error bits << 32 | (voxel+1) << 1 | label); smoothness and instance-seediness partial sums
//   4 bitonic sort    of every active slot's keys, descending (shared-memory chunks of 4096 + global steps)
//   5 lovasz_count / lovasz_apply   positives per chunk, then per sorted position the Jaccard increment
//                     g_i = J_i - J_{i-1} from integer prefix counts; loss += relu(e_i) g_i; dL/dp scattered to voxels
//   6 loss_backward_accumulate   per voxel: direct gradient wrt the embeddings + per-slot sums for dL/dmu, dL/dbw
//   7 loss_backward_distribute   per voxel: the masked-mean adjoints, smoothness and seediness gradients
//   8 loss_finalize   one thread: the three loss terms and their weighted sum
#include "common.cuh"

namespace stemseg {
namespace {

constexpr int kMaxE = 8;                       // embedding dims (reference modes: 2..5, embedding_utils.py:4-14)
constexpr int kMaxI = STEMSEG_MAX_LOSS_INSTANCES;
constexpr int kChunk = 4096;                   // keys sorted / scanned per thread block in shared memory
constexpr int kSortThreads = 512;
constexpr int kKeysPerThread = kChunk / kSortThreads;   // 8

// ---- workspace header (device) ---------------------------------------------------------------------------------
struct LossAcc {                // zeroed before every call
    // per ORIGINAL instance
    double count[kMaxI];
    double sum_emb[kMaxI][kMaxE];
    double sum_var[kMaxI][kMaxE];
    double sum_bw[kMaxI][kMaxE];
    double bg_count, bg_sq;
    // per kept slot
    double smooth_sq[kMaxI];    // sum over the slot's points and variance dims of (mean_var - var)^2
    double seed_sq[kMaxI];      // sum over the slot's points of (seediness - p)^2
    double lovasz[kMaxI];       // sum_i relu(e_i) g_i
    double a_mu[kMaxI][kMaxE];  // dL/dmu_e
    double b_bw[kMaxI][kMaxE];  // dL/dbw_e
};

struct LossSlots {              // written by loss_prepare
    int n_kept;
    int src[kMaxI];             // original instance id whose points define slot n's centre / bandwidth
    int active[kMaxI];          // masks[n] (original index n) is non-empty -> Lovasz + instance seediness terms exist
    float target_count[kMaxI];  // number of points of masks[n]
    float src_count[kMaxI];     // number of points of masks[src[n]]
    float mu[kMaxI][kMaxE];
    float bw[kMaxI][kMaxE];     // mean activated bandwidth, then the free-dim bandwidths
    float mean_var[kMaxI][kMaxE];
};

struct LossDims {
    long long m;                // voxels
    long long n_pad;            // power of two >= max(m, kChunk)
    int e, v, n_inst;
    float free_bw[kMaxE];
    float w_lovasz, w_smooth, w_seed, w;
};

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// block-wide sum of `k` per-thread doubles -> atomicAdd into dst[0..k) (one atomic per value per block)
template <int K>
__device__ __forceinline__ void block_accumulate(double (&vals)[K], double* dst, int k, double* s_red /*[32][K]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = (blockDim.x + 31) >> 5;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < K; ++q) {
        if (q < k) {                                   // k is block-uniform
            const double s = warp_sum(vals[q]);
            if (lane == 0) s_red[warp * K + q] = s;
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < k; q += blockDim.x) {
        double s = 0.0;
        for (int wi = 0; wi < warps; ++wi) s += s_red[wi * K + q];
        if (s != 0.0) atomicAdd(dst + q, s);
    }
}

// ---- 1: statistics ------------------------------------------------------------------------------------------------
// blockIdx.y = original instance id, or n_inst for the background row.
__global__ void __launch_bounds__(256) loss_stats_kernel(const float* __restrict__ out, const float* __restrict__ seed,
                                                         const uint8_t* __restrict__ masks,
                                                         const uint8_t* __restrict__ ignore, LossDims d,
                                                         LossAcc* __restrict__ acc) {
    __shared__ double s_red[8 * (1 + 3 * kMaxE)];
    const int inst = blockIdx.y;
    const long long stride = 1ll * gridDim.x * blockDim.x;
    double vals[1 + 3 * kMaxE];
#pragma unroll
    for (int q = 0; q < 1 + 3 * kMaxE; ++q) vals[q] = 0.0;
    if (inst < d.n_inst) {
        const uint8_t* mk = masks + 1ll * inst * d.m;
        for (long long vx = 1ll * blockIdx.x * blockDim.x + threadIdx.x; vx < d.m; vx += stride) {
            if (mk[vx] == 0) continue;
            vals[0] += 1.0;
#pragma unroll
            for (int e = 0; e < kMaxE; ++e)
                if (e < d.e) vals[1 + e] += static_cast<double>(out[1ll * e * d.m + vx]);
#pragma unroll
            for (int q = 0; q < kMaxE; ++q)
                if (q < d.v) {
                    const float var = out[1ll * (d.e + q) * d.m + vx];
                    vals[1 + kMaxE + q] += static_cast<double>(var);
                    vals[1 + 2 * kMaxE + q] += static_cast<double>(expf(var) * 10.0f);   // embedding_loss.py:121-124
                }
        }
        // LossAcc: count[kMaxI] | sum_emb | sum_var | sum_bw are separate arrays -> accumulate each group
        double cnt[1] = {vals[0]};
        block_accumulate<1>(cnt, acc->count + inst, 1, s_red);
        double g[kMaxE];
#pragma unroll
        for (int e = 0; e < kMaxE; ++e) g[e] = vals[1 + e];
        block_accumulate<kMaxE>(g, acc->sum_emb[inst], d.e, s_red);
#pragma unroll
        for (int e = 0; e < kMaxE; ++e) g[e] = vals[1 + kMaxE + e];
        block_accumulate<kMaxE>(g, acc->sum_var[inst], d.v, s_red);
#pragma unroll
        for (int e = 0; e < kMaxE; ++e) g[e] = vals[1 + 2 * kMaxE + e];
        block_accumulate<kMaxE>(g, acc->sum_bw[inst], d.v, s_red);
    } else {
        // background = no instance at all on the voxel (embedding_loss.py:112); ignored points count in the mean's
        // denominator but contribute zero (:115-116)
        double bg[2] = {0.0, 0.0};
        for (long long vx = 1ll * blockIdx.x * blockDim.x + threadIdx.x; vx < d.m; vx += stride) {
            bool any = false;
            for (int i = 0; i < d.n_inst; ++i) any |= masks[1ll * i * d.m + vx] != 0;
            if (any) continue;
            bg[0] += 1.0;
            if (ignore == nullptr || ignore[vx] == 0) {
                const float s = seed[vx];
                bg[1] += static_cast<double>(s * s);
            }
        }
        block_accumulate<2>(bg, &acc->bg_count, 2, s_red);
    }
}

// ---- 2: slots -----------------------------------------------------------------------------------------------------
__global__ void loss_prepare_kernel(LossDims d, const LossAcc* __restrict__ acc, LossSlots* __restrict__ slots) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int n = 0;
    for (int i = 0; i < d.n_inst; ++i) {
        if (acc->count[i] <= 0.0) continue;                    // `unique` over the nonzero points (:83-87)
        slots->src[n] = i;
        const double c = acc->count[i];
        slots->src_count[n] = static_cast<float>(c);
        for (int e = 0; e < d.e; ++e) slots->mu[n][e] = static_cast<float>(acc->sum_emb[i][e] / c);
        for (int q = 0; q < d.v; ++q) {
            slots->mean_var[n][q] = static_cast<float>(acc->sum_var[i][q] / c);
            slots->bw[n][q] = static_cast<float>(acc->sum_bw[i][q] / c);
        }
        for (int q = d.v; q < d.e; ++q) slots->bw[n][q] = d.free_bw[q - d.v];      // embedding_loss.py:170-171
        ++n;
    }
    slots->n_kept = n;
    for (int s = 0; s < kMaxI; ++s) {
        // slot s is paired with masks[s], the ORIGINAL index s (embedding_loss.py:128)
        const double tc = (s < n) ? acc->count[s] : 0.0;
        slots->target_count[s] = static_cast<float>(tc);
        slots->active[s] = (s < n && tc > 0.0) ? 1 : 0;                             // :129-130
    }
}

// ---- 3: probabilities, hinge errors, sort keys ------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_prob_kernel(const float* __restrict__ out, const float* __restrict__ seed,
                                                        const uint8_t* __restrict__ masks, LossDims d,
                                                        const LossSlots* __restrict__ slots, LossAcc* __restrict__ acc,
                                                        float* __restrict__ probs /*[I][m]*/,
                                                        unsigned long long* __restrict__ keys /*[I][n_pad]*/) {
    __shared__ double s_red[8 * 2];
    const int slot = blockIdx.y;
    if (slot >= slots->n_kept) return;
    const int src = slots->src[slot];
    const bool active = slots->active[slot] != 0;
    float mu[kMaxE], bw[kMaxE], mv[kMaxE];
#pragma unroll
    for (int e = 0; e < kMaxE; ++e) {
        mu[e] = e < d.e ? slots->mu[slot][e] : 0.f;
        bw[e] = e < d.e ? slots->bw[slot][e] : 0.f;
        mv[e] = e < d.v ? slots->mean_var[slot][e] : 0.f;
    }
    const uint8_t* src_mask = masks + 1ll * src * d.m;
    const uint8_t* tgt_mask = masks + 1ll * slot * d.m;
    unsigned long long* krow = keys + 1ll * slot * d.n_pad;
    float* prow = probs + 1ll * slot * d.m;
    double part[2] = {0.0, 0.0};                               // smoothness, instance seediness
    const long long stride = 1ll * gridDim.x * blockDim.x;
    for (long long vx = 1ll * blockIdx.x * blockDim.x + threadIdx.x; vx < d.n_pad; vx += stride) {
        if (vx >= d.m) {
            if (active) krow[vx] = 0ull;                       // padding sorts strictly after every real key
            continue;
        }
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < kMaxE; ++e)
            if (e < d.e) {
                const float df = __fsub_rn(out[1ll * e * d.m + vx], mu[e]);
                q = __fadd_rn(q, __fmul_rn(__fmul_rn(df, df), bw[e]));
            }
        const float p = expf(__fmul_rn(-0.5f, q));
        prow[vx] = p;
        if (src_mask[vx] != 0) {
#pragma unroll
            for (int k = 0; k < kMaxE; ++k)
                if (k < d.v) {
                    const float dv = mv[k] - out[1ll * (d.e + k) * d.m + vx];
                    part[0] += static_cast<double>(dv * dv);
                }
            if (active) {
                const float ds = seed[vx] - p;
                part[1] += static_cast<double>(ds * ds);
            }
        }
        if (active) {
            const unsigned label = tgt_mask[vx] != 0 ? 1u : 0u;
            const float logit = __fsub_rn(__fmul_rn(p, 2.0f), 1.0f);                  // embedding_loss.py:127
            const float err = label ? __fsub_rn(1.0f, logit) : __fadd_rn(1.0f, logit);   // _lovasz.py:150-151
            // err is in [0, 2]: non-negative floats order like their bit patterns
            krow[vx] = (static_cast<unsigned long long>(__float_as_uint(fmaxf(err, 0.f))) << 32) |
                       ((static_cast<unsigned long long>(vx) + 1ull) << 1) | label;
        }
    }
    {
        double a[1] = {part[0]};
        block_accumulate<1>(a, acc->smooth_sq + slot, 1, s_red);
        double b[1] = {part[1]};
        block_accumulate<1>(b, acc->seed_sq + slot, 1, s_red);
    }
}

// ---- 4: bitonic sort (descending) ------------------------------------------------------------------------------------
// Direction of the compare-exchange of elements (i, i^j) in the merge of size k: the network below sorts ASCENDING
// when (i & k) == 0; we want the final order DESCENDING, so the comparison is flipped.
__device__ __forceinline__ void cmp_swap(unsigned long long& a, unsigned long long& b, bool descending) {
    if ((a < b) == descending) {
        const unsigned long long t = a;
        a = b;
        b = t;
    }
}

// sorts each chunk of kChunk keys completely (k = 2 .. kChunk); chunk c ends up descending if ((c*kChunk) & kChunk)==0
// in the sense of the global network (so that the following global merges see bitonic sequences)
__global__ void __launch_bounds__(kSortThreads) bitonic_local_sort_kernel(unsigned long long* __restrict__ keys,
                                                                            long long n_pad,
                                                                            const LossSlots* __restrict__ slots) {
    __shared__ unsigned long long s[kChunk];
    const int slot = blockIdx.y;
    if (!slots->active[slot]) return;
    unsigned long long* row = keys + 1ll * slot * n_pad + 1ll * blockIdx.x * kChunk;
    const long long base = 1ll * blockIdx.x * kChunk;
    for (int i = threadIdx.x; i < kChunk; i += kSortThreads) s[i] = row[i];
    __syncthreads();
    for (int k = 2; k <= kChunk; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < kChunk / 2; t += kSortThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const bool desc = ((base + i) & k) == 0;
                cmp_swap(s[i], s[i | j], desc);
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < kChunk; i += kSortThreads) row[i] = s[i];
}

// one compare-exchange step (k, j) with j >= kChunk over the whole row
__global__ void __launch_bounds__(256) bitonic_global_step_kernel(unsigned long long* __restrict__ keys, long long n_pad,
                                                                  long long k, long long j,
                                                                  const LossSlots* __restrict__ slots) {
    const int slot = blockIdx.y;
    if (!slots->active[slot]) return;
    unsigned long long* row = keys + 1ll * slot * n_pad;
    const long long stride = 1ll * gridDim.x * blockDim.x;
    for (long long t = 1ll * blockIdx.x * blockDim.x + threadIdx.x; t < n_pad / 2; t += stride) {
        const long long i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        unsigned long long a = row[i], b = row[i | j];
        const bool desc = (i & k) == 0;
        if ((a < b) == desc) {
            row[i] = b;
            row[i | j] = a;
        }
    }
}

// the steps j = kChunk/2 .. 1 of the merge of size k (> kChunk), chunk-local
__global__ void __launch_bounds__(kSortThreads) bitonic_local_merge_kernel(unsigned long long* __restrict__ keys,
                                                                             long long n_pad, long long k,
                                                                             const LossSlots* __restrict__ slots) {
    __shared__ unsigned long long s[kChunk];
    const int slot = blockIdx.y;
    if (!slots->active[slot]) return;
    const long long base = 1ll * blockIdx.x * kChunk;
    unsigned long long* row = keys + 1ll * slot * n_pad + base;
    for (int i = threadIdx.x; i < kChunk; i += kSortThreads) s[i] = row[i];
    __syncthreads();
    const bool desc = (base & k) == 0;                 // constant over the chunk because k > kChunk
    for (int j = kChunk >> 1; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < kChunk / 2; t += kSortThreads) {
            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            cmp_swap(s[i], s[i | j], desc);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < kChunk; i += kSortThreads) row[i] = s[i];
}

// ---- 5: Lovasz gradient --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads) lovasz_count_kernel(const unsigned long long* __restrict__ keys,
                                                                      long long n_pad, const LossSlots* __restrict__ slots,
                                                                      int* __restrict__ chunk_pos /*[I][n_pad/kChunk]*/) {
    __shared__ int s_cnt[kSortThreads / 32];
    const int slot = blockIdx.y;
    if (!slots->active[slot]) return;
    const unsigned long long* row = keys + 1ll * slot * n_pad + 1ll * blockIdx.x * kChunk;
    int c = 0;
    for (int i = threadIdx.x; i < kChunk; i += kSortThreads) c += static_cast<int>(row[i] & 1ull);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int wi = 0; wi < kSortThreads / 32; ++wi) tot += s_cnt[wi];
        chunk_pos[1ll * slot * (n_pad / kChunk) + blockIdx.x] = tot;
    }
}

// Jaccard index after the first (pos + neg) sorted elements, exactly as lovasz_grad (_lovasz.py:18-31) computes it in
// fp32 from integer-valued floats: 1 - (P - cum_pos) / (P + cum_neg)
__device__ __forceinline__ float jaccard_at(float p_total, int cum_pos, int cum_neg) {
    const float inter = __fsub_rn(p_total, static_cast<float>(cum_pos));
    const float uni = __fadd_rn(p_total, static_cast<float>(cum_neg));
    return __fsub_rn(1.0f, __fdiv_rn(inter, uni));
}

__global__ void __launch_bounds__(kSortThreads) lovasz_apply_kernel(const unsigned long long* __restrict__ keys,
                                                                      long long n_pad, long long m,
                                                                      const LossSlots* __restrict__ slots,
                                                                      const int* __restrict__ chunk_pos,
                                                                      LossAcc* __restrict__ acc,
                                                                      float* __restrict__ dldp /*[I][m]*/) {
    __shared__ int s_warp[kSortThreads / 32];
    __shared__ int s_prefix;
    __shared__ double s_red[kSortThreads / 32];
    const int slot = blockIdx.y;
    if (!slots->active[slot]) return;
    const long long base = 1ll * blockIdx.x * kChunk;
    if (base >= m) return;                              // the whole chunk is padding
    const int n_chunks = static_cast<int>(n_pad / kChunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // positives in the chunks before this one
    int before = 0;
    for (int c = threadIdx.x; c < static_cast<int>(blockIdx.x); c += kSortThreads) before += chunk_pos[1ll * slot * n_chunks + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) s_warp[warp] = before;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int wi = 0; wi < kSortThreads / 32; ++wi) tot += s_warp[wi];
        s_prefix = tot;
    }
    __syncthreads();
    const int prefix = s_prefix;
    __syncthreads();
    // thread t owns the kKeysPerThread consecutive sorted positions base + t*kKeysPerThread ...
    const unsigned long long* row = keys + 1ll * slot * n_pad + base + 1ll * threadIdx.x * kKeysPerThread;
    unsigned long long kk[kKeysPerThread];
    int local = 0;
#pragma unroll
    for (int q = 0; q < kKeysPerThread; ++q) {
        kk[q] = row[q];
        local += static_cast<int>(kk[q] & 1ull);
    }
    // exclusive scan of `local` over the block
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int warp_off = 0;
    for (int wi = 0; wi < warp; ++wi) warp_off += s_warp[wi];
    int cum_pos = prefix + warp_off + incl - local;     // positives strictly before this thread's first position
    const float p_total = slots->target_count[slot];
    float* drow = dldp + 1ll * slot * m;
    double loss = 0.0;
#pragma unroll
    for (int q = 0; q < kKeysPerThread; ++q) {
        const unsigned long long key = kk[q];
        if (key == 0ull) continue;                      // padding
        const long long pos = base + 1ll * threadIdx.x * kKeysPerThread + q;       // 0-based sorted position
        const int label = static_cast<int>(key & 1ull);
        const long long vx = static_cast<long long>((key & 0xFFFFFFFFull) >> 1) - 1;
        const float err = __uint_as_float(static_cast<unsigned>(key >> 32));
        const int pos_before = cum_pos, neg_before = static_cast<int>(pos) - cum_pos;
        cum_pos += label;
        const int neg_after = static_cast<int>(pos) + 1 - cum_pos;
        const float j_now = jaccard_at(p_total, cum_pos, neg_after);
        const float j_prev = pos == 0 ? 0.f : jaccard_at(p_total, pos_before, neg_before);
        const float g = __fsub_rn(j_now, j_prev);
        float dp = 0.f;
        if (err > 0.f) {                                // relu'(0) = 0
            loss += static_cast<double>(err) * static_cast<double>(g);
            dp = label ? -2.0f * g : 2.0f * g;          // d err / d p = -2 sign
        }
        drow[vx] = dp;
    }
    loss = warp_sum(loss);
    if (lane == 0) s_red[warp] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int wi = 0; wi < kSortThreads / 32; ++wi) tot += s_red[wi];
        if (tot != 0.0) atomicAdd(acc->lovasz + slot, tot);
    }
}

// ---- 6/7: gradient -------------------------------------------------------------------------------------------------
// Direct term: dL/dx_e[v] = sum_slots g p (-(x_e - mu_e) bw_e) with g = dL/dp scaled by w w_lovasz / n_kept;
// per-slot sums A_e = sum_v g p (x_e - mu_e) bw_e (= dL/dmu_e) and B_e = sum_v g p (-0.5)(x_e - mu_e)^2 (= dL/dbw_e).
__global__ void __launch_bounds__(256) loss_backward_accumulate_kernel(const float* __restrict__ out, LossDims d,
                                                                        const LossSlots* __restrict__ slots,
                                                                        const float* __restrict__ probs,
                                                                        const float* __restrict__ dldp,
                                                                        LossAcc* __restrict__ acc,
                                                                        float* __restrict__ d_out /*[e+v][m]*/) {
    __shared__ double s_red[8 * 2 * kMaxE];
    const int n_kept = slots->n_kept;
    const float lov_scale = n_kept > 0 ? d.w * d.w_lovasz / static_cast<float>(n_kept) : 0.f;
    const long long stride = 1ll * gridDim.x * blockDim.x;
    for (int slot = 0; slot <= n_kept; ++slot) {
        // slot == n_kept: the pass that writes the accumulated direct term (keeps one pass per slot simple and the
        // per-slot block reductions uniform)
        if (slot < n_kept && !slots->active[slot]) continue;
        double sums[2 * kMaxE];
#pragma unroll
        for (int q = 0; q < 2 * kMaxE; ++q) sums[q] = 0.0;
        if (slot < n_kept) {
            float mu[kMaxE], bw[kMaxE];
#pragma unroll
            for (int e = 0; e < kMaxE; ++e) {
                mu[e] = e < d.e ? slots->mu[slot][e] : 0.f;
                bw[e] = e < d.e ? slots->bw[slot][e] : 0.f;
            }
            for (long long vx = 1ll * blockIdx.x * blockDim.x + threadIdx.x; vx < d.m; vx += stride) {
                const float t = dldp[1ll * slot * d.m + vx] * lov_scale * probs[1ll * slot * d.m + vx];
                if (t == 0.f) continue;
#pragma unroll
                for (int e = 0; e < kMaxE; ++e)
                    if (e < d.e) {
                        const float df = out[1ll * e * d.m + vx] - mu[e];
                        sums[e] += static_cast<double>(t * df * bw[e]);
                        sums[kMaxE + e] += static_cast<double>(-0.5f * t * df * df);
                    }
            }
            double a[kMaxE], b[kMaxE];
#pragma unroll
            for (int e = 0; e < kMaxE; ++e) {
                a[e] = sums[e];
                b[e] = sums[kMaxE + e];
            }
            block_accumulate<kMaxE>(a, acc->a_mu[slot], d.e, s_red);
            block_accumulate<kMaxE>(b, acc->b_bw[slot], d.e, s_red);
        } else {
            for (long long vx = 1ll * blockIdx.x * blockDim.x + threadIdx.x; vx < d.m; vx += stride) {
                float g[kMaxE];
#pragma unroll
                for (int e = 0; e < kMaxE; ++e) g[e] = 0.f;
                for (int s = 0; s < n_kept; ++s) {
                    if (!slots->active[s]) continue;
                    const float t = dldp[1ll * s * d.m + vx] * lov_scale * probs[1ll * s * d.m + vx];
                    if (t == 0.f) continue;
#pragma unroll
                    for (int e = 0; e < kMaxE; ++e)
                        if (e < d.e) g[e] -= t * (out[1ll * e * d.m + vx] - slots->mu[s][e]) * slots->bw[s][e];
                }
#pragma unroll
                for (int e = 0; e < kMaxE; ++e)
                    if (e < d.e) d_out[1ll * e * d.m + vx] = g[e];
            }
        }
    }
}

// Masked-mean adjoints + smoothness + seediness.  Runs after the accumulate kernel (same stream).
__global__ void __launch_bounds__(256) loss_backward_distribute_kernel(const float* __restrict__ out,
                                                                        const float* __restrict__ seed,
                                                                        const uint8_t* __restrict__ masks,
                                                                        const uint8_t* __restrict__ ignore, LossDims d,
                                                                        const LossSlots* __restrict__ slots,
                                                                        const LossAcc* __restrict__ acc,
                                                                        const float* __restrict__ probs,
                                                                        float* __restrict__ d_out, float* __restrict__ d_seed) {
    const int n_kept = slots->n_kept;
    const float smooth_scale = n_kept > 0 ? d.w * d.w_smooth / static_cast<float>(n_kept) : 0.f;   // batch of one
    const float seed_scale = n_kept > 0 ? d.w * d.w_seed / static_cast<float>(n_kept + 1) : 0.f;
    const float bg_count = static_cast<float>(acc->bg_count);
    const long long stride = 1ll * gridDim.x * blockDim.x;
    for (long long vx = 1ll * blockIdx.x * blockDim.x + threadIdx.x; vx < d.m; vx += stride) {
        float ge[kMaxE], gv[kMaxE];
#pragma unroll
        for (int e = 0; e < kMaxE; ++e) {
            ge[e] = 0.f;
            gv[e] = 0.f;
        }
        float gs = 0.f;
        const float s = seed[vx];
        bool any = false;
        for (int i = 0; i < d.n_inst; ++i) any |= masks[1ll * i * d.m + vx] != 0;
        if (!any && n_kept > 0 && (ignore == nullptr || ignore[vx] == 0)) gs += seed_scale * 2.0f * s / bg_count;
        for (int slot = 0; slot < n_kept; ++slot) {
            if (masks[1ll * slots->src[slot] * d.m + vx] == 0) continue;
            const float c = slots->src_count[slot];
            const bool active = slots->active[slot] != 0;
#pragma unroll
            for (int e = 0; e < kMaxE; ++e)
                if (e < d.e && active) ge[e] += static_cast<float>(acc->a_mu[slot][e]) / c;
#pragma unroll
            for (int q = 0; q < kMaxE; ++q)
                if (q < d.v) {
                    const float var = out[1ll * (d.e + q) * d.m + vx];
                    if (active) gv[q] += static_cast<float>(acc->b_bw[slot][q]) * (10.0f * expf(var)) / c;
                    gv[q] += smooth_scale * 2.0f * (var - slots->mean_var[slot][q]) / (c * static_cast<float>(d.v));
                }
            if (active) gs += seed_scale * 2.0f * (s - probs[1ll * slot * d.m + vx]) / c;
        }
#pragma unroll
        for (int e = 0; e < kMaxE; ++e)
            if (e < d.e) d_out[1ll * e * d.m + vx] += ge[e];
#pragma unroll
        for (int q = 0; q < kMaxE; ++q)
            if (q < d.v) d_out[1ll * (d.e + q) * d.m + vx] = gv[q];
        d_seed[vx] = gs;
    }
}

// ---- 8: loss terms -------------------------------------------------------------------------------------------------
__global__ void loss_finalize_kernel(LossDims d, const LossSlots* __restrict__ slots, const LossAcc* __restrict__ acc,
                                     float* __restrict__ losses /*[4]: total, lovasz, variance_smoothness, seediness*/) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n = slots->n_kept;
    if (n == 0) {                                              // embedding_loss.py:136-140
        losses[0] = losses[1] = losses[2] = losses[3] = 0.f;
        return;
    }
    double lov = 0.0, smooth = 0.0, seedl = acc->bg_sq / acc->bg_count;      // 0/0 = NaN like the reference's empty mean
    for (int s = 0; s < n; ++s) {
        const double c = static_cast<double>(slots->src_count[s]);
        smooth += acc->smooth_sq[s] / (c * d.v);
        if (slots->active[s]) {
            lov += acc->lovasz[s];
            seedl += acc->seed_sq[s] / c;
        }
    }
    lov /= n;
    smooth /= n;                                               // mean over the kept instances, then / batch size 1
    seedl /= (n + 1);
    losses[1] = static_cast<float>(lov);
    losses[2] = static_cast<float>(smooth);
    losses[3] = static_cast<float>(seedl);
    losses[0] = static_cast<float>((lov * d.w_lovasz + smooth * d.w_smooth + seedl * d.w_seed) * d.w);
}

__global__ void scale_by_device_scalar_kernel(float* __restrict__ x, long long n, const float* __restrict__ scalar) {
    const float s = *scalar;
    const long long stride = 1ll * gridDim.x * blockDim.x;
    for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= s;
}

inline long long next_pow2(long long x) {
    long long p = 1;
    while (p < x) p <<= 1;
    return p;
}

struct LossLayout {
    long long n_pad;
    size_t off_acc, off_slots, off_probs, off_dldp, off_keys, off_chunks, total;
};

LossLayout loss_layout(long long m, int n_inst) {
    LossLayout L;
    L.n_pad = next_pow2(m < kChunk ? kChunk : m);
    size_t o = 0;
    L.off_acc = o;    o += align_up(sizeof(LossAcc), 256);
    L.off_slots = o;  o += align_up(sizeof(LossSlots), 256);
    L.off_probs = o;  o += align_up(static_cast<size_t>(n_inst) * m * sizeof(float), 256);
    L.off_dldp = o;   o += align_up(static_cast<size_t>(n_inst) * m * sizeof(float), 256);
    L.off_keys = o;   o += align_up(static_cast<size_t>(n_inst) * L.n_pad * sizeof(unsigned long long), 256);
    L.off_chunks = o; o += align_up(static_cast<size_t>(n_inst) * (L.n_pad / kChunk) * sizeof(int), 256);
    L.total = o;
    return L;
}

}  // namespace
}  // namespace stemseg

using namespace stemseg;

extern "C" size_t stemseg_embedding_loss_workspace_bytes(int64_t voxels, int32_t n_instances) {
    if (voxels < 1 || n_instances < 0 || n_instances > kMaxI) return 0;
    return loss_layout(voxels, n_instances < 1 ? 1 : n_instances).total;
}

extern "C" int32_t stemseg_embedding_loss(const float* head_out, const float* seediness, const uint8_t* masks,
                                          const uint8_t* ignore, int64_t voxels, int32_t n_instances,
                                          int32_t embedding_dims, int32_t n_free_dims, const float* free_dim_stds,
                                          float w_lovasz, float w_variance_smoothness, float w_seediness, float w,
                                          float* losses, float* d_head_out, float* d_seediness, void* workspace,
                                          size_t workspace_bytes, void* stream_) {
    SS_REQUIRE(head_out && seediness && losses && d_head_out && d_seediness && workspace, "embedding_loss: null pointer");
    SS_REQUIRE(voxels >= 1 && voxels < (1ll << 30), "embedding_loss: voxels out of range");
    SS_REQUIRE(n_instances >= 0 && n_instances <= kMaxI, "embedding_loss: at most %d instances (got %d)", kMaxI, n_instances);
    SS_REQUIRE(masks != nullptr || n_instances == 0, "embedding_loss: masks is null");
    SS_REQUIRE(embedding_dims >= 1 && embedding_dims <= kMaxE && n_free_dims >= 0 && n_free_dims < embedding_dims,
               "embedding_loss: embedding_dims %d / n_free_dims %d out of range", embedding_dims, n_free_dims);
    SS_REQUIRE(n_free_dims == 0 || free_dim_stds != nullptr, "embedding_loss: free_dim_stds is null");
    int rc = require_sm100();
    if (rc != STEMSEG_OK) return rc;
    const int ni = n_instances < 1 ? 1 : n_instances;
    const LossLayout L = loss_layout(voxels, ni);
    if (workspace_bytes < L.total) {
        set_error("embedding_loss: workspace %zu < %zu bytes", workspace_bytes, L.total);
        return STEMSEG_ERR_WORKSPACE;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    LossAcc* acc = reinterpret_cast<LossAcc*>(ws + L.off_acc);
    LossSlots* slots = reinterpret_cast<LossSlots*>(ws + L.off_slots);
    float* probs = reinterpret_cast<float*>(ws + L.off_probs);
    float* dldp = reinterpret_cast<float*>(ws + L.off_dldp);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws + L.off_keys);
    int* chunks = reinterpret_cast<int*>(ws + L.off_chunks);

    LossDims d;
    d.m = voxels;
    d.n_pad = L.n_pad;
    d.e = embedding_dims;
    d.v = embedding_dims - n_free_dims;
    d.n_inst = n_instances;
    for (int q = 0; q < kMaxE; ++q) d.free_bw[q] = 0.f;
    for (int q = 0; q < n_free_dims; ++q) {
        SS_REQUIRE(free_dim_stds[q] > 0.f, "embedding_loss: free_dim_stds must be positive");
        d.free_bw[q] = 1.0f / (free_dim_stds[q] * free_dim_stds[q]);                      // embedding_loss.py:29
    }
    d.w_lovasz = w_lovasz; d.w_smooth = w_variance_smoothness; d.w_seed = w_seediness; d.w = w;

    SS_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(LossAcc), stream));
    const int sms = device_sm_count();
    long long bx = (voxels + 255) / 256;
    if (bx > 4ll * sms) bx = 4ll * sms;
    const unsigned gx = static_cast<unsigned>(bx < 1 ? 1 : bx);
    loss_stats_kernel<<<dim3(gx, n_instances + 1), 256, 0, stream>>>(head_out, seediness, masks, ignore, d, acc);
    loss_prepare_kernel<<<1, 32, 0, stream>>>(d, acc, slots);
    if (n_instances > 0) {
        long long bp = (L.n_pad + 255) / 256;
        if (bp > 4ll * sms) bp = 4ll * sms;
        loss_prob_kernel<<<dim3(static_cast<unsigned>(bp), n_instances), 256, 0, stream>>>(head_out, seediness, masks, d, slots,
                                                                                            acc, probs, keys);
        const unsigned n_chunks = static_cast<unsigned>(L.n_pad / kChunk);
        bitonic_local_sort_kernel<<<dim3(n_chunks, n_instances), kSortThreads, 0, stream>>>(keys, L.n_pad, slots);
        long long bg = (L.n_pad / 2 + 255) / 256;
        if (bg > 8ll * sms) bg = 8ll * sms;
        for (long long k = 2ll * kChunk; k <= L.n_pad; k <<= 1) {
            for (long long j = k >> 1; j >= kChunk; j >>= 1)
                bitonic_global_step_kernel<<<dim3(static_cast<unsigned>(bg), n_instances), 256, 0, stream>>>(keys, L.n_pad, k, j,
                                                                                                              slots);
            bitonic_local_merge_kernel<<<dim3(n_chunks, n_instances), kSortThreads, 0, stream>>>(keys, L.n_pad, k, slots);
        }
        lovasz_count_kernel<<<dim3(n_chunks, n_instances), kSortThreads, 0, stream>>>(keys, L.n_pad, slots, chunks);
        lovasz_apply_kernel<<<dim3(n_chunks, n_instances), kSortThreads, 0, stream>>>(keys, L.n_pad, voxels, slots, chunks,
                                                                                      acc, dldp);
    }
    loss_backward_accumulate_kernel<<<gx, 256, 0, stream>>>(head_out, d, slots, probs, dldp, acc, d_head_out);
    loss_backward_distribute_kernel<<<gx, 256, 0, stream>>>(head_out, seediness, masks, ignore, d, slots, acc, probs,
                                                            d_head_out, d_seediness);
    loss_finalize_kernel<<<1, 32, 0, stream>>>(d, slots, acc, losses);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_scale_by_device_scalar(float* x, int64_t n, const float* scalar, void* stream_) {
    SS_REQUIRE(x && scalar && n >= 0, "scale_by_device_scalar: bad arguments");
    if (n == 0) return STEMSEG_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    long long b = (n + 255) / 256;
    const long long cap = 8ll * device_sm_count();
    if (b > cap) b = cap;
    scale_by_device_scalar_kernel<<<static_cast<unsigned>(b), 256, 0, stream>>>(x, n, scalar);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}
