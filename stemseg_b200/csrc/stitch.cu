// Device side of the sub-clip stitch (sm_100a): label pair histograms and LUT relabelling.
//
// Replaces the per-pair mask reductions of OnlineChainer.associate_clusters (stemseg/inference/online_chainer.py:
// 315-328: K1*K2 times `(l1_active & l2_active).sum()` / `(l1_active | l2_active).sum()` over the overlap points),
// the per-association `torch.where` relabelling (online_chainer.py:219-224) and the per-frame `unique` / count loops of
// TrackContainer.get_track_mask_idxes (online_chainer.py:94-117).  Integer work, one pass over the label vectors;
// the <= 21 x 21 count table goes to the host for the Hungarian solve (scipy, microseconds).
#include "common.cuh"

namespace stemseg {
namespace {

constexpr int kHistThreads = 256;
constexpr int kSmemBins = 8192;      // 32 KB of int32 counters

// bin(v) = 0 for negative labels (outliers), v - base + 1 otherwise; values outside [0, nbins) are a caller error
__device__ __forceinline__ int label_bin(long long v, long long base, int nbins) {
    if (v < 0) return 0;
    const long long b = v - base + 1;
    return (b < 1 || b >= nbins) ? -1 : static_cast<int>(b);
}

__global__ void __launch_bounds__(kHistThreads) pair_histogram_kernel(const long long* __restrict__ a,
                                                                      const long long* __restrict__ b, long long n,
                                                                      long long a_base, long long b_base, int na, int nb,
                                                                      int* __restrict__ table, int* __restrict__ bad) {
    extern __shared__ int s_bins[];
    const int bins = na * nb;
    const bool use_smem = bins <= kSmemBins;
    if (use_smem) {
        for (int i = threadIdx.x; i < bins; i += kHistThreads) s_bins[i] = 0;
        __syncthreads();
    }
    for (long long i = blockIdx.x * 1ll * kHistThreads + threadIdx.x; i < n; i += 1ll * gridDim.x * kHistThreads) {
        const int ia = label_bin(a[i], a_base, na), ib = label_bin(b[i], b_base, nb);
        if (ia < 0 || ib < 0) {
            atomicAdd(bad, 1);
            continue;
        }
        if (use_smem) atomicAdd(&s_bins[ia * nb + ib], 1);
        else atomicAdd(&table[ia * nb + ib], 1);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < bins; i += kHistThreads)
            if (s_bins[i] != 0) atomicAdd(&table[i], s_bins[i]);
    }
}

__global__ void __launch_bounds__(256) relabel_lut_kernel(long long* __restrict__ labels, long long n, long long base,
                                                          const long long* __restrict__ lut, int nlut) {
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
        const long long v = labels[i];
        if (v < 0) continue;
        const long long k = v - base;
        if (k >= 0 && k < nlut) labels[i] = lut[k];
    }
}

}  // namespace
}  // namespace stemseg

using namespace stemseg;

extern "C" int32_t stemseg_label_pair_histogram(const int64_t* a, const int64_t* b, int64_t n, int64_t a_base,
                                                int64_t b_base, int32_t na, int32_t nb, int32_t* table,
                                                int32_t* out_of_range, void* stream_) {
    SS_REQUIRE(table && out_of_range, "label_pair_histogram: null output");
    SS_REQUIRE(na >= 1 && nb >= 1 && 1ll * na * nb <= (1 << 24), "label_pair_histogram: bad table size %d x %d", na, nb);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SS_CUDA_OK(cudaMemsetAsync(table, 0, sizeof(int32_t) * na * nb, stream));
    SS_CUDA_OK(cudaMemsetAsync(out_of_range, 0, sizeof(int32_t), stream));
    if (n == 0) return STEMSEG_OK;
    SS_REQUIRE(a && b && n > 0, "label_pair_histogram: null input");
    long long blocks = (n + kHistThreads * 8 - 1) / (kHistThreads * 8);
    const long long cap = 2ll * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const size_t smem = (1ll * na * nb <= kSmemBins) ? sizeof(int) * na * nb : 0;
    pair_histogram_kernel<<<static_cast<unsigned>(blocks), kHistThreads, smem, stream>>>(
        reinterpret_cast<const long long*>(a), reinterpret_cast<const long long*>(b), n, a_base, b_base, na, nb, table,
        out_of_range);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

extern "C" int32_t stemseg_relabel_lut(int64_t* labels, int64_t n, int64_t base, const int64_t* lut, int32_t nlut,
                                       void* stream_) {
    if (n == 0) return STEMSEG_OK;
    SS_REQUIRE(labels && lut && n > 0 && nlut >= 1, "relabel_lut: bad arguments");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    long long blocks = (n + 255) / 256;
    const long long cap = 8ll * device_sm_count();
    if (blocks > cap) blocks = cap;
    relabel_lut_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(reinterpret_cast<long long*>(labels), n, base,
                                                                          reinterpret_cast<const long long*>(lut), nlut);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Device-resident sequential stitch (OnlineChainer.process, online_chainer.py:162-236, for one sub-clip per call):
//   stitch_hist    one pass over the sub-clip's points: per-frame label counts + the joint histogram
//                  (existing track label x current local label) over the overlap frames
//   stitch_assign  one thread: label lists in CPython set order, IoU costs in the reference's fp32/fp64 arithmetic,
//                  scipy's assignment algorithm (assoc.cuh), relabelling tables, TrackContainer statistics
//                  (get_track_mask_idxes, online_chainer.py:94-117), next_track_label -- all kept on the device
//   stitch_relabel one pass: local labels -> track ids, non-overlap frames stored into the container
// Nothing is read back per sub-clip; the caller fetches the error word / statistics once at the end of the video.
// ---------------------------------------------------------------------------------------------------------------
#include "assoc.cuh"

namespace stemseg {
namespace {

constexpr int kStitchMaxFrames = 64;
constexpr int kStitchThreads = 256;

enum StitchError {
    kErrLabelRange = 1, kErrOverlapSize = 2, kErrLabelsOverlap = 4, kErrTooManyLabels = 8, kErrFrameExists = 16,
    kErrAssignment = 32
};

struct StitchFrames {
    int n_frames;
    int frame[kStitchMaxFrames];         // global frame number of every slot
    int overlap[kStitchMaxFrames];       // 1: the frame is already in the container (shared with the previous sub-clip)
};

struct StitchArgs {
    long long* labels;                   // [>= total points] local labels (start 1), rewritten to the returned labels
    const int* frame_counts;             // device [n_frames]
    const int* k_dev;                    // device: clusters of this sub-clip
    int max_instances;
    long long* frame_labels;             // container [num_frames][frame_capacity]
    int* frame_count;                    // container [num_frames], -1 = no labels yet
    long long frame_capacity;
    int* state;                          // [0] next_track_label [1] highest [2] error flags [3] sub-clips done
    long long* track_counts;             // [max_labels + 1], index = label + 1
    int* span_lo;                        // [max_labels + 1]
    int* span_hi;
    int max_labels;
    long long* meta_labels;              // [max_instances] out: track id of every local cluster ordinal
    // workspace
    int* joint;                          // [(max_labels + 1)][max_instances + 2]
    int* per_frame;                      // [max_instances + 2][n_frames]
    long long* lut_assoc;                // [max_instances + 1]
    long long* lut_local;                // [max_instances + 1]
    double* cost;                        // [kAssocMaxSide * kAssocMaxSide]
    LsapScratch* lsap;
    int is_first;
};

__device__ __forceinline__ int find_slot(const long long* starts, int n_frames, long long p) {
    int lo = 0, hi = n_frames - 1;
    while (lo < hi) {                    // last slot with starts[slot] <= p
        const int mid = (lo + hi + 1) >> 1;
        if (starts[mid] <= p) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// warp-aggregated increment: lanes that hit the same bin elect one leader that adds their count (spatially coherent
// labels make whole warps hit one bin)
__device__ __forceinline__ void hist_add(int* table, int bin, bool active) {
    const unsigned mask = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const unsigned peers = __match_any_sync(mask, bin);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(table + bin, __popc(peers));
}

// per-frame label counts [max_instances + 2][n_frames] and the joint histogram [max_labels + 1][max_instances + 2] are
// accumulated in shared memory per block (when they fit) and flushed once: a few hundred bins receive every point, so
// direct global atomics serialise (0.3 ms per sub-clip at 414 720 points; profiles/r02_bench_n8_before_stitch_fix.json)
__global__ void __launch_bounds__(kStitchThreads) stitch_hist_kernel(const StitchArgs a, const StitchFrames f,
                                                                     int smem_bins) {
    extern __shared__ int s_hist[];
    __shared__ long long starts[kStitchMaxFrames + 1];
    const int nb = a.max_instances + 2;
    const int pf_bins = nb * f.n_frames;
    const int joint_bins = (a.max_labels + 1) * nb;
    const bool privat = smem_bins >= pf_bins + joint_bins;
    int* pf = privat ? s_hist : a.per_frame;
    int* jt = privat ? s_hist + pf_bins : a.joint;
    if (privat)
        for (int i = threadIdx.x; i < pf_bins + joint_bins; i += kStitchThreads) s_hist[i] = 0;
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int j = 0; j < f.n_frames; ++j) { starts[j] = s; s += a.frame_counts[j]; }
        starts[f.n_frames] = s;
    }
    __syncthreads();
    const long long total = starts[f.n_frames];
    const int k = *a.k_dev;
    int err = 0;
    const long long span = 1ll * gridDim.x * kStitchThreads;
    for (long long p0 = blockIdx.x * 1ll * kStitchThreads; p0 < total; p0 += span) {      // warp-uniform trip count
        const long long p = p0 + threadIdx.x;
        bool ok = p < total;
        int j = 0, lbin = 0;
        if (ok) {
            j = find_slot(starts, f.n_frames, p);
            const long long l = a.labels[p];
            if (l == 0 || l > k || l < -1) { err |= kErrLabelRange; ok = false; }
            else lbin = l < 0 ? 0 : static_cast<int>(l);
        }
        hist_add(pf, lbin * f.n_frames + j, ok);
        bool ok2 = ok && !a.is_first && f.overlap[j];
        int gbin = 0;
        if (ok2) {
            const int t = f.frame[j];
            if (a.frame_count[t] != a.frame_counts[j]) { err |= kErrOverlapSize; ok2 = false; }
            else {
                const long long e = a.frame_labels[t * a.frame_capacity + (p - starts[j])];
                if (e == 0 || e > a.max_labels || e < -1) { err |= kErrLabelRange; ok2 = false; }
                else gbin = e < 0 ? 0 : static_cast<int>(e);
            }
        }
        hist_add(jt, gbin * nb + lbin, ok2);
    }
    if (privat) {
        __syncthreads();
        for (int i = threadIdx.x; i < pf_bins; i += kStitchThreads)
            if (s_hist[i] != 0) atomicAdd(a.per_frame + i, s_hist[i]);
        for (int i = threadIdx.x; i < joint_bins; i += kStitchThreads)
            if (s_hist[pf_bins + i] != 0) atomicAdd(a.joint + i, s_hist[pf_bins + i]);
    }
    if (err) atomicOr(&a.state[2], err);
}

constexpr int kAssignThreads = 128;
constexpr int kAssignMaxLabels = 4095;          // s_size_a lives in shared memory

struct AssignSmem {
    double cost[kAssocMaxSide * kAssocMaxSide];
    LsapScratch lsap;
    long long size_a[kAssignMaxLabels + 1];
    long long size_b[kAssocMaxSide + 2];
    long long u1[kAssocMaxSide + 1], u2[kAssocMaxSide + 1];
    int c1, c2, err, highest;
};

// One block.  The sequential parts (set ordering, assignment) run in thread 0 on shared memory; everything that only
// moves or reduces table entries (row / column sums of the joint histogram, the cost matrix, the per-frame statistics)
// is spread over the block -- a single thread walking thousands of cold global loads made this kernel the serial
// section of the multi-GPU stitch (0.5 ms per sub-clip; profiles/r02_bench_n8_before_stitch_fix.json).
__global__ void __launch_bounds__(kAssignThreads) stitch_assign_kernel(const StitchArgs a, const StitchFrames f) {
    extern __shared__ __align__(16) unsigned char assign_smem_raw[];
    AssignSmem& sm = *reinterpret_cast<AssignSmem*>(assign_smem_raw);
    const int tid = threadIdx.x;
    const int k = *a.k_dev;
    const int nb = a.max_instances + 2;
    const long long offset = a.state[0] - 1;
    const int highest0 = a.state[1];
    if (tid == 0) {
        sm.err = 0;
        sm.c1 = sm.c2 = 0;
        sm.highest = highest0;
        if (k > a.max_instances || offset + k > a.max_labels) sm.err |= kErrTooManyLabels;
    }
    for (int l = 1 + tid; l <= a.max_instances + 1; l += kAssignThreads) {
        a.lut_local[l - 1] = offset + l;
        a.lut_assoc[l - 1] = offset + l;
    }
    for (int c = tid; c < a.max_instances; c += kAssignThreads) a.meta_labels[c] = c < k ? offset + c + 1 : -1;
    __syncthreads();
    for (int j = tid; j < f.n_frames; j += kAssignThreads) {
        const int t = f.frame[j];
        if (!a.is_first && f.overlap[j]) {
            if (a.frame_count[t] != a.frame_counts[j]) atomicOr(&sm.err, kErrOverlapSize);
        } else if (a.frame_count[t] >= 0) {
            atomicOr(&sm.err, kErrFrameExists);
        }
    }
    const bool associate = !a.is_first && k <= a.max_instances && offset + k <= a.max_labels;
    if (associate) {
        // row / column sums of the joint histogram (existing id x current local label) over the overlap frames
        for (int g = tid; g <= highest0 && g <= a.max_labels; g += kAssignThreads) {
            long long sum = 0;
            for (int l = 0; l <= k; ++l) sum += a.joint[g * nb + l];
            sm.size_a[g] = sum;
        }
        for (int l = tid; l <= k && l < kAssocMaxSide + 2; l += kAssignThreads) {
            long long sum = 0;
            for (int g = 0; g <= highest0 && g <= a.max_labels; ++g) sum += a.joint[g * nb + l];
            sm.size_b[l] = sum;
        }
    }
    __syncthreads();
    // the two label lists are ordered independently: thread 0 takes the existing ids, thread 32 (another warp) the
    // current ones -- each emulation is a chain of dependent table probes, so running them side by side halves it
    if (associate && (tid == 0 || tid == 32)) {
        long long vals[kAssocMaxSide + 1];
        int n = 0, err = 0;
        if (tid == 0) {      // ids present in the overlap frames, ascending like Tensor.unique(), outliers (-1) first
            for (int g = 0; g <= highest0 && g <= a.max_labels; ++g)
                if (sm.size_a[g] > 0) {
                    if (n >= kAssocMaxSide) { err |= kErrTooManyLabels; break; }
                    vals[n++] = g == 0 ? -1 : g;
                }
        } else {
            for (int l = 0; l <= k && l < kAssocMaxSide + 2; ++l)
                if (sm.size_b[l] > 0) {
                    if (n >= kAssocMaxSide) { err |= kErrTooManyLabels; break; }
                    vals[n++] = l == 0 ? -1 : offset + l;
                }
        }
        int c = 0;
        if (!err) {
            c = pyset_order(vals, n, tid == 0 ? sm.u1 : sm.u2);
            if (c < 0) { err |= kErrTooManyLabels; c = 0; }
            // every current label is > every existing one (offset = next_track_label - 1 >= highest existing label)
            if (tid == 0)
                for (int i = 0; i < c; ++i)
                    if (sm.u1[i] > offset) err |= kErrLabelsOverlap;
        }
        if (tid == 0) sm.c1 = c; else sm.c2 = c;
        if (err) atomicOr(&sm.err, err);
    }
    __syncthreads();
    if (tid == 0 && sm.err) sm.c1 = sm.c2 = 0;
    __syncthreads();
    const int c1 = sm.c1, c2 = sm.c2;
    for (int idx = tid; idx < c1 * c2; idx += kAssignThreads) {
        const int i1 = idx / c2, i2 = idx % c2;
        const int l = static_cast<int>(sm.u2[i2] - offset);
        const long long inter = a.joint[sm.u1[i1] * nb + l];
        const long long uni = sm.size_a[sm.u1[i1]] + sm.size_b[l] - inter;
        const float iou = __fdiv_rn(static_cast<float>(inter), static_cast<float>(uni));   // fp32 tensors in the reference
        const float c32 = static_cast<float>(1.0 - static_cast<double>(iou));            // 1. - iou.item() into float32
        sm.cost[idx] = static_cast<double>(c32);
    }
    __syncthreads();
    if (tid == 0 && c1 > 0 && c2 > 0) {
        int rows[kAssocMaxSide], cols[kAssocMaxSide];
        const int pairs = lsap_solve(sm.cost, c1, c2, rows, cols, sm.lsap);
        if (pairs < 0) sm.err |= kErrAssignment;
        for (int q = 0; q < pairs; ++q) {
            const long long associated = sm.u1[rows[q]], current = sm.u2[cols[q]];
            a.lut_assoc[current - offset - 1] = associated;
            for (int c = 0; c < k; ++c)                   // meta_info['instance_labels'].index(current)
                if (a.meta_labels[c] == current) { a.meta_labels[c] = associated; break; }
        }
    }
    __syncthreads();
    // TrackContainer bookkeeping for the frames added by this sub-clip (one thread per (frame, label bin))
    const int bins = (k < a.max_instances ? k : a.max_instances) + 1;
    for (int idx = tid; idx < f.n_frames * bins; idx += kAssignThreads) {
        const int j = idx / bins, lbin = idx % bins;
        if (!a.is_first && f.overlap[j]) continue;
        const int t = f.frame[j];
        const int c = a.per_frame[lbin * f.n_frames + j];
        if (c == 0) continue;
        const long long label = lbin == 0 ? -1 : a.lut_assoc[lbin - 1];
        if (label > a.max_labels) { atomicOr(&sm.err, kErrTooManyLabels); continue; }
        atomicAdd(reinterpret_cast<unsigned long long*>(a.track_counts + label + 1), static_cast<unsigned long long>(c));
        atomicMin(a.span_lo + label + 1, t);
        atomicMax(a.span_hi + label + 1, t);
        atomicMax(&sm.highest, static_cast<int>(label));
    }
    for (int j = tid; j < f.n_frames; j += kAssignThreads)
        if (a.is_first || !f.overlap[j]) a.frame_count[f.frame[j]] = a.frame_counts[j];
    __syncthreads();
    if (tid == 0) {
        a.state[1] = sm.highest;
        a.state[0] = sm.highest + 1;
        a.state[3] += 1;
        if (sm.err) atomicOr(&a.state[2], sm.err);
    }
}

__global__ void __launch_bounds__(kStitchThreads) stitch_relabel_kernel(const StitchArgs a, const StitchFrames f) {
    __shared__ long long starts[kStitchMaxFrames + 1];
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int j = 0; j < f.n_frames; ++j) { starts[j] = s; s += a.frame_counts[j]; }
        starts[f.n_frames] = s;
    }
    __syncthreads();
    const long long total = starts[f.n_frames];
    for (long long p = blockIdx.x * 1ll * kStitchThreads + threadIdx.x; p < total; p += 1ll * gridDim.x * kStitchThreads) {
        const int j = find_slot(starts, f.n_frames, p);
        const long long l = a.labels[p];
        const bool over = !a.is_first && f.overlap[j];
        long long out = -1;
        if (l >= 1 && l <= a.max_instances + 1) out = over ? a.lut_local[l - 1] : a.lut_assoc[l - 1];
        a.labels[p] = out;
        if (!over) a.frame_labels[f.frame[j] * a.frame_capacity + (p - starts[j])] = out;
    }
}

size_t stitch_ws_layout(int max_instances, int n_frames, int max_labels, size_t* off) {
    size_t o = 0;
    off[0] = o; o = align_up(o + sizeof(int) * (static_cast<size_t>(max_labels) + 1) * (max_instances + 2), 256);   // joint
    off[1] = o; o = align_up(o + sizeof(int) * static_cast<size_t>(max_instances + 2) * n_frames, 256);              // per_frame
    off[2] = o; o = align_up(o + sizeof(long long) * (max_instances + 1), 256);                                       // lut_assoc
    off[3] = o; o = align_up(o + sizeof(long long) * (max_instances + 1), 256);                                       // lut_local
    off[4] = o; o = align_up(o + sizeof(double) * kAssocMaxSide * kAssocMaxSide, 256);                                // cost
    off[5] = o; o = align_up(o + sizeof(LsapScratch), 256);
    return o;
}

}  // namespace
}  // namespace stemseg

extern "C" size_t stemseg_stitch_workspace_bytes(int32_t max_instances, int32_t n_frames, int32_t max_labels) {
    size_t off[6];
    return stitch_ws_layout(max_instances, n_frames, max_labels, off);
}

extern "C" int32_t stemseg_stitch_subclip(int64_t* labels, int64_t capacity, const int32_t* frame_counts_dev,
                                          const int32_t* k_dev, const int32_t* frames_host, const int32_t* overlap_host,
                                          int32_t n_frames, int32_t is_first, int32_t max_instances,
                                          int64_t* frame_labels, int32_t* frame_count, int64_t frame_capacity,
                                          int32_t num_frames, int32_t* state, int64_t* track_counts, int32_t* span_lo,
                                          int32_t* span_hi, int32_t max_labels, int64_t* meta_labels_out,
                                          void* workspace, size_t ws_bytes, void* stream_) {
    SS_REQUIRE(labels && frame_counts_dev && k_dev && frames_host && overlap_host && frame_labels && frame_count && state &&
                   track_counts && span_lo && span_hi && meta_labels_out && workspace,
               "stitch_subclip: null pointer");
    SS_REQUIRE(n_frames >= 1 && n_frames <= kStitchMaxFrames, "stitch_subclip: n_frames %d out of range [1, %d]", n_frames,
               kStitchMaxFrames);
    SS_REQUIRE(max_instances >= 1 && max_instances + 2 <= kAssocMaxSide, "stitch_subclip: max_instances %d out of range",
               max_instances);
    SS_REQUIRE(max_labels >= max_instances && capacity >= 0 && frame_capacity >= 1 && num_frames >= 1,
               "stitch_subclip: bad sizes");
    SS_REQUIRE(max_labels <= kAssignMaxLabels, "stitch_subclip: max_labels %d > %d", max_labels, kAssignMaxLabels);
    size_t off[6];
    const size_t need = stitch_ws_layout(max_instances, n_frames, max_labels, off);
    SS_REQUIRE(ws_bytes >= need, "stitch_subclip: workspace too small (%zu < %zu)", ws_bytes, need);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    StitchFrames f;
    f.n_frames = n_frames;
    for (int j = 0; j < n_frames; ++j) {
        SS_REQUIRE(frames_host[j] >= 0 && frames_host[j] < num_frames, "stitch_subclip: frame %d outside the video",
                   frames_host[j]);
        f.frame[j] = frames_host[j];
        f.overlap[j] = overlap_host[j] ? 1 : 0;
    }
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    StitchArgs a;
    a.labels = reinterpret_cast<long long*>(labels);
    a.frame_counts = frame_counts_dev;
    a.k_dev = k_dev;
    a.max_instances = max_instances;
    a.frame_labels = reinterpret_cast<long long*>(frame_labels);
    a.frame_count = frame_count;
    a.frame_capacity = frame_capacity;
    a.state = state;
    a.track_counts = reinterpret_cast<long long*>(track_counts);
    a.span_lo = span_lo;
    a.span_hi = span_hi;
    a.max_labels = max_labels;
    a.meta_labels = reinterpret_cast<long long*>(meta_labels_out);
    a.joint = reinterpret_cast<int*>(ws + off[0]);
    a.per_frame = reinterpret_cast<int*>(ws + off[1]);
    a.lut_assoc = reinterpret_cast<long long*>(ws + off[2]);
    a.lut_local = reinterpret_cast<long long*>(ws + off[3]);
    a.cost = reinterpret_cast<double*>(ws + off[4]);
    a.lsap = reinterpret_cast<LsapScratch*>(ws + off[5]);
    a.is_first = is_first ? 1 : 0;
    SS_CUDA_OK(cudaMemsetAsync(ws + off[0], 0, off[2] - off[0], stream));          // joint + per_frame
    long long blocks = (capacity + kStitchThreads * 4 - 1) / (kStitchThreads * 4);
    const long long cap = 2ll * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const long long hist_bins = 1ll * (max_instances + 2) * n_frames + 1ll * (max_labels + 1) * (max_instances + 2);
    const int smem_bins = hist_bins * 4 <= 40 * 1024 ? static_cast<int>(hist_bins) : 0;
    stitch_hist_kernel<<<static_cast<unsigned>(blocks), kStitchThreads, smem_bins * sizeof(int), stream>>>(a, f, smem_bins);
    SS_CUDA_OK(cudaGetLastError());
    SS_CUDA_OK(cudaFuncSetAttribute(stitch_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(sizeof(AssignSmem))));
    stitch_assign_kernel<<<1, kAssignThreads, sizeof(AssignSmem), stream>>>(a, f);
    SS_CUDA_OK(cudaGetLastError());
    stitch_relabel_kernel<<<static_cast<unsigned>(blocks), kStitchThreads, 0, stream>>>(a, f);
    SS_CUDA_OK(cudaGetLastError());
    return STEMSEG_OK;
}
